"""ctypes binding of libslfp_b200.so (the C ABI declared in include/slfp_b200.h).

There is no CPU fallback and no other backend: if the library is missing or a tensor is not on a
CUDA device, the callers raise.
"""
import ctypes
import os

import torch

from .build import LIB_PATH, source_hash

FMT_SFP33, FMT_SLFP34_ACT, FMT_SLFP34_WGT, FMT_SFP44_OUT, FMT_SLFP34_RELU, FMT_SFP33_RELU, FMT_SFP33_SFAST, FMT_E4M3, FMT_F16Q = 0, 1, 2, 3, 4, 5, 6, 7, 8
ACT_STL, ACT_SWISH, ACT_SIGMOID = 0, 1, 2
SGD_NORMAL, SGD_DSGD, SGD_SSGD = 0, 1, 2
Q_LAYEROUT_ZERO_IS_ZERO = 1
CONV_SPLIT_OPERANDS = 1
CONV_E4M3_OPERANDS = 2
CONV_FOLD_W = 4

c_vp, c_sz, c_f, c_i, c_ll, c_d = (ctypes.c_void_p, ctypes.c_size_t, ctypes.c_float, ctypes.c_int,
                                   ctypes.c_longlong, ctypes.c_double)


class SlfpConvDesc(ctypes.Structure):
    _fields_ = [(n, c_i) for n in ("n", "h", "w", "c", "c_phys", "k", "r", "s", "stride_h", "stride_w", "pad_h",
                                   "pad_w", "dil_h", "dil_w", "groups", "fmt", "pad_h_extra", "pad_w_extra", "flags")]


class SlfpEpilogue(ctypes.Structure):
    _fields_ = [("bias_q", c_vp), ("post_a", c_f), ("post_b", c_f), ("ch_scale", c_vp), ("ch_shift", c_vp),
                ("residual", c_vp), ("residual_f16", c_i), ("relu", c_i), ("y_f32", c_vp), ("y_f16", c_vp),
                ("y_codes", c_vp), ("next_k_div", c_f), ("next_fmt", c_i), ("k_phys_out", c_i),
                ("y_codes2", c_vp), ("next_k_div2", c_f), ("ch_mul", c_vp), ("ch_add", c_vp), ("layerout", c_i), ("store_f16", c_i)]


class SlfpGatherChan(ctypes.Structure):
    _fields_ = [("src", c_vp), ("stride", c_i), ("ch", c_i)]


class SlfpGatherRun(ctypes.Structure):
    _fields_ = [("src", c_vp), ("stride", c_i), ("ch0", c_i), ("len", c_i), ("dst_start", c_i), ("dst_step", c_i),
                ("magic", ctypes.c_uint), ("shift", ctypes.c_uint)]


class SlfpWeightJob(ctypes.Structure):
    _fields_ = [("desc", ctypes.POINTER(SlfpConvDesc)), ("w", c_vp), ("w_stride", c_ll * 4), ("kw", c_f), ("w_f16", c_vp),
                ("w_codes", c_vp), ("out_pitch", c_sz), ("out_offset", c_sz), ("row_scale", c_vp), ("lo_offset", c_sz)]


_SIGS = {
    "slfp_version": (c_i, []),
    "slfp_last_error": (ctypes.c_char_p, []),
    "slfp_build_id": (ctypes.c_char_p, []),
    "slfp_quantize_dyn_f32": (c_i, [c_vp, c_sz, c_vp, c_d, c_i, ctypes.c_uint, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "slfp_quantize_f32": (c_i, [c_vp, c_sz, c_f, c_i, ctypes.c_uint, c_vp, c_vp, c_vp, c_vp]),
    "slfp_quantize_nhwc_f32": (c_i, [c_vp, c_sz, c_i, c_i, c_f, c_i, c_vp, c_vp]),
    "slfp_quantize_nchw_f32": (c_i, [c_vp, c_i, c_i, c_sz, c_i, c_f, c_i, c_vp, c_vp]),
    "slfp_quantize_nchw_s2d_f32": (c_i, [c_vp, c_i, c_i, c_i, c_i, c_i, c_f, c_i, c_vp, c_vp]),
    "slfp_quantize_nchw_s2d_f16q": (c_i, [c_vp, c_i, c_i, c_i, c_f, c_i, c_i, c_i, c_i, c_i, c_vp, c_vp]),
    "slfp_quantize_nchw_im2col3x3_f16q": (c_i, [c_vp, c_i, c_i, c_i, c_f, c_i, c_vp, c_vp]),
    "slfp_dequantize": (c_i, [c_vp, c_sz, c_i, c_vp, c_vp]),
    "slfp_gather_quantize_f16": (c_i, [c_vp, c_sz, c_i, c_i, c_f, c_i, c_vp, c_vp]),
    "slfp_gather_quantize_runs_f16": (c_i, [c_vp, c_i, c_sz, c_i, c_i, c_i, c_f, c_i, c_vp, c_vp]),
    "slfp_magic_u32": (None, [ctypes.c_uint, ctypes.POINTER(ctypes.c_uint), ctypes.POINTER(ctypes.c_uint)]),
    "slfp_absmax_f32": (c_i, [c_vp, c_sz, c_vp, c_i, c_vp]),
    "slfp_conv_wpitch": (c_sz, [ctypes.POINTER(SlfpConvDesc)]),
    "slfp_prepare_weights": (c_i, [ctypes.POINTER(SlfpConvDesc), c_vp, c_ll, c_ll, c_ll, c_ll, c_f, c_i, c_vp, c_vp,
                                   c_vp, c_vp]),
    "slfp_prepare_weights_jobs": (c_i, [c_i, c_vp, c_i, c_vp]),
    "slfp_conv2d_fwd_dual": (c_i, [ctypes.POINTER(SlfpConvDesc), c_vp, ctypes.POINTER(SlfpConvDesc), c_vp, c_vp,
                                   ctypes.POINTER(SlfpEpilogue), c_vp]),
    "slfp_conv2d_fwd": (c_i, [ctypes.POINTER(SlfpConvDesc), c_vp, c_vp, ctypes.POINTER(SlfpEpilogue), c_vp]),
    "slfp_conv2d_bwd": (c_i, [ctypes.POINTER(SlfpConvDesc), c_vp, c_vp, c_vp, c_i, c_f, c_f, c_vp, c_vp, c_ll, c_ll,
                              c_ll, c_ll, c_vp, c_vp]),
    "slfp_conv2d_bwd_workspace_size": (c_sz, [ctypes.POINTER(SlfpConvDesc), c_i, c_i]),
    "slfp_conv2d_bwd_ws": (c_i, [ctypes.POINTER(SlfpConvDesc), c_vp, c_vp, c_vp, c_i, c_f, c_f, c_vp, c_vp, c_ll, c_ll,
                                 c_ll, c_ll, c_vp, c_vp, c_sz, c_vp]),
    "slfp_conv2d_bwd_ws_absmax": (c_i, [ctypes.POINTER(SlfpConvDesc), c_vp, c_vp, c_vp, c_vp, c_i, c_f, c_f, c_vp, c_vp, c_ll, c_ll,
                                        c_ll, c_ll, c_vp, c_vp, c_sz, c_vp]),
    "slfp_act_fwd": (c_i, [c_vp, c_sz, c_i, c_vp, c_vp]),
    "slfp_act_bwd": (c_i, [c_vp, c_vp, c_sz, c_i, c_vp, c_vp]),
    "slfp_sgd_step": (c_i, [c_i, c_vp, c_vp, c_vp, c_vp, c_i, c_i, c_d, c_d, c_d, c_d, c_i, c_i, c_vp]),
    "slfp_maxpool_codes": (c_i, [c_vp, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_vp, c_vp]),
    "slfp_avgpool_nhwc": (c_i, [c_vp, c_i, c_i, c_i, c_i, c_vp, c_vp]),
    "slfp_avgpool_quantize_nhwc_f16": (c_i, [c_vp, c_i, c_i, c_i, c_vp, c_f, c_i, c_vp, c_vp]),
    "slfp_bn_act_workspace_floats": (c_sz, [c_i]),
    "slfp_bn_act_fwd_train": (c_i, [c_vp, c_sz, c_i, c_vp, c_vp, c_vp, c_i, c_f, c_f, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "slfp_bn_act_fwd_train_quant": (c_i, [c_vp, c_sz, c_i, c_vp, c_vp, c_vp, c_f, c_f, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i, c_i,
                                          ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_void_p), c_vp]),
    "slfp_bn_act_bwd": (c_i, [c_vp, c_vp, c_vp, c_sz, c_i, c_vp, c_vp, c_vp, c_vp, c_i, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "slfp_maxpool3x3s2_fwd_f32": (c_i, [c_vp, c_i, c_i, c_i, c_i, c_vp, c_vp, c_vp]),
    "slfp_maxpool3x3s2_bwd_f32": (c_i, [c_vp, c_vp, c_i, c_i, c_i, c_i, c_vp, c_vp]),
    "slfp_debug_set_buffer": (c_i, [c_vp]),
    "slfp_quantize_host_f32": (c_i, [c_vp, c_sz, c_f, c_i, c_vp, c_vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)
_lib = None

# Every entry point that launches kernels on the caller's stream.  The proxy below counts those calls
# (bench.py reports them as `gpu_launches`) and, when a profile dict is installed, brackets each call
# with CUDA events on the launching stream (bench.py's per-kernel roofline pass).
_LAUNCHING = {"slfp_gather_quantize_runs_f16", "slfp_gather_quantize_f16", "slfp_quantize_dyn_f32", "slfp_prepare_weights_jobs", "slfp_conv2d_fwd_dual", "slfp_quantize_f32", "slfp_quantize_nhwc_f32", "slfp_dequantize", "slfp_absmax_f32", "slfp_prepare_weights",
              "slfp_conv2d_fwd", "slfp_conv2d_bwd", "slfp_conv2d_bwd_ws", "slfp_conv2d_bwd_ws_absmax", "slfp_act_fwd", "slfp_act_bwd", "slfp_sgd_step", "slfp_maxpool_codes",
              "slfp_avgpool_nhwc", "slfp_avgpool_quantize_nhwc_f16", "slfp_bn_act_fwd_train", "slfp_bn_act_fwd_train_quant", "slfp_bn_act_bwd", "slfp_maxpool3x3s2_fwd_f32", "slfp_maxpool3x3s2_bwd_f32", "slfp_quantize_nchw_f32", "slfp_quantize_nchw_s2d_f32", "slfp_quantize_nchw_s2d_f16q", "slfp_quantize_nchw_im2col3x3_f16q"}
launch_count = 0
# (data_ptr, numel, device scalar) of the most recent gradient tensor whose producer tracked max |.| while writing it
# (utils/bn_act.py backward); the Conv2d_Q backward that receives exactly that tensor uses it and clears the slot
pending_grad_absmax = None
profile = None          # None, or {name: [(start_event, end_event, tag), ...]}
profile_tag = None


class _Lib:
    def __init__(self, handle):
        self._h = handle
        for name in _SIGS:
            fn = getattr(handle, name)
            setattr(self, name, self._wrap(name, fn) if name in _LAUNCHING else fn)

    @staticmethod
    def _wrap(name, fn):
        def call(*args):
            global launch_count
            launch_count += 1
            if profile is None:
                return fn(*args)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rc = fn(*args)
            b.record()
            profile.setdefault(name, []).append((a, b, launch_count))      # launch_count orders events across entry points
            return rc
        return call


def lib():
    """Load the native library; fail loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m cnns_slfp_quantization_b200.build` "
                "(nvcc, sm_100a).  There is no CPU / PyTorch fallback for the SLFP hot path.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        built, want = handle.slfp_build_id().decode(), source_hash()
        if built != want and not os.environ.get("SLFP_ALLOW_STALE_LIB"):
            raise ImportError(f"{LIB_PATH} was built from other sources (build id {built}, sources {want}): rebuild with "
                              "`python -m cnns_slfp_quantization_b200.build`")
        _lib = _Lib(handle)
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().slfp_last_error()
        raise RuntimeError(f"libslfp_b200 error {rc}: {msg.decode() if msg else ''}")


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def require_cuda(t, who):
    if not t.is_cuda:
        raise RuntimeError(f"{who}: expected a CUDA tensor (the SLFP hot path has hand-written sm_100a kernels only; "
                           "there is no CPU fallback)")
    if t.dtype != torch.float32:
        raise TypeError(f"{who}: expected float32, got {t.dtype}")


def fmt_for(q_bit, kind):
    """q_bit in {7, 8}; kind in {'act', 'weight'} -> storage format of the 8-bit codes."""
    if q_bit == 7:
        return FMT_SFP33
    if q_bit == 8:
        return FMT_SLFP34_ACT if kind == "act" else FMT_SLFP34_WGT
    raise ValueError(f"no 8-bit code format for q_bit={q_bit}")


def relu_fmt(fmt):
    """The unsigned post-ReLU code format with the same grid as the signed activation format `fmt`."""
    return FMT_SFP33_RELU if fmt in (FMT_SFP33, FMT_SFP33_RELU) else FMT_SLFP34_RELU


def dense_flat(t):
    """A tensor whose storage can be processed as a flat array in memory order (any dense layout)."""
    if t.is_contiguous():
        return t
    if t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last):
        return t
    return t.contiguous()
