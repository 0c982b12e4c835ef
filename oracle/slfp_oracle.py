"""numpy front-end of oracle/slfp_oracle.c (ctypes).  TEST INFRASTRUCTURE ONLY.

Every function cites the reference lines it restates (paths under /root/reference).
Parity pin: tests/golden/*.npz generated from the reference itself by
tests/golden/make_golden.py; see the header of slfp_oracle.c.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libslfp_oracle.so")

FMT_SFP33, FMT_SLFP34_ACT, FMT_SLFP34_WGT, FMT_SFP44_OUT = 0, 1, 2, 3
ACT_STL, ACT_SWISH, ACT_SIGMOID = 0, 1, 2
SGD_NORMAL, SGD_DSGD, SGD_SSGD = 0, 1, 2


def build(force=False):
    src = os.path.join(_HERE, "slfp_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "_build/libslfp_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.slfp_oracle_decode1.restype = ctypes.c_float
        _lib.slfp_oracle_absmax.restype = ctypes.c_float
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t)) if a is not None else None


def fmt_for(q_bit, kind):
    """kind in {'act','weight','layerout'}; q_bit in {7,8} (sfp_quant.py:14,32,63,80,111)."""
    if kind == "layerout":
        return FMT_SFP44_OUT
    if q_bit == 7:
        return FMT_SFP33
    if q_bit == 8:
        return FMT_SLFP34_ACT if kind == "act" else FMT_SLFP34_WGT
    raise ValueError(q_bit)


def quantize(x, fmt, kdiv=1.0, bugcompat=True, want_codes=True):
    """(x / float32(kdiv)) -> (codes uint8 | None, fake-quant float32).

    sfp_quant.py:14-47 (weights), :63-96 (activations), :111-126 (layerout);
    pre-scale division conv2d_func.py:21-22.
    """
    x = np.ascontiguousarray(x, dtype=np.float32)
    fq = np.empty_like(x)
    codes = np.empty(x.shape, dtype=np.uint8) if (want_codes and fmt != FMT_SFP44_OUT) else None
    lib().slfp_oracle_quantize(_p(x, ctypes.c_float), ctypes.c_size_t(x.size),
                               ctypes.c_float(np.float32(kdiv)), ctypes.c_int(fmt),
                               ctypes.c_int(1 if bugcompat else 0),
                               _p(codes, ctypes.c_uint8), _p(fq, ctypes.c_float))
    return codes, fq


def decode(codes, fmt):
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    out = np.empty(codes.shape, dtype=np.float32)
    lib().slfp_oracle_decode(_p(codes, ctypes.c_uint8), ctypes.c_size_t(codes.size),
                             ctypes.c_int(fmt), _p(out, ctypes.c_float))
    return out


def decode_relu(codes, sfp33=False):
    """Value of the fused pipeline's unsigned post-ReLU codes (include/slfp_b200.h, SLFP_FMT_*_RELU), restated
    from the reference's grids.  c = 0 -> 0 (sfp_quant.py:92 / :74: |x| < 0.0625 -> 1e-10, which is 0 as a float16
    tensor-core operand); u = c - 1 = E * 2H + h (H = 16 / 8 mantissa steps, h a HALF-step index): E = 0 -> 0.125
    (:93 / :75); E >= 1 -> mantissa index i = (h + 1) >> 1 = round(H m) of :88 / :69 (ties up), pushed through the
    log converter of :89 for SLFP (log index = i + [2 <= i <= 14]); clamped to the top value (:95 / :77)."""
    c = np.asarray(codes).astype(np.int64)
    u = np.maximum(c - 1, 0)
    if sfp33:
        E, i = u >> 4, ((u & 15) + 1) >> 1
        val = np.minimum((1.0 + i / 8.0) * np.exp2(E - 4.0), 15.0)
    else:
        E, i = u >> 5, ((u & 31) + 1) >> 1
        L = i + ((i >= 2) & (i <= 14))
        # float32(2^(j/16)) as the reference's pow() returns it (tests/golden/make_golden.log); index 16 = 2.0
        tab = np.array([0x3f800000, 0x3f85aac3, 0x3f8b95c2, 0x3f91c3d3, 0x3f9837f0, 0x3f9ef532, 0x3fa5fed7, 0x3fad583f,
                        0x3fb504f3, 0x3fbd08a4, 0x3fc5672a, 0x3fce248c, 0x3fd744fd, 0x3fe0ccdf, 0x3feac0c7, 0x3ff5257d,
                        0x40000000], dtype=np.uint32).view(np.float32).astype(np.float64)
        val = np.minimum(tab[L] * np.exp2(E - 4.0), tab[15] * 8.0)
    val = np.where(E == 0, 0.125, val)
    return np.where(c == 0, 0.0, val).astype(np.float32)


def absmax(x):
    """max(|x|) -- cifar100_train_eval.py:261-271 (calibration)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    return float(lib().slfp_oracle_absmax(_p(x, ctypes.c_float), ctypes.c_size_t(x.size)))


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def conv2d_q(xq, wq, bias_q, stride, padding, dilation, groups, ka, kw):
    """F.conv2d(xq, wq, bias_q, ...) * Ka * Kw with double accumulation (conv2d_func.py:23-24)."""
    xq = np.ascontiguousarray(xq, dtype=np.float32)
    wq = np.ascontiguousarray(wq, dtype=np.float32)
    if bias_q is not None:
        bias_q = np.ascontiguousarray(bias_q, dtype=np.float32)
    N, C, H, W = xq.shape
    O, Cg, R, S = wq.shape
    assert Cg * groups == C
    (sh, sw), (ph, pw), (dh, dw) = _pair(stride), _pair(padding), _pair(dilation)
    Ho = (H + 2 * ph - dh * (R - 1) - 1) // sh + 1
    Wo = (W + 2 * pw - dw * (S - 1) - 1) // sw + 1
    y = np.empty((N, O, Ho, Wo), dtype=np.float32)
    lib().slfp_oracle_conv2d(_p(xq, ctypes.c_float), _p(wq, ctypes.c_float), _p(bias_q, ctypes.c_float),
                             N, C, H, W, O, R, S, sh, sw, ph, pw, dh, dw, groups,
                             ctypes.c_float(np.float32(ka)), ctypes.c_float(np.float32(kw)),
                             _p(y, ctypes.c_float))
    return y


def conv2d_q_bwd(xq, wq, gy, stride, padding, dilation, groups, ka, kw, with_bias=False):
    """Module-level gradients (dL/dinput, dL/dweight, dL/dbias) of Conv2d_Q.forward given gy.

    conv2d_func.py:20-25 differentiated with the identity STE of sfp_quant.py:50-53.
    """
    xq = np.ascontiguousarray(xq, dtype=np.float32)
    wq = np.ascontiguousarray(wq, dtype=np.float32)
    gy = np.ascontiguousarray(gy, dtype=np.float32)
    N, C, H, W = xq.shape
    O, Cg, R, S = wq.shape
    (sh, sw), (ph, pw), (dh, dw) = _pair(stride), _pair(padding), _pair(dilation)
    dx = np.zeros(xq.shape, dtype=np.float64)
    dwt = np.zeros(wq.shape, dtype=np.float64)
    db = np.zeros((O,), dtype=np.float64) if with_bias else None
    lib().slfp_oracle_conv2d_bwd(_p(xq, ctypes.c_float), _p(wq, ctypes.c_float), _p(gy, ctypes.c_float),
                                 N, C, H, W, O, R, S, sh, sw, ph, pw, dh, dw, groups,
                                 ctypes.c_float(np.float32(ka)), ctypes.c_float(np.float32(kw)),
                                 _p(dx, ctypes.c_double), _p(dwt, ctypes.c_double), _p(db, ctypes.c_double))
    return dx, dwt, db


def conv2d_Q_forward(x, w, bias, ka, kw, q_bit, stride=1, padding=0, dilation=1, groups=1):
    """Whole Conv2d_Q.forward (conv2d_func.py:20-25 / :41-47): returns (input_q, weight_q, output)."""
    if q_bit == 32:
        xq = (np.asarray(x, np.float32) / np.float32(ka)).astype(np.float32)
        wq = (np.asarray(w, np.float32) / np.float32(kw)).astype(np.float32)
    else:
        _, xq = quantize(x, fmt_for(q_bit, "act"), ka, want_codes=False)
        _, wq = quantize(w, fmt_for(q_bit, "weight"), kw, want_codes=False)
    bq = None
    if bias is not None:
        bq = ((np.asarray(bias, np.float32) / np.float32(ka)) / np.float32(kw)).astype(np.float32)
    return xq, wq, conv2d_q(xq, wq, bq, stride, padding, dilation, groups, ka, kw)


def linear_Q_forward(x, w, bias, ka, kw, q_bit):
    """Linear_Q.forward (conv2d_func.py:60-65): bias/Kw/Ka, then *Kw*Ka."""
    x = np.asarray(x, np.float32)
    w = np.asarray(w, np.float32)
    if q_bit == 32:
        xq, wq = x / np.float32(ka), w / np.float32(kw)
    else:
        _, xq = quantize(x, fmt_for(q_bit, "act"), ka, want_codes=False)
        _, wq = quantize(w, fmt_for(q_bit, "weight"), kw, want_codes=False)
    acc = xq.astype(np.float64) @ wq.astype(np.float64).T
    out = acc.astype(np.float32)
    if bias is not None:
        out = out + ((np.asarray(bias, np.float32) / np.float32(kw)) / np.float32(ka)).astype(np.float32)
    out = (out * np.float32(kw)).astype(np.float32) * np.float32(ka)
    return xq, wq, out.astype(np.float32)


def act_fwd(x, kind):
    """activation_func.py:10 (STL), :30-32 (Swish), :34-36 (Sigmoid)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    lib().slfp_oracle_act_fwd(_p(x, ctypes.c_float), ctypes.c_size_t(x.size), ctypes.c_int(kind),
                              _p(y, ctypes.c_float))
    return y


def act_bwd(x, gy, kind):
    """activation_func.py:16 (STL backward clips by |grad|); autograd of Swish/Sigmoid."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    gy = np.ascontiguousarray(gy, dtype=np.float32)
    gx = np.empty_like(x)
    lib().slfp_oracle_act_bwd(_p(x, ctypes.c_float), _p(gy, ctypes.c_float), ctypes.c_size_t(x.size),
                              ctypes.c_int(kind), _p(gx, ctypes.c_float))
    return gx


def sgd_step(p, grad, buf, mode, q_bit, lr, momentum=0.0, dampening=0.0, weight_decay=0.0,
             nesterov=False, first_step=True):
    """In-place optimizer step on numpy arrays (optimizer.py:41-64 / :109-131 / :165-190)."""
    assert p.dtype == np.float32 and grad.dtype == np.float32 and p.flags.c_contiguous
    qfmt = -1 if q_bit == 32 else fmt_for(q_bit, "weight")
    if buf is None:
        buf = np.zeros_like(p)
    lib().slfp_oracle_sgd_step(_p(p, ctypes.c_float), _p(grad, ctypes.c_float), _p(buf, ctypes.c_float),
                               ctypes.c_size_t(p.size), ctypes.c_int(mode), ctypes.c_int(qfmt),
                               ctypes.c_double(lr), ctypes.c_double(momentum), ctypes.c_double(dampening),
                               ctypes.c_double(weight_decay), ctypes.c_int(1 if nesterov else 0),
                               ctypes.c_int(1 if first_step else 0))
    return buf
