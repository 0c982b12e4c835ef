"""Helpers shared by the -m gpu tests: thin wrappers that call the C ABI on torch CUDA tensors."""
import ctypes

import numpy as np
import torch

from cnns_slfp_quantization_b200 import _native as nv


def dev():
    return torch.device("cuda:0")


def quantize_gpu(x_np, fmt, kdiv=1.0, flags=0, want_codes=True, want_f16=False):
    x = torch.from_numpy(np.ascontiguousarray(x_np, np.float32)).to(dev())
    codes = torch.empty(x.shape, dtype=torch.uint8, device=dev()) if (want_codes and fmt != nv.FMT_SFP44_OUT) else None
    fq = torch.empty_like(x)
    f16 = torch.empty(x.shape, dtype=torch.float16, device=dev()) if want_f16 else None
    nv.check(nv.lib().slfp_quantize_f32(x.data_ptr(), x.numel(), float(np.float32(kdiv)), fmt, flags, nv.ptr(codes),
                                        fq.data_ptr(), nv.ptr(f16), nv.stream()))
    torch.cuda.synchronize()
    return (None if codes is None else codes.cpu().numpy()), fq.cpu().numpy(), (None if f16 is None else f16.cpu().numpy())


def conv_fwd_gpu(x_nchw, w_oihw, bias_q, ka, kw, q_bit, stride, pad, dil, groups, epi_kwargs=None):
    """Raw C-ABI path: quantize -> prepare weights -> conv.  Returns dict of numpy outputs (NCHW)."""
    lib = nv.lib()
    x = torch.from_numpy(np.ascontiguousarray(x_nchw, np.float32)).to(dev()).permute(0, 2, 3, 1).contiguous()
    w = torch.from_numpy(np.ascontiguousarray(w_oihw, np.float32)).to(dev())
    N, H, W, C = x.shape
    K, Cg, R, S = w.shape
    dense = groups == 1
    Cp = (4 if C <= 4 else (C + 15) // 16 * 16) if dense else (C + 3) // 4 * 4
    afmt, wfmt = nv.fmt_for(q_bit, "act"), nv.fmt_for(q_bit, "weight")
    d = nv.SlfpConvDesc(N, H, W, C, Cp, K, R, S, stride, stride, pad, pad, dil, dil, groups, afmt)
    Ho = (H + 2 * pad - dil * (R - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (S - 1) - 1) // stride + 1
    st = nv.stream()
    xc = torch.empty((N, H, W, Cp), dtype=torch.uint8, device=dev())
    nv.check(lib.slfp_quantize_nhwc_f32(x.data_ptr(), N * H * W, C, Cp, float(np.float32(ka)), afmt, xc.data_ptr(), st))
    pitch = lib.slfp_conv_wpitch(ctypes.byref(d))
    wc = torch.empty((K * pitch,), dtype=torch.uint8, device=dev())
    wh = torch.empty((K * pitch,), dtype=torch.float16, device=dev())
    so, sc, sr, ss = w.stride()
    nv.check(lib.slfp_prepare_weights(ctypes.byref(d), w.data_ptr(), so, sc, sr, ss, float(np.float32(kw)), wfmt,
                                      wh.data_ptr(), wc.data_ptr(), None, st))
    epi = nv.SlfpEpilogue()
    keep = []
    y32 = torch.full((N, Ho, Wo, K), float("nan"), dtype=torch.float32, device=dev())
    epi.y_f32 = y32.data_ptr()
    epi.post_a, epi.post_b = float(np.float32(ka)), float(np.float32(kw))
    if bias_q is not None:
        b = torch.from_numpy(np.ascontiguousarray(bias_q, np.float32)).to(dev())
        keep.append(b)
        epi.bias_q = b.data_ptr()
    out = {}
    ek = epi_kwargs or {}
    if "ch_scale" in ek:
        s_ = torch.from_numpy(ek["ch_scale"].astype(np.float32)).to(dev())
        h_ = torch.from_numpy(ek["ch_shift"].astype(np.float32)).to(dev())
        keep += [s_, h_]
        epi.ch_scale, epi.ch_shift = s_.data_ptr(), h_.data_ptr()
    if "residual" in ek:
        r = torch.from_numpy(np.ascontiguousarray(ek["residual"])).to(dev()).permute(0, 2, 3, 1).contiguous()
        keep.append(r)
        epi.residual = r.data_ptr()
        epi.residual_f16 = 1 if r.dtype == torch.float16 else 0
    epi.relu = 1 if ek.get("relu") else 0
    y16 = yc = yc2 = None
    if ek.get("want_f16"):
        y16 = torch.zeros((N, Ho, Wo, K), dtype=torch.float16, device=dev())
        epi.y_f16 = y16.data_ptr()
    if "next_k" in ek:
        kp = (K + 15) // 16 * 16
        yc = torch.full((N, Ho, Wo, kp), 0x55, dtype=torch.uint8, device=dev())
        epi.y_codes, epi.next_k_div, epi.next_fmt, epi.k_phys_out = yc.data_ptr(), float(np.float32(ek["next_k"])), afmt, kp
        if "next_k2" in ek:
            yc2 = torch.full((N, Ho, Wo, kp), 0x55, dtype=torch.uint8, device=dev())
            epi.y_codes2, epi.next_k_div2 = yc2.data_ptr(), float(np.float32(ek["next_k2"]))
    nv.check(lib.slfp_conv2d_fwd(ctypes.byref(d), xc.data_ptr(), (wh if dense else wc).data_ptr(), ctypes.byref(epi), st))
    torch.cuda.synchronize()
    out["y"] = y32.permute(0, 3, 1, 2).cpu().numpy()
    out["x_codes"] = xc.cpu().numpy()
    out["w_codes"] = wc.cpu().numpy().reshape(K, pitch)
    out["w_f16"] = wh.cpu().numpy().reshape(K, pitch)
    if y16 is not None:
        out["y_f16"] = y16.permute(0, 3, 1, 2).cpu().numpy()
    if yc is not None:
        out["y_codes"] = yc.cpu().numpy()
    if yc2 is not None:
        out["y_codes2"] = yc2.cpu().numpy()
    out["desc"] = d
    out["dev"] = dict(xc=xc, wc=wc)
    return out


def decode_e4m3(codes):
    """Value of OCP e4m3 bytes [s][e:4][m:3], bias 7 (SLFP_FMT_E4M3: SFP<3,3> values stored as the fp8 number itself)."""
    c = np.asarray(codes).astype(np.int64)
    e, m = (c >> 3) & 15, c & 7
    mag = np.where(e == 0, m * 2.0 ** -9, (1.0 + m / 8.0) * np.exp2(e - 7.0))
    mag = np.where((c & 0x7f) == 0x7f, np.nan, mag)
    return np.where(c & 0x80, -mag, mag).astype(np.float32)


def decode_codes(orc, codes, fmt):
    """Value of activation codes in ANY of the formats the kernels exchange (include/slfp_b200.h)."""
    codes = np.asarray(codes)
    if fmt in (nv.FMT_SLFP34_RELU, nv.FMT_SFP33_RELU):
        return orc.decode_relu(codes, fmt == nv.FMT_SFP33_RELU)
    if fmt == nv.FMT_SFP33_SFAST:
        m = orc.decode_relu(codes & 0x7f, True)
        return np.where(codes & 0x80, -m, m).astype(np.float32)
    if fmt == nv.FMT_E4M3:
        return decode_e4m3(codes)
    return orc.decode(codes, fmt)


def decode_q16(orc, halves, qfmt):
    """SLFP_FMT_F16Q tensor (float16 images of the codes of format qfmt) -> the exact grid values: every half must BE the
    float16 image of a code's value (asserted), and is mapped back to that value."""
    table = decode_codes(orc, np.arange(256, dtype=np.uint8), qfmt).astype(np.float32)
    table = table[np.isfinite(table)]
    img = table.astype(np.float16)
    order = np.argsort(img, kind="stable")
    img_s, tab_s = img[order], table[order]
    h = np.asarray(halves)
    pos = np.clip(np.searchsorted(img_s, h), 0, len(img_s) - 1)
    assert (img_s[pos] == h).all(), "a float16 activation that is not the image of any code"
    return tab_s[pos]


def decode_tensor(orc, t):
    """Values of an engine tensor of kind 'codes' or 'q16'."""
    arr = t.buf.cpu().numpy()
    if getattr(t, "pad", None) is not None:                 # physically zero-padded buffer: the interior is the tensor
        top, left, _, _ = t.pad
        assert (np.delete(arr, np.s_[top:top + t.h], axis=1) == 0).all() and (np.delete(arr, np.s_[left:left + t.w], axis=2) == 0).all()
        arr = arr[:, top:top + t.h, left:left + t.w]
    if getattr(t, "im2col", False):                         # im2col matrix of a 3x3 / pad 1 RGB stem: entry (r * 3 + s) * 4 + c
        v = decode_q16(orc, arr, t.qfmt)
        n, h, w, _ = v.shape
        center = v[..., 16:19]
        for r in range(3):                                  # every tap is the shifted centre tap, zero outside the image
            for s_ in range(3):
                want = np.zeros_like(center)
                ys, xs = slice(max(0, 1 - r), min(h, h + 1 - r)), slice(max(0, 1 - s_), min(w, w + 1 - s_))
                yd, xd = slice(max(0, r - 1), min(h, h + r - 1)), slice(max(0, s_ - 1), min(w, w + s_ - 1))
                want[:, ys, xs] = center[:, yd, xd]
                e = (r * 3 + s_) * 4
                assert (v[..., e:e + 3] == want).all() and (v[..., e + 3] == 0).all()
        assert (v[..., 36:] == 0).all()
        return center
    return decode_q16(orc, arr, t.qfmt) if t.fmt == nv.FMT_F16Q else decode_codes(orc, arr, t.fmt)
