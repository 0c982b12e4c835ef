#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the REFERENCE itself.

Runs only in the build container (needs /root/reference, which is read-only and does not
exist on the GPU box).  It imports the unmodified reference modules (utils/sfp_quant.py,
utils/conv2d_func.py, utils/activation_func.py, utils/optimizer.py), feeds them seeded
synthetic inputs and stores input/output pairs as small .npz files.  It also performs the
exhaustive sweep that pins the C oracle: every float32 mantissa (2^23) at 11 exponents through
every quantizer, reference bits == oracle bits.  The log of this script is committed as
make_golden.log.

    python tests/golden/make_golden.py            # everything
    python tests/golden/make_golden.py --no-sweep # fixtures only
"""
import argparse
import os
import sys
import time
import types

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.modules.setdefault("torchsummary", types.SimpleNamespace(summary=lambda *a, **k: None))

import numpy as np
import torch

from utils import sfp_quant as ref_q          # noqa: E402  (reference)
from utils import conv2d_func as ref_c        # noqa: E402
from utils import activation_func as ref_a    # noqa: E402
from utils import optimizer as ref_o          # noqa: E402
from oracle import slfp_oracle as orc         # noqa: E402

torch.set_num_threads(8)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def same_bits(ref, got):
    """Bit equality, except that any NaN equals any NaN."""
    ref, got = np.asarray(ref, np.float32), np.asarray(got, np.float32)
    eq = bits(ref) == bits(got)
    eq |= np.isnan(ref) & np.isnan(got)
    return eq


REF_FNS = {
    "sfp33_act": (lambda: ref_q.quantize_act(7), orc.FMT_SFP33),
    "sfp33_wgt": (lambda: ref_q.quantize_weight(7), orc.FMT_SFP33),
    "slfp34_act": (lambda: ref_q.quantize_act(8), orc.FMT_SLFP34_ACT),
    "slfp34_wgt": (lambda: ref_q.quantize_weight(8), orc.FMT_SLFP34_WGT),
    "sfp44_out": (lambda: ref_q.quantize_layerout(8), orc.FMT_SFP44_OUT),
}


def run_ref(name, x):
    with torch.no_grad():
        return REF_FNS[name][0]()(torch.from_numpy(x)).numpy()


def sweep():
    mant = np.arange(1 << 23, dtype=np.uint32)
    for name, (_, fmt) in REF_FNS.items():
        t0 = time.time()
        exps = range(-6, 5) if fmt != orc.FMT_SFP44_OUT else list(range(-9, 9)) + [-126, -127, -100, 60]
        bad = 0
        for e in exps:
            if e == -127:          # the denormal range: exponent field 0
                x = mant.copy().view(np.float32)
            else:
                x = (mant | np.uint32((e + 127) << 23)).view(np.float32)
            for sgn in (1.0, -1.0) if e in (0, -4, 3) else (1.0,):
                xs = (x * np.float32(sgn)).astype(np.float32)
                ref = run_ref(name, xs)
                _, got = orc.quantize(xs, fmt, 1.0, bugcompat=True, want_codes=False)
                bad += int((~same_bits(ref, got)).sum())
        print(f"sweep {name:11s}: exponents {list(exps)[0]}..{list(exps)[-1]} x 2^23 mantissas, "
              f"mismatches = {bad}  ({time.time() - t0:.1f}s)", flush=True)
        assert bad == 0, name
    # codes round-trip: decode(code) == fake-quant for every 8-bit format
    for fmt in (orc.FMT_SFP33, orc.FMT_SLFP34_ACT, orc.FMT_SLFP34_WGT):
        x = (np.random.default_rng(1).standard_normal(1 << 20) * 6).astype(np.float32)
        c, fq = orc.quantize(x, fmt)
        assert same_bits(fq, orc.decode(c, fmt)).all()
    print("codes round-trip ok")


def derive_tables():
    """Read the decode / threshold tables off the reference and store them as a fixture."""
    mant = np.arange(1 << 23, dtype=np.uint32)
    x = (mant | np.uint32(127 << 23)).view(np.float32)          # [1, 2)
    out = {}
    w = bits(run_ref("slfp34_wgt", x))
    vals = np.unique(w)
    assert len(vals) == 17
    out["pow2frac_bits"] = vals[:16].astype(np.uint32)
    out["wgt_thresh_bits"] = np.array([bits(x)[np.argmax(w == v)] for v in vals[1:]], dtype=np.uint32)
    a = bits(run_ref("slfp34_act", x))
    idx = np.searchsorted(vals, a)
    # first mantissa mapping to each distinct act output, and which log-code it is
    firsts = np.flatnonzero(np.diff(idx, prepend=-1) != 0)
    out["act_first_mantissa_bits"] = bits(x)[firsts]
    out["act_logcode"] = idx[firsts].astype(np.uint8)
    s = bits(run_ref("sfp33_act", x))
    sv = np.unique(s)
    out["sfp33_values_bits"] = sv.astype(np.uint32)
    out["sfp33_first_mantissa_bits"] = np.array([bits(x)[np.argmax(s == v)] for v in sv], dtype=np.uint32)
    consts = np.array([1e-10, 15.32165, 0.0625, 0.125, 15.0, 248.0], dtype=np.float32)
    out["const_bits"] = bits(consts)
    np.savez(os.path.join(HERE, "quant_tables.npz"), **out)
    print("pow2frac:", " ".join(f"{v:08x}" for v in out["pow2frac_bits"]))
    print("wgt thresholds:", " ".join(f"{v:08x}" for v in out["wgt_thresh_bits"]))
    print("act log-codes used:", out["act_logcode"].tolist())


def edge_values():
    e = [0.0, -0.0, 1e-30, -1e-30, 1e-40, -3e-39, 0.0624999, 0.0625, 0.06251, 0.1, 0.1249999, 0.125,
         0.2, 1.0, -1.0, 1.03, 1.0625, 1.09375, 14.74, 14.75, 14.99, 14.9999, 15.0, 15.32165, 15.3216505,
         15.3217, 15.9, 16.0, 100.0, 247.9, 248.0, 300.0, 3e38, float("inf"), -float("inf"), float("nan"),
         0.01, 0.5, 2.0, 4.0, 8.0, 0.25, 7.999999, 3.9999998, 1.9999999, 0.99999994, 0.12499999]
    e = np.array(e, dtype=np.float32)
    return np.concatenate([e, -e])


def quant_samples():
    rng = np.random.default_rng(20261018)
    x = np.concatenate([
        edge_values(),
        (rng.standard_normal(6000) * 4).astype(np.float32),
        np.exp(rng.uniform(np.log(1e-3), np.log(40), 3000)).astype(np.float32) * rng.choice([-1, 1], 3000).astype(np.float32),
        rng.uniform(-16, 16, 3000).astype(np.float32),
    ]).astype(np.float32)
    out = {"x": x}
    for name in REF_FNS:
        out[name] = run_ref(name, x)
    # KAT printed by the reference's own __main__ (utils/sfp_quant.py:177-182)
    kat = np.array([0.01, 0.06251, 0.125, 0.1, 0.2, 1, 15], dtype=np.float32)
    out["kat_x"] = kat
    out["kat_slfp34_act"] = run_ref("slfp34_act", kat)
    # pre-scale: tensor / 0-dim float64 tensor (conv2d_func.py:17-22)
    ks = np.array([2.640000104904175 / 15.5, 0.7817208766937256 / 15.5, 0.33, 1.0, 3.0], dtype=np.float64)
    xs = (rng.standard_normal(4096) * 3).astype(np.float32)
    out["prescale_x"] = xs
    out["prescale_k"] = ks
    out["prescale_y"] = np.stack([(torch.from_numpy(xs) / torch.tensor(float(k))).numpy() for k in ks])
    assert all((out["prescale_y"][i] == xs / np.float32(k)).all() for i, k in enumerate(ks)), \
        "pre-scale is not an IEEE float32 division by float32(K)"
    # quantize with pre-scale, as the modules do
    for name in ("slfp34_act", "slfp34_wgt", "sfp33_act"):
        out["scaled_" + name] = np.stack(
            [REF_FNS[name][0]()(torch.from_numpy(xs) / torch.tensor(float(k))).numpy() for k in ks])
    np.savez_compressed(os.path.join(HERE, "quant_samples.npz"), **out)
    print("quant_samples:", x.size, "inputs")


CONV_CASES = [
    # name, qbit, factory, N, C, H, W, O, k, stride, pad, dil, groups, bias
    ("c3x3_s1", 8, "conv2d_Q", 2, 16, 9, 11, 24, 3, 1, 1, 1, 1, False),
    ("c1x1_s1", 8, "conv2d_Q", 3, 64, 7, 7, 32, 1, 1, 0, 1, 1, False),
    ("c3x3_s2", 8, "conv2d_Q", 2, 32, 10, 10, 16, 3, 2, 1, 1, 1, False),
    ("c1x1_s2", 8, "conv2d_Q", 2, 32, 8, 8, 48, 1, 2, 0, 1, 1, False),
    ("c7x7_stem", 8, "conv2d_Q", 2, 3, 20, 20, 16, 7, 2, 3, 1, 1, False),
    ("c3x3_bias", 8, "conv2d_Q_bias", 2, 16, 8, 8, 32, 3, 1, 1, 1, 1, True),
    ("c3x3_dil2", 8, "conv2d_Q", 1, 16, 12, 12, 16, 3, 1, 2, 2, 1, False),
    ("dw3x3_s1", 7, "conv2d_Q", 2, 32, 9, 9, 32, 3, 1, 1, 1, 32, False),
    ("dw3x3_s2", 7, "conv2d_Q", 2, 24, 10, 10, 24, 3, 2, 1, 1, 24, False),
    ("grp4_3x3", 8, "conv2d_Q", 2, 32, 6, 6, 64, 3, 1, 1, 1, 4, False),
    ("c1x1_sfp7", 7, "conv2d_Q", 2, 58, 6, 6, 58, 1, 1, 0, 1, 1, False),
    ("c5x5_bias", 8, "conv2d_Q_bias", 1, 16, 9, 9, 16, 5, 1, 2, 1, 1, True),
    ("c3x3_fp32", 32, "conv2d_Q", 1, 8, 6, 6, 8, 3, 1, 1, 1, 1, False),
]


def conv_cases():
    out = {}
    names = []
    for (name, qbit, fac, N, C, H, W, O, k, st, pad, dil, groups, bias) in CONV_CASES:
        g = torch.Generator().manual_seed(hash(name) % (1 << 31) if False else sum(map(ord, name)))
        x = torch.randn(N, C, H, W, generator=g) * 1.7
        w = torch.randn(O, C // groups, k, k, generator=g) * 0.2
        b = torch.randn(O, generator=g) * 0.5 if bias else None
        ka = float(x.abs().max()) / 15.5 * 1.1       # a little saturation on purpose
        kw = float(w.abs().max()) / 15.5 * 1.05
        cls = getattr(ref_c, fac)(qbit, kw, ka)
        m = cls(C, O, k, stride=st, padding=pad, dilation=dil, groups=groups)
        with torch.no_grad():
            m.weight.copy_(w)
            if bias:
                m.bias.copy_(b)
        xin = x.clone().requires_grad_(True)
        y = m(xin)
        gy = torch.randn(y.shape, generator=g)
        y.backward(gy)
        out[name + ".x"] = x.numpy(); out[name + ".w"] = w.numpy()
        if bias:
            out[name + ".b"] = b.numpy(); out[name + ".db"] = m.bias.grad.numpy()
        out[name + ".cfg"] = np.array([qbit, N, C, H, W, O, k, st, pad, dil, groups, int(bias)], dtype=np.int64)
        out[name + ".k"] = np.array([ka, kw], dtype=np.float64)
        out[name + ".input_q"] = m.input_q.detach().numpy()
        out[name + ".weight_q"] = m.weight_q.detach().numpy()
        out[name + ".y"] = y.detach().numpy()
        out[name + ".gy"] = gy.numpy()
        out[name + ".dx"] = xin.grad.numpy()
        out[name + ".dw"] = m.weight.grad.numpy()
        names.append(name)
    # Linear_Q
    for name, qbit, B, I, O in [("fc_8", 8, 5, 64, 10), ("fc_7", 7, 4, 48, 12)]:
        g = torch.Generator().manual_seed(sum(map(ord, name)))
        x = torch.randn(B, I, generator=g) * 2
        w = torch.randn(O, I, generator=g) * 0.1
        b = torch.randn(O, generator=g) * 0.3
        ka = float(x.abs().max()) / 15.5; kw = float(w.abs().max()) / 15.5
        m = ref_c.linear_Q(qbit, kw, ka)(I, O)
        with torch.no_grad():
            m.weight.copy_(w); m.bias.copy_(b)
        xin = x.clone().requires_grad_(True)
        y = m(xin)
        gy = torch.randn(y.shape, generator=g)
        y.backward(gy)
        out[name + ".x"] = x.numpy(); out[name + ".w"] = w.numpy(); out[name + ".b"] = b.numpy()
        out[name + ".cfg"] = np.array([qbit, B, I, O], dtype=np.int64)
        out[name + ".k"] = np.array([ka, kw], dtype=np.float64)
        out[name + ".input_q"] = m.input_q.detach().numpy(); out[name + ".weight_q"] = m.weight_q.detach().numpy()
        out[name + ".y"] = y.detach().numpy(); out[name + ".gy"] = gy.numpy()
        out[name + ".dx"] = xin.grad.numpy(); out[name + ".dw"] = m.weight.grad.numpy()
        out[name + ".db"] = m.bias.grad.numpy()
        names.append(name)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "conv_cases.npz"), **out)
    print("conv_cases:", len(names))


def act_cases():
    g = torch.Generator().manual_seed(7)
    x = torch.cat([torch.randn(2000, generator=g) * 3, torch.tensor([0.0, 1.0, -1.0, 1.0000001, -20.0, 20.0, 88.0, -88.0])])
    gy = torch.cat([torch.randn(2000, generator=g) * 2, torch.tensor([0.0, 1.0, -1.0, 1.5, -3.0, 0.5, 1e-3, 7.0])])
    out = {"x": x.numpy(), "gy": gy.numpy()}
    for name, mod in [("stl", ref_a.STL()), ("swish", ref_a.Swish()), ("sigmoid", ref_a.Sigmoid())]:
        xin = x.clone().requires_grad_(True)
        y = mod(xin)
        y.backward(gy)
        out[name + ".y"] = y.detach().numpy()
        out[name + ".gx"] = xin.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "act_cases.npz"), **out)
    print("act_cases ok")


def sgd_cases():
    out = {}
    names = []
    cfgs = [("dsgd8", "DSGD", 8, dict(lr=0.05, momentum=0.9, weight_decay=5e-4)),
            ("dsgd7", "DSGD", 7, dict(lr=0.02, momentum=0.9, weight_decay=0.0)),
            ("dsgd8_nomom", "DSGD", 8, dict(lr=0.1, momentum=0.0, weight_decay=1e-3)),
            ("dsgd8_nest", "DSGD", 8, dict(lr=0.05, momentum=0.8, weight_decay=5e-4, nesterov=True)),
            ("ssgd8", "SSGD", 8, dict(lr=0.03, momentum=0.9, weight_decay=5e-4)),
            ("dsgd32", "DSGD", 32, dict(lr=0.05, momentum=0.9, weight_decay=5e-4)),
            ("normal", "NormalSGD", None, dict(lr=0.05, momentum=0.9, weight_decay=5e-4, dampening=0.1))]
    import warnings
    warnings.simplefilter("ignore")
    for name, cls, qbit, kw in cfgs:
        g = torch.Generator().manual_seed(sum(map(ord, name)))
        p0 = torch.randn(3000, generator=g) * 1.5
        p0[:8] = torch.tensor([0.0, 0.06, 0.0625, 0.124, 15.3, -15.4, 1.0, -1.0])
        p = torch.nn.Parameter(p0.clone())
        opt = getattr(ref_o, cls)([p], qbit, **kw) if qbit is not None else getattr(ref_o, cls)([p], **kw)
        grads, ps = [], []
        for step in range(3):
            gr = torch.randn(3000, generator=g) * (0.5 if step else 0.05)
            p.grad = gr.clone()
            grads.append(gr.numpy().copy())
            opt.step()
            ps.append(p.detach().numpy().copy())
        out[name + ".p0"] = p0.numpy(); out[name + ".grads"] = np.stack(grads); out[name + ".ps"] = np.stack(ps)
        out[name + ".hp"] = np.array([kw.get("lr"), kw.get("momentum", 0), kw.get("dampening", 0),
                                      kw.get("weight_decay", 0), float(kw.get("nesterov", False)),
                                      -1 if qbit is None else qbit], dtype=np.float64)
        out[name + ".cls"] = np.array(cls)
        names.append(name)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "sgd_cases.npz"), **out)
    print("sgd_cases:", len(names))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-sweep", action="store_true")
    args = ap.parse_args()
    print("torch", torch.__version__, "numpy", np.__version__, "threads", torch.get_num_threads())
    derive_tables()
    quant_samples()
    conv_cases()
    act_cases()
    sgd_cases()
    if not args.no_sweep:
        sweep()
    print("done")
