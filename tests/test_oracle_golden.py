"""CPU: the oracle (oracle/slfp_oracle.c) against the fixtures generated from the REFERENCE
(tests/golden/make_golden.py).  This is what pins the oracle; the exhaustive sweep done at fixture
generation time is logged in tests/golden/make_golden.log."""
import numpy as np
import pytest

from conftest import bits, same_bits

FMTS = {"sfp33_act": 0, "sfp33_wgt": 0, "slfp34_act": 1, "slfp34_wgt": 2, "sfp44_out": 3}


def test_make_golden_log_records_clean_sweep():
    import os
    log = open(os.path.join(os.path.dirname(__file__), "golden", "make_golden.log")).read()
    assert log.count("mismatches = 0") == 5 and "codes round-trip ok" in log


def test_reference_kat(orc, g_quant):
    # utils/sfp_quant.py:177-182, the only known-answer material in the reference
    _, fq = orc.quantize(g_quant["kat_x"], orc.FMT_SLFP34_ACT)
    assert same_bits(g_quant["kat_slfp34_act"], fq).all()
    np.testing.assert_allclose(fq, [1e-10, 0.125, 0.125, 0.125, 0.20131129, 1.0, 15.32165241], rtol=1e-7)


@pytest.mark.parametrize("name", list(FMTS))
def test_quantizers_bit_exact(orc, g_quant, name):
    codes, fq = orc.quantize(g_quant["x"], FMTS[name])
    assert same_bits(g_quant[name], fq).all()
    if codes is not None:
        assert same_bits(fq, orc.decode(codes, FMTS[name])).all()


@pytest.mark.parametrize("name", ["slfp34_act", "slfp34_wgt", "sfp33_act"])
def test_prescaled_quantizers(orc, g_quant, name):
    for i, k in enumerate(g_quant["prescale_k"]):
        _, fq = orc.quantize(g_quant["prescale_x"], FMTS[name], kdiv=k)
        assert same_bits(g_quant["scaled_" + name][i], fq).all()


def test_tables_match_reference(orc, g_tables):
    one_to_two = g_tables["wgt_thresh_bits"].view(np.float32)
    # each threshold maps to its log-code, the float just below maps to the previous one
    c_hi, _ = orc.quantize(one_to_two, orc.FMT_SLFP34_WGT)
    c_lo, _ = orc.quantize((g_tables["wgt_thresh_bits"] - 1).view(np.float32), orc.FMT_SLFP34_WGT)
    assert (c_hi.astype(int) - 64 == np.arange(1, 17)).all()
    assert (c_lo.astype(int) - 64 == np.arange(0, 16)).all()
    dec = orc.decode(np.arange(64, 80, dtype=np.uint8), orc.FMT_SLFP34_ACT)
    assert (bits(dec) == g_tables["pow2frac_bits"]).all()
    assert g_tables["act_logcode"].tolist() == [0, 1, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16]
    c, _ = orc.quantize(g_tables["act_first_mantissa_bits"].view(np.float32), orc.FMT_SLFP34_ACT)
    assert (c.astype(int) - 64 == g_tables["act_logcode"]).all()
    assert bits(np.array([1e-10, 15.32165, 0.0625, 0.125, 15.0, 248.0], np.float32)).tolist() == \
        g_tables["const_bits"].tolist()


def _case(g, name):
    cfg = g[name + ".cfg"]
    return cfg, {k[len(name) + 1:]: g[k] for k in g.files if k.startswith(name + ".")}


def test_conv_cases_forward_and_backward(orc, g_conv):
    for name in g_conv["names"]:
        name = str(name)
        if name.startswith("fc_"):
            continue
        cfg, d = _case(g_conv, name)
        qbit, N, C, H, W, O, k, st, pad, dil, groups, has_bias = [int(v) for v in cfg]
        ka, kw = d["k"]
        xq, wq, y = orc.conv2d_Q_forward(d["x"], d["w"], d.get("b"), ka, kw, qbit, st, pad, dil, groups)
        if qbit != 32:
            assert same_bits(d["input_q"], xq).all(), name
            assert same_bits(d["weight_q"], wq).all(), name
        # float32 conv: summation order differs between oneDNN and the oracle's double accumulation
        l1 = orc.conv2d_q(np.abs(xq), np.abs(wq), None, st, pad, dil, groups, ka, kw)
        assert (np.abs(y - d["y"]) <= 4e-6 * l1 + 1e-6).all(), name
        dx, dw, db = orc.conv2d_q_bwd(xq, wq, d["gy"], st, pad, dil, groups, ka, kw, with_bias=bool(has_bias))
        np.testing.assert_allclose(dx, d["dx"], rtol=2e-4, atol=2e-4 * np.abs(d["dx"]).max(), err_msg=name)
        np.testing.assert_allclose(dw, d["dw"], rtol=2e-4, atol=2e-4 * np.abs(d["dw"]).max(), err_msg=name)
        if has_bias:
            np.testing.assert_allclose(db, d["db"], rtol=2e-4, atol=1e-4, err_msg=name)


def test_linear_cases(orc, g_conv):
    for name in ("fc_8", "fc_7"):
        cfg, d = _case(g_conv, name)
        qbit = int(cfg[0])
        ka, kw = d["k"]
        xq, wq, y = orc.linear_Q_forward(d["x"], d["w"], d["b"], ka, kw, qbit)
        assert same_bits(d["input_q"], xq).all() and same_bits(d["weight_q"], wq).all()
        np.testing.assert_allclose(y, d["y"], rtol=1e-5, atol=1e-5)


def test_activation_cases(orc, g_act):
    for kind, name in ((orc.ACT_STL, "stl"), (orc.ACT_SWISH, "swish"), (orc.ACT_SIGMOID, "sigmoid")):
        np.testing.assert_allclose(orc.act_fwd(g_act["x"], kind), g_act[name + ".y"], rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(orc.act_bwd(g_act["x"], g_act["gy"], kind), g_act[name + ".gx"], rtol=1e-5, atol=1e-7)


def test_sgd_cases(orc, g_sgd):
    for name in g_sgd["names"]:
        name = str(name)
        lr, mom, damp, wd, nest, qbit = g_sgd[name + ".hp"]
        cls = str(g_sgd[name + ".cls"])
        mode = {"NormalSGD": orc.SGD_NORMAL, "DSGD": orc.SGD_DSGD, "SSGD": orc.SGD_SSGD}[cls]
        p = g_sgd[name + ".p0"].copy()
        buf = np.zeros_like(p)
        for step, (gr, want) in enumerate(zip(g_sgd[name + ".grads"], g_sgd[name + ".ps"])):
            g = gr.copy()
            orc.sgd_step(p, g, buf, mode, 32 if qbit < 0 else int(qbit), lr, mom, damp, wd, bool(nest), step == 0)
            # bit-exact: ATen's add(alpha) is a fused multiply-add, which the oracle spells as fmaf
            assert (bits(p) == bits(want)).all(), (name, step)
