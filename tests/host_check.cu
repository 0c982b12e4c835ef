// host_check.cu -- compiles the PRODUCT's encode/decode device functions (slfp_common.cuh) for the
// host, so the CPU test-suite can sweep them against the oracle without a GPU.  Test-only.
#include "../cnns_slfp_quantization_b200/csrc/slfp_common.cuh"

using namespace slfp;

extern "C" void hostcheck_quantize(const float* x, size_t n, float k_div, int fmt, int zero_is_zero,
                                   uint8_t* codes, float* fakeq) {
    for (size_t i = 0; i < n; ++i) {
        const float v = div_rn(x[i], k_div);
        uint32_t c = 0;
        float q;
        switch (fmt) {
            case SLFP_FMT_SFP33: c = encode<SLFP_FMT_SFP33>(v); q = decode<true>(c, h_pow2frac); break;
            case SLFP_FMT_SLFP34_ACT: c = encode<SLFP_FMT_SLFP34_ACT>(v); q = decode<false>(c, h_pow2frac); break;
            case SLFP_FMT_SLFP34_WGT: c = encode<SLFP_FMT_SLFP34_WGT>(v); q = decode<false>(c, h_pow2frac); break;
            default: q = layerout_quantize(v, zero_is_zero != 0); break;
        }
        if (codes) codes[i] = (uint8_t)c;
        if (fakeq) fakeq[i] = q;
    }
}

extern "C" void hostcheck_decode(const uint8_t* codes, size_t n, int fmt, float* out) {
    for (size_t i = 0; i < n; ++i)
        out[i] = fmt == SLFP_FMT_SFP33 ? decode<true>(codes[i], h_pow2frac) : decode<false>(codes[i], h_pow2frac);
}

// div_k (reciprocal + two FMAs) against the IEEE quotient; returns the number of mismatching elements.
extern "C" size_t hostcheck_divk_mismatches(const float* x, size_t n, float k) {
    const DivK d = make_divk(k);
    size_t bad = 0;
    for (size_t i = 0; i < n; ++i) {
        const float a = div_k(x[i], d), b = x[i] / k;
        uint32_t ua, ub;
        memcpy(&ua, &a, 4); memcpy(&ub, &b, 4);
        bad += (ua != ub) && !(a != a && b != b);
    }
    return bad;
}

// encode_relu(q) must equal encode(q) for every q >= +0 that is not NaN; returns the mismatch count.
extern "C" size_t hostcheck_encode_relu_mismatches(const float* q, size_t n, int fmt) {
    size_t bad = 0;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t a = fmt == SLFP_FMT_SFP33 ? encode_relu<SLFP_FMT_SFP33>(q[i]) : encode_relu<SLFP_FMT_SLFP34_ACT>(q[i]);
        const uint32_t b = fmt == SLFP_FMT_SFP33 ? encode<SLFP_FMT_SFP33>(q[i]) : encode<SLFP_FMT_SLFP34_ACT>(q[i]);
        bad += a != b;
    }
    return bad;
}
// encode_q(div_k_fused(x, K), x) against encode(x / K)
extern "C" size_t hostcheck_fused_quant_mismatches(const float* x, size_t n, float k, int fmt) {
    const DivK d = make_divk(k);
    size_t bad = 0;
    for (size_t i = 0; i < n; ++i) {
        const float q = div_k_fused(x[i], d);
        const uint32_t a = fmt == SLFP_FMT_SFP33 ? encode_q<SLFP_FMT_SFP33>(q, x[i]) : encode_q<SLFP_FMT_SLFP34_ACT>(q, x[i]);
        const uint32_t b = fmt == SLFP_FMT_SFP33 ? encode<SLFP_FMT_SFP33>(x[i] / k) : encode<SLFP_FMT_SLFP34_ACT>(x[i] / k);
        bad += a != b;
    }
    return bad;
}

// post-ReLU code formats: codes and their values through the product's encode_relu_fast / decode_relu
extern "C" void hostcheck_relu_codes(const float* q, size_t n, int sfp33, uint8_t* codes, float* values, uint8_t* codes16) {
    for (size_t i = 0; i < n; ++i) {
        const uint32_t c = sfp33 ? encode_relu_fast<true>(q[i]) : encode_relu_fast<false>(q[i]);
        codes[i] = (uint8_t)c;
        values[i] = sfp33 ? decode_relu<true>(c, h_pow2frac) : decode_relu<false>(c, h_pow2frac);
        // the epilogue's form: clamp(q/16, 0, 1) -> raw16 -> saturating pack
        float q16 = q[i] * 0.0625f;
        q16 = q16 > 0.f ? (q16 > 1.f ? 1.f : q16) : 0.f;
        int32_t t = sfp33 ? encode_relu_fast_raw16<true>(q16) : encode_relu_fast_raw16<false>(q16);
        codes16[i] = (uint8_t)(t < 0 ? 0 : (t > 255 ? 255 : t));
    }
}

// encode_inrange<FMT>(v) against encode<FMT>(v) for |v| <= limit; returns the mismatch count
extern "C" size_t hostcheck_encode_inrange_mismatches(const float* v, size_t n, int fmt) {
    size_t bad = 0;
    for (size_t i = 0; i < n; ++i) {
        uint32_t a, b;
        switch (fmt) {
            case SLFP_FMT_SFP33: a = encode_inrange<SLFP_FMT_SFP33>(v[i]); b = encode<SLFP_FMT_SFP33>(v[i]); break;
            case SLFP_FMT_SLFP34_ACT: a = encode_inrange<SLFP_FMT_SLFP34_ACT>(v[i]); b = encode<SLFP_FMT_SLFP34_ACT>(v[i]); break;
            default: a = encode_inrange<SLFP_FMT_SLFP34_WGT>(v[i]); b = encode<SLFP_FMT_SLFP34_WGT>(v[i]); break;
        }
        bad += a != b;
    }
    return bad;
}

// encode_balanced<FMT>(v) against encode<FMT>(v); returns the mismatch count
extern "C" size_t hostcheck_encode_balanced_mismatches(const float* v, size_t n, int fmt) {
    size_t bad = 0;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t a = fmt == SLFP_FMT_SFP33 ? encode_balanced<SLFP_FMT_SFP33>(v[i]) : encode_balanced<SLFP_FMT_SLFP34_ACT>(v[i]);
        const uint32_t b = fmt == SLFP_FMT_SFP33 ? encode<SLFP_FMT_SFP33>(v[i]) : encode<SLFP_FMT_SLFP34_ACT>(v[i]);
        bad += a != b;
    }
    return bad;
}

// the table encoder of the stand-alone quantizer (enc_lut_index + enc_lut_entry + sign) against encode<FMT>(q); the
// dividend only tells zero from non-zero on that path.  Returns the mismatch count.
extern "C" size_t hostcheck_encode_lut_mismatches(const float* q, size_t n, int fmt) {
    static uint8_t tab[2][kEncLutBytes];
    static bool init = false;
    if (!init) {
        for (int i = 0; i < kEncLutBytes; ++i) {
            tab[0][i] = (uint8_t)enc_lut_entry<SLFP_FMT_SFP33>((uint32_t)i);
            tab[1][i] = (uint8_t)enc_lut_entry<SLFP_FMT_SLFP34_ACT>((uint32_t)i);
        }
        init = true;
    }
    size_t bad = 0;
    for (size_t i = 0; i < n; ++i) {
        const float x = q[i] == 0.0f ? 0.0f : 1.0f;
        const uint32_t idx = fmt == SLFP_FMT_SFP33 ? enc_lut_index<SLFP_FMT_SFP33>(q[i], x) : enc_lut_index<SLFP_FMT_SLFP34_ACT>(q[i], x);
        if (idx >= (uint32_t)kEncLutBytes) { ++bad; continue; }
        const uint32_t a = (uint32_t)tab[fmt == SLFP_FMT_SFP33 ? 0 : 1][idx] | ((f2u(q[i]) >> 24) & 0x80u);
        const uint32_t b = fmt == SLFP_FMT_SFP33 ? encode<SLFP_FMT_SFP33>(q[i]) : encode<SLFP_FMT_SLFP34_ACT>(q[i]);
        bad += a != b;
    }
    return bad;
}

// encode_wgt_bucket (threshold search by bucket table) against encode<SLFP_FMT_SLFP34_WGT>; returns the mismatch count
extern "C" size_t hostcheck_encode_wgt_bucket_mismatches(const float* v, size_t n) {
    static uint2 tbl[32];
    static bool init = false;
    if (!init) {
        for (uint32_t b = 0; b < 32; ++b) wgt_bucket_entry(b, tbl[b].x, tbl[b].y);
        init = true;
    }
    size_t bad = 0;
    for (size_t i = 0; i < n; ++i) bad += encode_wgt_bucket(v[i], tbl) != encode<SLFP_FMT_SLFP34_WGT>(v[i]);
    return bad;
}

// layerout_relu (fast epilogues: three FMA-pipe operations) against relu(quantize_layerout(y)) with exact 0 -> 0;
// returns the mismatch count
extern "C" size_t hostcheck_layerout_relu_mismatches(const float* y, size_t n) {
    size_t bad = 0;
    for (size_t i = 0; i < n; ++i) {
        const float a = layerout_relu(y[i]);
        float b = layerout_quantize(y[i], true);
        b = b > 0.0f ? b : 0.0f;
        bad += f2u(a) != f2u(b);
    }
    return bad;
}

// signed fast SFP<3,3> codes (SLFP_FMT_SFP33_SFAST; the depthwise kernels' encoder) and their decoded values
extern "C" void hostcheck_sfast_codes(const float* q, size_t n, uint8_t* codes, float* values) {
    for (size_t i = 0; i < n; ++i) {
        float a16 = fabsf(q[i]) * 0.0625f;
        a16 = a16 > 1.f ? 1.f : a16;
        int32_t m = encode_relu_fast_raw16<true>(a16);
        m = m < 0 ? 0 : (m > 127 ? 127 : m);
        const uint32_t c = (uint32_t)m | ((f2u(q[i]) >> 24) & 0x80u);
        codes[i] = (uint8_t)c;
        values[i] = decode_act_any(c, SLFP_FMT_SFP33_SFAST, h_pow2frac);
    }
}

// encode_wgt_lut (the one-look-up weight encoder of wprep_rows_kernel) against encode<SLFP34_WGT> on its domain: finite
// non-zero values.  Also checks the float16 image: float16(decode(code)).
extern "C" size_t hostcheck_encode_wgt_lut_mismatches(const float* v, size_t n) {
    static WgtLutEntry lut[kWgtLutEntries];
    static bool built = false;
    if (!built) {
        for (uint32_t i = 0; i < (uint32_t)kWgtLutEntries; ++i) lut[i] = wgt_lut_entry(i, h_pow2frac);
        built = true;
    }
    size_t bad = 0;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t qb = f2u(v[i]), qa = qb & 0x7fffffffu;
        if (qa == 0u || qa >= 0x7f800000u) continue;                       // outside the look-up's domain
        uint32_t h16;
        const uint32_t code = encode_wgt_lut(qb, lut[wgt_lut_index(qa)], h16);
        const uint32_t want = encode<SLFP_FMT_SLFP34_WGT>(v[i]);
        const uint32_t hw = f16_bits_rn(decode<false>(want, h_pow2frac));
        bad += (code != want) || (h16 != hw);
    }
    return bad;
}

