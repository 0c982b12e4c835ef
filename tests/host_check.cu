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
