// slfp_common.cuh -- shared device code of libslfp_b200: the SLFP / SFP number formats.
//
// encode<FMT>(v): float32 -> 8-bit storage code, bit-exact with the reference quantizers
//   utils/sfp_quant.py:14-30,63-78 (SFP<3,3>), :80-96 (SLFP<3,4> activations),
//   :32-47 (SLFP<3,4> weights).  The reference rounds with float32 log2/pow; here the same
//   decision is taken in integer arithmetic on the IEEE-754 bit pattern (its log2f/powf results
//   were characterised exhaustively, see tests/golden/make_golden.log), so the device libm never
//   enters the result.
// decode: code -> the exact float32 the reference's fake-quant tensor holds.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <string.h>
#include <math.h>

#include "../../include/slfp_b200.h"

namespace slfp {

constexpr uint32_t kBitsTiny = 0x2edbe6ffu;  // float32(1e-10)      sfp_quant.py:26,43,74,92
constexpr uint32_t kBitsSat8 = 0x4175257au;  // float32(15.32165)   sfp_quant.py:46,95
constexpr uint32_t kBits0625 = 0x3d800000u;
constexpr uint32_t kBits0125 = 0x3e000000u;
constexpr uint32_t kBits15 = 0x41700000u;
constexpr uint32_t kBits248 = 0x43780000u;
constexpr uint32_t kBitsInf = 0x7f800000u;
constexpr uint32_t kCodeZero = 0, kCodeTiny = 1, kCodeSat = 2, kCodeNaN = 3;

// float32(2^(j/16)): what the reference's pow(2, e + j/16) returns (make_golden.log).
#define SLFP_POW2FRAC_TABLE                                                                       \
    {0x3f800000u, 0x3f85aac3u, 0x3f8b95c2u, 0x3f91c3d3u, 0x3f9837f0u, 0x3f9ef532u, 0x3fa5fed7u,   \
     0x3fad583fu, 0x3fb504f3u, 0x3fbd08a4u, 0x3fc5672au, 0x3fce248cu, 0x3fd744fdu, 0x3fe0ccdfu,   \
     0x3feac0c7u, 0x3ff5257du}
// first float32 mantissa pattern in [1,2) that quantize_weight(8) maps to log-code j = 1..16.
#define SLFP_WGT_THRESH_TABLE                                                                     \
    {0x3f82cd87u, 0x3f88980fu, 0x3f8ea43au, 0x3f94f4f0u, 0x3f9b8d3au, 0x3fa27043u, 0x3fa9a15bu,   \
     0x3fb123f6u, 0x3fb8fbb0u, 0x3fc12c4du, 0x3fc9b9beu, 0x3fd2a81eu, 0x3fdbfbb8u, 0x3fe5b907u,   \
     0x3fefe4bau, 0x3ffa83b3u}

static __device__ __constant__ uint32_t c_pow2frac[16] = SLFP_POW2FRAC_TABLE;
static const uint32_t h_pow2frac[16] = SLFP_POW2FRAC_TABLE;   // host copy (host-compiled checks)

// bit casts / clz usable from host-compiled code too: the encode/decode logic below is compiled
// for the host by tests/host_check.cu and swept against the oracle without a GPU.
__host__ __device__ __forceinline__ uint32_t f2u(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
__host__ __device__ __forceinline__ float u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
__host__ __device__ __forceinline__ int clz32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __clz((int)v);
#else
    return v ? __builtin_clz(v) : 32;
#endif
}

// IEEE-754 round-to-nearest float32 division: `input / self.Ka` on the reference's CPU path is a
// true division by float32(K) (not a multiply by the reciprocal); SURVEY.md section 8c.
__host__ __device__ __forceinline__ float div_rn(float x, float k) {
#ifdef __CUDA_ARCH__
    return __fdiv_rn(x, k);
#else
    return x / k;
#endif
}

// The same correctly rounded quotient without the division instruction sequence: the divisor K is a
// per-layer constant, so rk = RN(1/K) is computed once on the host (`1.0f / k`, IEEE) and
//      q0 = RN(x * rk);  r = x - q0*K (exact, one FMA);  q1 = RN(q0 + r * rk)
// is RN(x / K) (Markstein's final-step theorem: q0 is within one ulp and rk is the correctly rounded
// reciprocal).  The argument needs q0, r and q1 to stay inside the normal range, so anything with an
// extreme exponent (|x| outside [2^-60, 2^60], Inf, NaN, 0) takes the true division.  `valid` is
// decided on the host: K normal and 2^-30 <= K <= 2^30.  tests/test_host_compiled_kernels.py sweeps
// this against x / K for every float32 mantissa.
struct DivK {
    float k, rk;
    int fast;
};
__host__ __device__ static inline DivK make_divk(float k) {
    DivK d;
    d.k = k;
    d.rk = 1.0f / k;
    const float a = k < 0 ? -k : k;
    d.fast = (k >= 9.3132257e-10f && k <= 1.0737418e9f) ? 1 : 0;     // positive scales only (sign of a zero quotient)
    return d;
}
__host__ __device__ __forceinline__ float div_k(float x, const DivK& d) {
    const uint32_t ax = f2u(x) & 0x7fffffffu;
    // 2^-60 = 0x21800000, 2^60 = 0x5d800000
    if (d.fast && (ax - 0x21800000u) < (0x5d800000u - 0x21800000u)) {
        const float q0 = x * d.rk;
        const float r = fmaf(-q0, d.k, x);
        return fmaf(r, d.rk, q0);
    }
    return div_rn(x, d.k);
}

// Branch-free variant for the fused conv epilogue: always the reciprocal sequence.  Outside the range
// where the sequence is exact the quotient only decides a CLASS (saturated / tiny / zero / NaN), which it
// still gets right: an overflowing intermediate yields NaN or Inf (encode_q maps both to "saturated"
// unless the dividend itself was NaN) and a vanishing one yields a sub-0.0625 value ("tiny") or 0 for 0.
// (Only a dividend below 2^-119 whose true quotient rounds to exactly 0 can be classed "tiny" instead of
// "zero"; conv outputs that small do not occur and the stand-alone quantizer uses div_k.)
__host__ __device__ __forceinline__ float div_k_fused(float x, const DivK& d) {
    const float q0 = x * d.rk;
    const float r = fmaf(-q0, d.k, x);
    return fmaf(r, d.rk, q0);
}

// ---- encode ------------------------------------------------------------------------------------
template <int FMT>
__host__ __device__ __forceinline__ uint32_t encode(float v) {
    constexpr uint32_t kThresh[16] = SLFP_WGT_THRESH_TABLE;
    const uint32_t b = f2u(v);
    const uint32_t a = b & 0x7fffffffu;
    const uint32_t sg = (b >> 24) & 0x80u;
    uint32_t u;
    if (FMT == SLFP_FMT_SFP33) {
        // round-half-even of 8*m at mantissa bit 20; the carry runs into the exponent field
        const uint32_t r = a + 0x7ffffu + ((a >> 20) & 1u);
        u = (r >> 20) - ((127u - 4u) << 3);
        u = (a >= kBits15) ? 63u : u;                        // a >= 15 -> 15      (:29,:77)
        u = (a < kBits0125) ? 8u : u;                        // [0.0625,0.125) -> 0.125
    } else {
        const uint32_t mant = a & 0x7fffffu;
        uint32_t L;
        if (FMT == SLFP_FMT_SLFP34_ACT) {
            // m_q = rne(16 m)/16 (:88), then the log converter (:89): 0,1,3,4,...,15,15,16
            uint32_t i = mant >> 19;
            const uint32_t frac = mant & 0x7ffffu;
            i += (frac > 0x40000u || (frac == 0x40000u && (i & 1u))) ? 1u : 0u;
            L = i + ((i >= 2u && i <= 14u) ? 1u : 0u);
        } else {
            // L = round(16 log2 m) (:40) == number of thresholds <= m (16 immediate compares)
            const uint32_t mb = mant | 0x3f800000u;
            L = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) L += (mb >= kThresh[j]) ? 1u : 0u;
        }
        u = (((a >> 23) - (127u - 4u)) << 4) + L;            // L == 16 carries into E
        u = (a > kBitsSat8) ? kCodeSat : u;                  // a > 15.32165 -> literal (:46,:95)
        u = (a < kBits0125) ? 16u : u;
    }
    u = (a < kBits0625) ? kCodeTiny : u;                     // a < 0.0625 -> 1e-10
    u |= sg;
    u = (a == 0u) ? kCodeZero : u;                           // sign(+-0) = 0
    u = (a > kBitsInf) ? kCodeNaN : u;
    return u;
}

// encode() for a quotient q = div_k_fused(src, K): NaN is decided by the dividend; a NaN / Inf quotient
// produced by overflow inside the reciprocal sequence falls into the saturation class.
template <int FMT>
__host__ __device__ __forceinline__ uint32_t encode_q(float q, float src) {
    const uint32_t bq = f2u(q);
    uint32_t u;
    if ((bq & 0x7fffffffu) > kBitsInf) {        // quotient NaN (overflow artefact or NaN dividend)
        const uint32_t bs = f2u(src);
        u = ((bs & 0x7fffffffu) > kBitsInf) ? kCodeNaN : ((FMT == SLFP_FMT_SFP33 ? 63u : kCodeSat) | ((bs >> 24) & 0x80u));
    } else {
        u = encode<FMT>(q);
    }
    return u;
}

// encode() specialised for q >= +0 and not NaN (every quantize-on-store in the nets follows a ReLU):
// no sign / NaN handling, and the SLFP round-half-even of 16 m plus the log-converter correction are done
// with float adds on the FMA pipe (magic-number rounding, saturating adds) instead of integer-ALU ops.
__host__ __device__ __forceinline__ float sat01(float x) {
#ifdef __CUDA_ARCH__
    return __saturatef(x);
#else
    return x < 0.f ? 0.f : (x > 1.f ? 1.f : x);
#endif
}
template <int FMT>
__host__ __device__ __forceinline__ uint32_t encode_relu(float q) {
    const uint32_t b = f2u(q);
    uint32_t u;
    if (FMT == SLFP_FMT_SFP33) {
        const uint32_t r = b + 0x7ffffu + ((b >> 20) & 1u);
        u = (r >> 20) - ((127u - 4u) << 3);
        u = (b >= kBits15) ? 63u : u;
        u = (b < kBits0125) ? 8u : u;
    } else {
        const float m16 = u2f((b & 0x007fffffu) | 0x41800000u);          // 16 m in [16, 32)
        const float r = m16 + 12582912.0f;                               // rne(16 m) in the low mantissa bits
        const float c1 = sat01(r - 12582929.0f);                         // 1 if rne >= 18  (i >= 2)
        const float c2 = sat01(r - 12582942.0f);                         // 1 if rne >= 31  (i >= 15)
        const float r2 = (r + c1) - c2;                                  // + log-converter correction
        u = ((b >> 19) & 0xff0u) + f2u(r2) - (0x4b400000u + 1984u);
        u = (b > kBitsSat8) ? kCodeSat : u;
        u = (b < kBits0125) ? 16u : u;
    }
    const uint32_t tz = b < 1u ? b : 1u;                                 // 0 for +0, else "tiny"
    u = (b < kBits0625) ? tz : u;
    return u;
}

// encode<FMT>() for every v that is not NaN: ~19 instructions instead of ~30 (the stand-alone quantizer is
// HBM-bound only if the encoder stays under ~22 instructions per element).  The caller checks max |v| of a
// 4-element group and takes encode<>() for the rare group that contains a NaN.  Bit-exact with encode<FMT>() on
// that domain: swept for every mantissa by
// tests/test_host_compiled_kernels.py::test_fast_encoder_equals_reference_encoder.
template <int FMT>
__host__ __device__ __forceinline__ uint32_t encode_inrange(float v) {
    const uint32_t b = f2u(v);
    const uint32_t a = b & 0x7fffffffu;
    const uint32_t sg = (b >> 24) & 0x80u;
    uint32_t u;
    if (FMT == SLFP_FMT_SFP33) {
        const uint32_t r = a + 0x7ffffu + ((a >> 20) & 1u);          // round-half-even of 8 m, carry into the exponent
        u = (r >> 20) - ((127u - 4u) << 3);
        u = (int32_t)u < 8 ? 8u : u;                                 // [0.0625, 0.125) -> 0.125 (smaller handled below)
        u = u > 63u ? 63u : u;                                       // a >= 15 -> 15
    } else if (FMT == SLFP_FMT_SLFP34_ACT) {
        const uint32_t r = a + 0x3ffffu + ((a >> 19) & 1u);          // round-half-even of 16 m
        const uint32_t t = r >> 19;                                  // (exponent << 4) | i, i = 16 carried
        // log converter 0,1,3,4,...,15,15,16: +1 for 2 <= i <= 14; bit (t & 31) of the mask (exponent parity x i)
#ifdef __CUDA_ARCH__
        const uint32_t corr = (__funnelshift_r(0x7ffc7ffcu, 0u, t) ) & 1u;
#else
        const uint32_t corr = (0x7ffc7ffcu >> (t & 31u)) & 1u;
#endif
        u = t + corr - ((127u - 4u) << 4);
        u = (int32_t)u < 16 ? 16u : u;
        u = (a > kBitsSat8) ? kCodeSat : u;                          // a > 15.32165 (and Inf) -> the literal
    } else {
        // weights: number of thresholds <= mantissa, found from the top 5 mantissa bits (each 1/32 bucket holds at
        // most one of the 16 thresholds): (thresholds below the bucket, the threshold inside it or ~0)
        constexpr uint32_t kThresh[16] = SLFP_WGT_THRESH_TABLE;
        const uint32_t mb = (a & 0x007fffffu) | 0x3f800000u;
        uint32_t L = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) L += (mb >= kThresh[j]) ? 1u : 0u;
        u = (((a >> 23) - (127u - 4u)) << 4) + L;
        u = (int32_t)u < 16 ? 16u : u;
        u = (a > kBitsSat8) ? kCodeSat : u;
    }
    const uint32_t z = a < 1u ? a : 1u;                              // 0 for +-0, else 1
    const uint32_t low = z * sg + z;                                 // +-0 -> 0 (sign dropped), tiny -> 1 | sign
    return (a < kBits0625) ? low : (u | sg);
}

// The same encoder with its work split between the two CUDA-core pipes.  ALU (LOP3 / SHF / IMNMX / ISETP / SEL /
// FMNMX) and FMA (FADD / FMUL / FFMA / IMAD) each issue one warp instruction per two cycles and sub-partition: an
// integer-only encoder is ALU-bound at 64 lanes/clk/SM and holds the stand-alone quantizer at ~0.6 of the HBM
// peak.  Here the mantissa rounding is a Veltkamp split on the FMA pipe - t = |v| (2^s + 1); hi = t - (t - |v|)
// is |v| rounded to nearest-even at 23 - s mantissa bits, verified against the integer rounding for every
// mantissa - and the sign / bias arithmetic are IMADs.  Domain: v not NaN (Inf is fine: it is clamped first).
template <int FMT>
__host__ __device__ __forceinline__ uint32_t encode_balanced(float v) {
    static_assert(FMT == SLFP_FMT_SFP33 || FMT == SLFP_FMT_SLFP34_ACT, "activation formats");
    const uint32_t b = f2u(v);
    const float av = fabsf(v);
    const uint32_t sbit = b >> 31;
    uint32_t u;
    float avc;
    if (FMT == SLFP_FMT_SFP33) {
        avc = av < 15.0f ? av : 15.0f;                                // a >= 15 -> 15 (:29 / :77); also tames Inf
        const float t = avc * 1048577.0f;                              // 2^20 + 1
        const float d = t - avc;
        const float hi = t - d;                                        // round-half-even at 3 mantissa bits
        u = (f2u(hi) >> 20) - ((127u - 4u) << 3);
        u = (int32_t)u < 8 ? 8u : u;                                   // [0.0625, 0.125) -> 0.125
    } else {
        avc = av < 16.0f ? av : 16.0f;                                 // keeps the split finite; saturation is decided on av
        const float t = avc * 524289.0f;                               // 2^19 + 1
        const float d = t - avc;
        const float hi = t - d;                                        // round-half-even of 16 m (:88)
        const uint32_t tt = f2u(hi) >> 19;                             // (exponent << 4) | i, i = 16 carried
#ifdef __CUDA_ARCH__
        const uint32_t corr = __funnelshift_r(0x7ffc7ffcu, 0u, tt) & 1u;   // log converter (:89): +1 for 2 <= i <= 14
#else
        const uint32_t corr = (0x7ffc7ffcu >> (tt & 31u)) & 1u;
#endif
        u = tt + corr - ((127u - 4u) << 4);
        u = (int32_t)u < 16 ? 16u : u;
        u = (f2u(av) > kBitsSat8) ? kCodeSat : u;                      // a > 15.32165 (and Inf) -> the literal (:95)
    }
    const uint32_t a = f2u(avc);
    const uint32_t z = a < 1u ? a : 1u;                                // 0 for +-0, else 1
    const uint32_t hi_code = sbit * 128u + u;
    const uint32_t lo_code = (z * sbit) * 128u + z;                    // +-0 -> 0 (sign dropped), tiny -> 1 | sign
    return (a < kBits0625) ? lo_code : hi_code;
}

// ---- table encoder (stand-alone activation quantizer, codes output) ------------------------------------------
// encode_balanced<> still spends ~15 ALU-pipe instructions per element on class selection (clamps, tiny / zero /
// saturation selects, the log-converter correction, sign insertion) and that pipe issues 64 lanes/clk/SM: the kernel
// sat at 0.64 of the HBM peak.  Here EVERY class decision is folded into the value that gets rounded, with
// saturating FMA-pipe arithmetic, and the unsigned code is one byte load from a shared-memory table indexed by the
// top bits of the rounded float:
//     m    = |q| * 2^100                                   (exact, may overflow to +Inf)
//     nz   = sat(|x| * 2^126)                              1 unless the dividend is +-0
//     tiny = sat(nz * 2^96 - m)                            1 iff 0 < |q| < 0.0625          (:92 / :74)
//     sat  = sat(m - 15.32165 * 2^100)                     1 iff |q| > 15.32165            (:95; SLFP only)
//     a'   = min(|q|, 16 | 15) + 48 sat + 128 tiny         saturated -> 64, tiny -> 128, +-0 -> 0, else unchanged
//     hi   = Veltkamp split of a' at 4 | 3 mantissa bits   (round-half-even, :88 / :69)
//     code = table[bits(hi) >> 19 | 20]                    table holds the log converter (:89) and [0.0625,0.125) -> 0.125
// 13 FMA-pipe + 2 ALU-pipe instructions per element; sign bits are gathered four at a time (PRMT) and OR-ed into the
// packed word.  Domain (the quantizer's group probe guarantees it): q finite, x == +0 or |x| >= 2^-119, never -0
// (a zero code carries no sign).  Swept against encode<> for every mantissa by tests/test_host_compiled_kernels.py.
constexpr int kEncLutBytes = 2176;
template <int FMT>
__host__ __device__ __forceinline__ uint32_t enc_lut_entry(uint32_t idx) {
    static_assert(FMT == SLFP_FMT_SFP33 || FMT == SLFP_FMT_SLFP34_ACT, "activation formats");
    constexpr uint32_t mb = FMT == SLFP_FMT_SFP33 ? 3u : 4u;
    if (idx == 0u) return kCodeZero;
    if (idx == (134u << mb)) return kCodeTiny;                    // 128.0
    if (idx >= (133u << mb)) return kCodeSat;                     // 64.0 (and the unreachable rest)
    const uint32_t eb = idx >> mb, i = idx & ((1u << mb) - 1u);
    if (eb < 123u) return kCodeTiny;                              // unreachable (tiny values were moved to 128)
    const uint32_t E = eb - 123u;
    if (E == 0u) return 1u << mb;                                 // [0.0625, 0.125) -> 0.125
    if (FMT == SLFP_FMT_SFP33) return (E << 3) + i;
    return (E << 4) + i + ((i >= 2u && i <= 14u) ? 1u : 0u);      // log converter 0,1,3,4,...,15,15,(16 = carried)
}
template <int FMT>
__host__ __device__ __forceinline__ uint32_t enc_lut_index(float q, float x) {
    const float av = fabsf(q);
    const float nz = sat01(fabsf(x) * 0x1p126f);
    const float m = av * 0x1p100f;
    const float tiny = sat01(fmaf(nz, 0x1p96f, -m));
    float a;
    if (FMT == SLFP_FMT_SFP33) {
        a = fmaf(tiny, 128.0f, fminf(av, 15.0f));                 // a >= 15 -> 15 (:77)
    } else {
        const float st = sat01(m - u2f(kBitsSat8 + (100u << 23)));
        a = fmaf(tiny, 128.0f, fmaf(st, 48.0f, fminf(av, 16.0f)));
    }
    const float t = a * (FMT == SLFP_FMT_SFP33 ? 1048577.0f : 524289.0f);
    const float d = t - a;
    const float hi = t - d;
    return f2u(hi) >> (FMT == SLFP_FMT_SFP33 ? 20 : 19);
}

// ---- SLFP<3,4> weight encoder, threshold search by table ----------------------------------------------------------
// quantize_weight(8) has no linear pre-round (sfp_quant.py:40): L = number of the 16 thresholds 2^((2j-1)/32) that are
// <= the mantissa.  Consecutive thresholds are >= 0.043 apart, so each 1/32-wide mantissa bucket holds at most one:
// entry b = (thresholds at or below the bucket's start, the threshold inside the bucket or ~0) and
// L = cnt + (mantissa >= thr): one 8-byte load and a compare instead of 16 compares.  Bit-exact with encode<>
// (swept for every mantissa by tests/test_host_compiled_kernels.py).
__host__ __device__ __forceinline__ void wgt_bucket_entry(uint32_t b, uint32_t& cnt, uint32_t& thr) {
    constexpr uint32_t kThresh[16] = SLFP_WGT_THRESH_TABLE;
    const uint32_t lo = 0x3f800000u | (b << 18), hi = lo + (1u << 18);
    cnt = 0; thr = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (kThresh[j] <= lo) ++cnt;
        else if (kThresh[j] < hi) thr = kThresh[j];
    }
}
__host__ __device__ __forceinline__ uint32_t encode_wgt_bucket(float v, const uint2* __restrict__ tbl) {
    const uint32_t b = f2u(v);
    const uint32_t a = b & 0x7fffffffu;
    const uint32_t mb = (a & 0x007fffffu) | 0x3f800000u;
    const uint2 e = tbl[(a >> 18) & 31u];
    uint32_t u = (((a >> 23) - (127u - 4u)) << 4) + e.x + (mb >= e.y ? 1u : 0u);   // L == 16 carries into E
    u = (int32_t)u < 16 ? 16u : u;                           // [0.0625, 0.125) -> 0.125
    u = (a > kBitsSat8) ? kCodeSat : u;
    u = (a < kBits0625) ? kCodeTiny : u;
    u |= (b >> 24) & 0x80u;
    u = (a == 0u) ? kCodeZero : u;
    u = (a > kBitsInf) ? kCodeNaN : u;
    return u;
}

__host__ __device__ __forceinline__ uint32_t encode_rt(float v, int fmt) {
    if (fmt == SLFP_FMT_SFP33) return encode<SLFP_FMT_SFP33>(v);
    if (fmt == SLFP_FMT_SLFP34_ACT) return encode<SLFP_FMT_SLFP34_ACT>(v);
    return encode<SLFP_FMT_SLFP34_WGT>(v);
}

// quantize_layerout, SFP<4,4> (sfp_quant.py:111-126).  Lines 122-123 of the reference use '^'
// (XOR), so there is no low clamp and exact 0 gives NaN; zero_is_zero selects the intended value.
__host__ __device__ __forceinline__ float layerout_quantize(float v, bool zero_is_zero) {
    const uint32_t b = f2u(v);
    const uint32_t a = b & 0x7fffffffu, s = b & 0x80000000u;
    uint32_t r;
    if (a >= 0x00800000u) {
        r = (a + 0x3ffffu + ((a >> 19) & 1u)) & ~0x7ffffu;
    } else {                                                  // denormal input
        const int sh = (31 - clz32(a | 1u)) - 4;
        r = a;
        if (sh > 0) r = (a + ((1u << (sh - 1)) - 1u) + ((a >> sh) & 1u)) & ~((1u << sh) - 1u);
    }
    r = (a >= kBits248) ? kBits248 : r;
    r |= s;
    if (a == 0u) r = zero_is_zero ? 0u : 0xffc00000u;
    if (a > kBitsInf) r = 0x7fc00000u;
    return u2f(r);
}

// ---- decode ------------------------------------------------------------------------------------
// tab: the 16-entry pow2frac table (shared or constant memory)
template <bool SFP33>
__host__ __device__ __forceinline__ float decode(uint32_t code, const uint32_t* __restrict__ tab) {
    const uint32_t s = (code & 0x80u) << 24;
    const uint32_t u = code & 0x7fu;
    uint32_t r;
    if (SFP33) {
        r = (((u >> 3) + 123u) << 23) | ((u & 7u) << 20);
        r = (u < 8u) ? 0x7fc00000u : r;
    } else {
        r = tab[u & 15u] + (((u >> 4) - 4u) << 23);
        r = (u < 16u) ? 0x7fc00000u : r;
        r = (u == kCodeSat) ? kBitsSat8 : r;
    }
    r = (u == kCodeTiny) ? kBitsTiny : r;
    r |= s;
    r = (u == kCodeZero) ? 0u : r;
    r = (u == kCodeNaN) ? 0x7fc00000u : r;
    return u2f(r);
}

// Operand value fed to the tensor cores for a code: RN_fp16(decode(code)).  +-1e-10 underflows to
// +-0 and the saturation literal rounds to the same half as the top grid value.
template <bool SFP33>
__device__ __forceinline__ uint16_t decode_f16_bits(uint32_t code, const uint32_t* __restrict__ tab) {
    const float f = decode<SFP33>(code, tab);
    return __half_as_ushort(__float2half_rn(f));
}

// ---- post-ReLU activation codes of the fused pipeline (SLFP_FMT_SLFP34_RELU / SLFP_FMT_SFP33_RELU) ----
// Between two fused layers the producer's epilogue has already applied the ReLU, so the value is >= +0 and
// the code needs no sign.  The unsigned code is the TRUNCATED float32 bit pattern, re-based, with one more
// mantissa bit than the format keeps (a half step):
//     SLFP<3,4>:  c = (bits(q) >> 18) - 0xF5F      q = 0.0625 -> 1, 0.125 -> 33, 15.25.. -> 255
//     SFP<3,3> :  c = (bits(q) >> 19) - 0x7AF      q = 0.0625 -> 1, 0.125 -> 17, 15      -> 127
// saturated to [0, 255].  Truncation keeps every class boundary of the reference exact (a < 0.0625, a < 0.125,
// the saturation thresholds compare the UN-rounded value, sfp_quant.py:92-95 / :74-77); the rounding itself
// - the reference's linear pre-round `round(2^mbits m)` (:88 / :69) - is finished by the CONSUMER's decode
// table from the half-step bit, which also applies the SLFP log converter of :89:
//     c = 0               -> 0         (everything below 0.0625: the reference's 1e-10 is 0 in float16)
//     u = c - 1 = E*2H+h  -> E = 0: 0.125 ([0.0625, 0.125) -> 0.125, :93 / :75); E >= 1: mantissa index
//                            i = (h + 1) >> 1, value 2^(E-4) * grid[i], clamped to the top value (:95 / :77)
// Two deviations from encode<>, both far below the float32 noise of the convolution that feeds the encoder:
// exact ties of the mantissa rounding (all 18 / 19 dropped bits equal to 10...0) go up instead of to even,
// and the quotient q = y * (1/Ka) is a reciprocal multiply.
template <bool SFP33>
__host__ __device__ __forceinline__ int32_t encode_relu_fast_raw(float q_nonneg) {
    const int32_t b = (int32_t)f2u(q_nonneg);
    return SFP33 ? ((b >> 19) - 0x7AF) : ((b >> 18) - 0xF5F);
}
// The same on q/16 (the epilogue computes clamp(q/16, 0, 1) with one saturating FMA: ReLU, the scale by 1/Ka
// and an upper clamp in a single instruction); 1.0 lands above the top code and saturates in the pack.
template <bool SFP33>
__host__ __device__ __forceinline__ int32_t encode_relu_fast_raw16(float q16_sat) {
    const int32_t b = (int32_t)f2u(q16_sat);
    return SFP33 ? ((b >> 19) - 0x76F) : ((b >> 18) - 0xEDF);
}
template <bool SFP33>
__host__ __device__ __forceinline__ uint32_t encode_relu_fast(float q) {
    const int32_t t = encode_relu_fast_raw<SFP33>(q > 0.0f ? q : 0.0f);
    return (uint32_t)(t < 0 ? 0 : (t > 255 ? 255 : t));
}
template <bool SFP33>
__host__ __device__ __forceinline__ float decode_relu(uint32_t code, const uint32_t* __restrict__ tab) {
    const uint32_t c = code & 0xffu;
    if (c == 0u) return 0.0f;
    const uint32_t u = c - 1u;
    if (SFP33) {
        const uint32_t E = u >> 4, i = ((u & 15u) + 1u) >> 1;            // i in 0..8
        if (E == 0u) return 0.125f;
        const uint32_t r = ((E + 123u) << 23) + (i << 20);               // i = 8 carries into the exponent
        return u2f(r > kBits15 ? kBits15 : r);
    }
    const uint32_t E = u >> 5, i = ((u & 31u) + 1u) >> 1;                // i in 0..16
    if (E == 0u) return 0.125f;
    const uint32_t L = i + ((i >= 2u && i <= 14u) ? 1u : 0u);            // log converter; 16 = next octave
    const uint32_t r = (L == 16u ? 0x40000000u : tab[L]) + ((E - 4u) << 23);
    const uint32_t top = tab[15] + (3u << 23);                           // 2^(3 + 15/16)
    return u2f(r > top ? top : r);
}

// relu(quantize_layerout(y)) for the fused fast epilogues (conv -> BN -> layerout_quantize_func -> ReLU,
// nets_cifar/shufflenet_v2.py:64-72): round-half-even to 4 mantissa bits, clamp at 248, negative -> 0.  The fast paths
// map an exact 0 to 0 (the reference's XOR typo yields NaN there; the generic epilogue reproduces that when asked).
__host__ __device__ __forceinline__ float layerout_relu(float y) {
    // Veltkamp split at 2^19 + 1: hi = y rounded half-even to 5 significant bits (the same three FMA-pipe operations the
    // table encoder uses, enc_lut_index); the library is compiled with -fmad=false, so nothing here is contracted.
    // Bit-exact with relu(layerout_quantize(y, true)) for every normal float (tests/test_host_compiled_kernels.py).
    const float t0 = fminf(fmaxf(y, 0.0f), 512.0f);
    const float t = t0 * 524289.0f;
    const float d = t - t0;
    return fminf(t - d, 248.0f);
}

// ---- e4m3 storage of SFP<3,3> values (SLFP_FMT_E4M3) -----------------------------------------------------------------
// decode: [s][e:4][m:3], bias 7 (sub-normals m * 2^-9; the encoders only ever produce +-0 of them)
__host__ __device__ __forceinline__ float decode_e4m3(uint32_t c) {
    const uint32_t e = (c >> 3) & 15u, m = c & 7u, s = (c & 0x80u) << 24;
    if (e == 0u) return u2f(s | f2u((float)m * 0.001953125f));
    return u2f(s | ((e + 120u) << 23) | (m << 20));
}
// the SFP<3,3> quantizer (utils/sfp_quant.py:63-78) of an already pre-scaled value q, as an e4m3 byte:
// |q| < 0.0625 -> +-0 (the reference's +-1e-10), [0.0625, 0.125) -> 0.125, >= 15 -> 15, else round-half-even to 3
// mantissa bits.  NaN -> 0x7f.
__host__ __device__ __forceinline__ uint32_t encode_e4m3(float q) {
    const uint32_t b = f2u(q), s = (b >> 24) & 0x80u;
    uint32_t a = b & 0x7fffffffu;
    if (a > kBitsInf) return 0x7fu;
    if (a < 0x3d800000u) return s;                            // < 0.0625
    if (a < 0x3e000000u) a = 0x3e000000u;                     // [0.0625, 0.125) -> 0.125
    if (a >= kBits15) a = kBits15;
    const uint32_t r = (a + 0x7ffffu + ((a >> 20) & 1u)) >> 20;     // round half even to 3 mantissa bits: (exp << 3) | m
    const uint32_t top = (130u << 3) | 7u;                    // 15 = 1.875 * 2^3
    return s | ((r > top ? top : r) - (120u << 3));
}
// two pre-scaled values -> two e4m3 bytes (lo in bits 0-7): the clamps on the FMA / ALU pipes, the rounding in ONE
// cvt.rn.satfinite.e4m3x2.f32 (round-half-even).  Bit-exact with encode_e4m3 (tests/test_gpu_quantizer.py sweeps it).
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t encode_e4m3x2(float lo, float hi) {
#if defined(__CUDA_ARCH__)
    float al = fminf(fabsf(lo), 15.0f), ah = fminf(fabsf(hi), 15.0f);
    al = al < 0.0625f ? 0.0f : fmaxf(al, 0.125f);
    ah = ah < 0.0625f ? 0.0f : fmaxf(ah, 0.125f);
    uint16_t d;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(d) : "f"(copysignf(ah, hi)), "f"(copysignf(al, lo)));
    return (uint32_t)d;
#else
    return encode_e4m3(lo) | (encode_e4m3(hi) << 8);
#endif
}
#endif
// the same after a ReLU (negative inputs -> 0): no sign handling, ~4 clamp instructions per element + half a cvt
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t encode_e4m3x2_relu(float lo, float hi) {
#if defined(__CUDA_ARCH__)
    float al = fminf(lo, 15.0f), ah = fminf(hi, 15.0f);
    al = al < 0.0625f ? 0.0f : fmaxf(al, 0.125f);               // also sends negatives (and -0) to +0
    ah = ah < 0.0625f ? 0.0f : fmaxf(ah, 0.125f);
    uint16_t d;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(d) : "f"(ah), "f"(al));
    return (uint32_t)d;
#else
    return encode_e4m3(lo > 0.f ? lo : 0.f) | (encode_e4m3(hi > 0.f ? hi : 0.f) << 8);
#endif
}
#endif
// SFP<3,3> weight / activation code ([s][E:3][m:3], escapes below 8) -> the e4m3 byte of the same value
__host__ __device__ __forceinline__ uint32_t sfp33_code_to_e4m3(uint32_t code) {
    const uint32_t u = code & 0x7fu;
    if (u < 8u) return code & 0x80u;                          // 0 and +-1e-10 -> +-0 (NaN never reaches a weight operand)
    return (code & 0x80u) | (u + 24u);                        // E + 3 in the exponent field: ((E + 3) << 3) | m
}

// ---- SLFP<3,4> weight encoder as ONE table look-up (wprep_rows_kernel in quantize.cu) ---------------------------------
// Domain: finite non-zero quotients (the kernel's group probe guarantees it).  Entry (|q| bits >> 18) - clamped to
// [below 0.0625 | the 8 x 32 buckets of [0.0625, 16) | 16 and above] - holds the bucket's only decision threshold on the
// full bit pattern (a rounding threshold 2^((2j-1)/32) of sfp_quant.py:40, or the saturation bound 15.32165 of :46 - never
// both: the bound lies in bucket 29 of [8, 16), the nearest thresholds in 27 and 30), the unsigned codes below /
// at-and-above it and the float16 images of their values.  11 instead of 26 instructions per weight; bit-exact with
// encode_wgt_bucket() / encode<SLFP34_WGT>() on that domain: swept for every mantissa on the host
// (tests/test_host_compiled_kernels.py::test_weight_lut_encoder_equals_reference_encoder) and on the GPU against the oracle.
constexpr int kWgtLutEntries = 258;
struct WgtLutEntry { uint32_t thr, lo, hi, halves; };        // halves: float16(decode(lo)) | float16(decode(hi)) << 16
__host__ __device__ __forceinline__ uint32_t f16_bits_rn(float v) { return (uint32_t)__half_as_ushort(__float2half_rn(v)); }
__host__ __device__ inline WgtLutEntry wgt_lut_entry(uint32_t i, const uint32_t* __restrict__ tab) {
    uint32_t thr = 0xffffffffu, lo, hi;
    if (i == 0u) lo = hi = kCodeTiny;
    else if (i == (uint32_t)kWgtLutEntries - 1u) lo = hi = kCodeSat;
    else {
        const uint32_t e = (i - 1u) >> 5, m5 = (i - 1u) & 31u;             // |q| in 2^(e-4) * [1 + m5/32, 1 + (m5+1)/32)
        if (e == 0u) lo = hi = 16u;                                        // [0.0625, 0.125) -> 0.125
        else {
            uint32_t cnt, t;
            wgt_bucket_entry(m5, cnt, t);
            lo = (e << 4) + cnt; hi = lo + 1u;                             // L == 16 carries into the exponent
            if (t != 0xffffffffu) thr = ((123u + e) << 23) | (t & 0x007fffffu);
            const uint32_t b0 = ((123u + e) << 23) | (m5 << 18);
            if (b0 > kBitsSat8) lo = hi = kCodeSat;                        // whole bucket above 15.32165
            else if (b0 + (1u << 18) > kBitsSat8) { thr = kBitsSat8 + 1u; hi = kCodeSat; }
        }
    }
    WgtLutEntry en;
    en.thr = thr; en.lo = lo; en.hi = hi;
    en.halves = f16_bits_rn(decode<false>(lo, tab)) | (f16_bits_rn(decode<false>(hi, tab)) << 16);
    return en;
}
__host__ __device__ __forceinline__ uint32_t wgt_lut_index(uint32_t qa) {
    int t = (int)(qa >> 18) - (123 * 32 - 1);
    t = t < 0 ? 0 : t;
    return (uint32_t)(t > kWgtLutEntries - 1 ? kWgtLutEntries - 1 : t);
}
// signed code of the quotient with bit pattern qb; half16: float16 image of its value (sign included)
__host__ __device__ __forceinline__ uint32_t encode_wgt_lut(uint32_t qb, const WgtLutEntry& en, uint32_t& half16) {
    const bool up = (qb & 0x7fffffffu) >= en.thr;
    half16 = (up ? (en.halves >> 16) : (en.halves & 0xffffu)) | ((qb >> 16) & 0x8000u);
    return (up ? en.hi : en.lo) | ((qb >> 24) & 0x80u);
}

// value of an activation code in any of the code formats a dense conv accepts
__host__ __device__ __forceinline__ float decode_act_any(uint32_t code, int fmt, const uint32_t* __restrict__ tab) {
    switch (fmt) {
        case SLFP_FMT_SFP33: return decode<true>(code, tab);
        case SLFP_FMT_SLFP34_RELU: return decode_relu<false>(code, tab);
        case SLFP_FMT_SFP33_RELU: return decode_relu<true>(code, tab);
        case SLFP_FMT_E4M3: return decode_e4m3(code);
        case SLFP_FMT_SFP33_SFAST: {
            const float m = decode_relu<true>(code & 0x7fu, tab);
            return (code & 0x80u) ? -m : m;
        }
        default: return decode<false>(code, tab);
    }
}

// ---- host-side error plumbing ---------------------------------------------------------------------
int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);
int num_sms();

// "done once PER DEVICE" flag: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device property, and one process may
// drive several GPUs (the deployment model is one process per GPU, but nothing may break otherwise).  Racing threads at
// worst set the attribute twice.
struct DeviceOnce {
    bool done[64] = {};
    bool& flag() {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); dev = 0; }
        return done[dev];
    }
};

static inline __host__ __device__ size_t ceil_div_sz(size_t a, size_t b) { return (a + b - 1) / b; }

}  // namespace slfp
