// conv_igemm_sm100.cu -- dense SLFP/SFP convolution forward as an implicit GEMM on the Blackwell
// tensor cores (tcgen05.mma, accumulators in TMEM), sm_100a only.
//
// Replaces Conv2d_Q.forward / Linear_Q.forward of the reference (utils/conv2d_func.py:20-25,
// 41-47, 60-65): F.conv2d(input_q, weight_q, bias_q) * Ka * Kw on fake-quant float32 tensors.
// Here the activations stay 8-bit codes in HBM (NHWC), the weights are the float16 image of
// weight_q (KRSC), and the kernel computes
//      D[m, n] = sum_k A[m, k] * B[n, k]      m = output pixel, n = output channel, k = (r, s, c)
// with A gathered + de-quantized on the fly and the reference's post-scale / bias (and, for the
// fused eval pipeline, folded BatchNorm, residual add, ReLU and the NEXT layer's quantizer) applied
// in the epilogue straight out of TMEM.
//
// One persistent CTA per SM, 18 warps, static round-robin tile schedule (tile = 128 pixels x BLOCK_N
// channels):
//   warps 0-15  workers.  For every tile they first gather + decode its A operand, K block by K block:
//                 cp.async the code bytes (any stride / padding / dilation; zero fill for padding) into
//                 a landing ring, look the codes up in a bank-conflict-free shared-memory table
//                 (code -> float16) and write the 128x64 float16 A tile in the 128B-swizzled K-major
//                 UMMA layout; then they run the epilogue of the PREVIOUS tile (whose MMAs have drained
//                 by then): tcgen05.ld 16 columns at a time -> affine / residual / ReLU -> float32 /
//                 float16 / quantize-on-store codes.  All ALU work (decode and epilogue) is thus spread
//                 over the same 16 warps, whichever of the two dominates for a layer.
//   warp 16     TMA producer: weight tile (BLOCK_N x 64 halves) with cp.async.bulk.tensor, SW128
//   warp 17     MMA issuer: one thread, 4 x tcgen05.mma (K=16) per 64-wide K block; frees the smem
//                 stage with tcgen05.commit; accumulator double-buffered in TMEM
// Barriers: full[s]/empty[s] (smem ring), tmem_full[b]/tmem_empty[b] (accumulator ring).
#include <cudaTypedefs.h>

#include "slfp_common.cuh"
#include "sm100_ptx.cuh"

namespace slfp {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kWorkWarps = 16;
constexpr int kWorkThreads = kWorkWarps * 32;          // 512
constexpr int kTmaWarp = 16;
constexpr int kMmaWarp = 17;
constexpr int kThreads = 18 * 32;                      // 576
constexpr int kLand = 3;                               // landing-ring depth (cp.async groups in flight)
constexpr int kLandBytes = kWorkThreads * 16;          // 8 KB per landing stage
constexpr int kLutBytes = 256 * 32 * 4;                // code -> f16, one copy per bank
constexpr int kABytes = kBlockM * kBlockK * 2;         // 16 KB

struct FastDiv {
    uint32_t d, mul, shr;
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return d == 1 ? n : (__umulhi(n, mul) >> shr); }
};
static FastDiv make_fastdiv(uint32_t d) {
    FastDiv f{d, 0, 0};
    if (d > 1) {
        uint32_t lg = 31 - __builtin_clz(d);
        if (d & (d - 1)) ++lg;
        const uint32_t p = 31 + lg;
        f.mul = (uint32_t)(((1ull << p) + d - 1) / d);
        f.shr = p - 32;
    }
    return f;
}

struct IgemmParams {
    const uint8_t* x;
    int H, W, Cp, Ho, Wo;
    int R, S, sh, sw, ph, pw, dh, dw;
    uint32_t M;
    int Kout;
    int num_kb, taps;
    int m_tiles, n_tiles, num_tiles;
    FastDiv div_hw, div_w, div_cpt, div_s;
    int cblocks;               // Cp / 64 when Cp % 64 == 0 (uniform-tap gather), else 0
    int sfp33;                 // activation code layout: 1 = SFP<3,3>, 0 = SLFP<3,4>
    SlfpEpilogue epi;
    DivK next_div, next_div2;  // quantize-on-store divisors with their host-computed reciprocals
};

template <int BLOCK_N>
struct Cfg {
    static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (BLOCK_N >= 256) ? 3 : 4;
    static constexpr int kTmemCols = 2 * BLOCK_N;
    static constexpr int kSmemBytes = kStages * kStageBytes + kLand * kLandBytes + kLutBytes + 256 + 1024;
};

// ---- epilogue of one tile, executed by all 16 worker warps ------------------------------------------------
template <int BLOCK_N>
__device__ __forceinline__ void epilogue_tile(const IgemmParams& p, int tile, uint32_t tmem_acc, int warp, int lane) {
    const SlfpEpilogue& e = p.epi;
    const int Kout = p.Kout;
    const bool vec4 = (Kout & 3) == 0, vec8 = (Kout & 7) == 0;
    const int quad = warp & 3;                        // TMEM lane quadrant this warp may access
    const int cgrp = warp >> 2;                       // which quarter of the tile's columns
    constexpr int kColsPerWarp = BLOCK_N / 4;
    const uint32_t m = (uint32_t)(tile / p.n_tiles) * kBlockM + (uint32_t)(quad * 32 + lane);
    const int n_base = (tile % p.n_tiles) * BLOCK_N + cgrp * kColsPerWarp;
    const bool row_ok = m < p.M;
    const bool folded = e.ch_mul != nullptr;
#pragma unroll 1
    for (int ch = 0; ch < kColsPerWarp / 16; ++ch) {
        const int n0 = n_base + ch * 16;
        const bool store_f = n0 < Kout;                                      // warp-uniform
        const bool store_c = (e.y_codes != nullptr) && n0 < e.k_phys_out;     // warp-uniform
        if (!store_f && !store_c) continue;
        uint32_t acc[16];
        ptx::tmem_ld16(tmem_acc + ((uint32_t)(quad * 32) << 16) + cgrp * kColsPerWarp + ch * 16, acc);
        const bool full = n0 + 16 <= Kout;                                     // warp-uniform
        const size_t off = (size_t)m * Kout + n0;
        // residual: issued before the TMEM wait so its latency overlaps
        float res[16];
        const bool has_res = e.residual != nullptr;
        if (has_res) {
            if (row_ok && full && vec8 && e.residual_f16) {
                const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(e.residual) + off);
                const uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
                const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&rw[i]));
                    res[2 * i] = f.x; res[2 * i + 1] = f.y;
                }
            } else if (row_ok && full && vec4 && !e.residual_f16) {
                const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(e.residual) + off);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 f = __ldg(rp + i);
                    res[4 * i] = f.x; res[4 * i + 1] = f.y; res[4 * i + 2] = f.z; res[4 * i + 3] = f.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    res[i] = 0.f;
                    if (row_ok && n0 + i < Kout)
                        res[i] = e.residual_f16 ? __half2float(reinterpret_cast<const __half*>(e.residual)[off + i])
                                                : reinterpret_cast<const float*>(e.residual)[off + i];
                }
            }
        }
        ptx::tmem_ld_wait();
        float v[16];
        if (full) {
            // per-channel vectors: warp-uniform 16-byte loads (L1 broadcast)
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                if (folded) {
                    const float4 m4 = __ldg(reinterpret_cast<const float4*>(e.ch_mul + n0) + g);
                    const float4 a4 = __ldg(reinterpret_cast<const float4*>(e.ch_add + n0) + g);
                    v[4 * g + 0] = fmaf(__uint_as_float(acc[4 * g + 0]), m4.x, a4.x);
                    v[4 * g + 1] = fmaf(__uint_as_float(acc[4 * g + 1]), m4.y, a4.y);
                    v[4 * g + 2] = fmaf(__uint_as_float(acc[4 * g + 2]), m4.z, a4.z);
                    v[4 * g + 3] = fmaf(__uint_as_float(acc[4 * g + 3]), m4.w, a4.w);
                } else {
                    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), s4 = make_float4(1.f, 1.f, 1.f, 1.f), h4 = b4;
                    if (e.bias_q) b4 = __ldg(reinterpret_cast<const float4*>(e.bias_q + n0) + g);
                    if (e.ch_scale) {
                        s4 = __ldg(reinterpret_cast<const float4*>(e.ch_scale + n0) + g);
                        h4 = __ldg(reinterpret_cast<const float4*>(e.ch_shift + n0) + g);
                    }
                    const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w}, hh[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float t = __uint_as_float(acc[4 * g + j]);
                        if (e.bias_q) t += bb[j];
                        t = t * e.post_a;
                        t = t * e.post_b;
                        if (e.ch_scale) t = fmaf(t, ss[j], hh[j]);
                        v[4 * g + j] = t;
                    }
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int n = n0 + i;
                const int nc = n < Kout ? n : Kout - 1;
                float t = __uint_as_float(acc[i]);
                if (folded) {
                    t = fmaf(t, __ldg(e.ch_mul + nc), __ldg(e.ch_add + nc));
                } else {
                    if (e.bias_q) t += __ldg(e.bias_q + nc);
                    t = t * e.post_a;
                    t = t * e.post_b;
                    if (e.ch_scale) t = fmaf(t, __ldg(e.ch_scale + nc), __ldg(e.ch_shift + nc));
                }
                v[i] = t;
            }
        }
        if (has_res) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += res[i];
        }
        if (e.relu) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.0f);
        }
        if (!row_ok) continue;
        if (e.y_f32 && store_f) {
            float* yp = e.y_f32 + off;
            if (vec4) {
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                    if (n0 + i < Kout) *reinterpret_cast<float4*>(yp + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) if (n0 + i < Kout) yp[i] = v[i];
            }
        }
        if (e.y_f16 && store_f) {
            __half* yp = reinterpret_cast<__half*>(e.y_f16) + off;
            if (vec8) {
#pragma unroll
                for (int i = 0; i < 16; i += 8) {
                    if (n0 + i < Kout) {
                        __half2 h0 = __floats2half2_rn(v[i], v[i + 1]), h1 = __floats2half2_rn(v[i + 2], v[i + 3]);
                        __half2 h2 = __floats2half2_rn(v[i + 4], v[i + 5]), h3 = __floats2half2_rn(v[i + 6], v[i + 7]);
                        uint4 pk;
                        pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                        pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                        *reinterpret_cast<uint4*>(yp + i) = pk;
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) if (n0 + i < Kout) yp[i] = __float2half_rn(v[i]);
            }
        }
        if (store_c) {
            // quantize-on-store: the next layer's `quantize_act(input / Ka)` fused here.  After a ReLU the
            // value is >= +0, which the specialised encoder exploits (a NaN would already have been
            // flushed to 0 by the ReLU's fmaxf; NaN activations are outside the fused pipeline's domain).
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                uint8_t* yc = pass ? e.y_codes2 : e.y_codes;
                if (!yc) continue;
                const DivK kd = pass ? p.next_div2 : p.next_div;
                uint32_t c[16];
                const bool relu_path = e.relu && kd.k > 0.f;
                if (e.next_fmt == SLFP_FMT_SFP33) {
                    if (relu_path) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) c[i] = encode_relu<SLFP_FMT_SFP33>(div_k_fused(v[i], kd));
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) c[i] = encode_q<SLFP_FMT_SFP33>(div_k_fused(v[i], kd), v[i]);
                    }
                } else {
                    if (relu_path) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) c[i] = encode_relu<SLFP_FMT_SLFP34_ACT>(div_k_fused(v[i], kd));
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) c[i] = encode_q<SLFP_FMT_SLFP34_ACT>(div_k_fused(v[i], kd), v[i]);
                    }
                }
                if (!full) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) c[i] = (n0 + i < Kout) ? c[i] : 0u;     // zero code in pad channels
                }
                uint32_t pk[4];
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    pk[g] = __byte_perm(__byte_perm(c[4 * g], c[4 * g + 1], 0x0040), __byte_perm(c[4 * g + 2], c[4 * g + 3], 0x0040), 0x5410);
                *reinterpret_cast<uint4*>(yc + (size_t)m * e.k_phys_out + n0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
    }
}

template <int BLOCK_N, int GRAN>
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmap_w, const IgemmParams p) {
    using C = Cfg<BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    // SW128 operand tiles need 1024-byte alignment
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_a = smem;                                   // [stages][128 rows][128 B]
    uint8_t* s_b = s_a + C::kStages * kABytes;             // [stages][BLOCK_N rows][128 B]
    uint8_t* s_land = s_b + C::kStages * C::kBBytes;       // [kLand][8 KB]
    uint32_t* s_lut = reinterpret_cast<uint32_t*>(s_land + kLand * kLandBytes);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_lut) + kLutBytes);
    uint64_t* bar_full = s_bar;                            // [stages]  16 worker warps + 1 TMA arrive
    uint64_t* bar_empty = s_bar + C::kStages;              // [stages]  1 tcgen05.commit
    uint64_t* bar_tfull = s_bar + 2 * C::kStages;          // [2]       1 tcgen05.commit
    uint64_t* bar_tempty = bar_tfull + 2;                  // [2]       16 worker warps
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_tempty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- one-time setup ---------------------------------------------------------------------------
    for (int i = tid; i < 256 * 32; i += kThreads)
        s_lut[i] = p.sfp33 ? decode_f16_bits<true>((uint32_t)(i >> 5), c_pow2frac)
                           : decode_f16_bits<false>((uint32_t)(i >> 5), c_pow2frac);
    if (warp == kTmaWarp && lane == 0) {
        ptx::prefetch_tmap(&tmap_w);
        for (int s = 0; s < C::kStages; ++s) {
            ptx::mbar_init(ptx::smem_u32(&bar_full[s]), kWorkWarps + 1);
            ptx::mbar_init(ptx::smem_u32(&bar_empty[s]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(ptx::smem_u32(&bar_tfull[b]), 1);
            ptx::mbar_init(ptx::smem_u32(&bar_tempty[b]), kWorkWarps);
        }
        ptx::fence_mbar_init();
    }
    if (warp == kMmaWarp) ptx::tmem_alloc<C::kTmemCols>(ptx::smem_u32(s_tmem));
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    const int my_tiles = ((int)blockIdx.x < p.num_tiles)
                             ? (p.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp < kWorkWarps) {
        // =========================== workers: gather + decode, then epilogue of the previous tile ==========
        const int row = tid & 127, quarter = tid >> 7;         // 16 of the K block's 64 k-values per thread
        const uint32_t a_row_off = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
        const uint32_t land_base = ptx::smem_u32(s_land);
        const uint32_t lut_base = ptx::smem_u32(s_lut);            // 128-byte aligned: (code << 7) | lane*4 never carries
        const uint32_t lane4 = (uint32_t)lane * 4u;
        const uint32_t a_base = ptx::smem_u32(s_a);
        const int total = my_tiles * p.num_kb;

        // gather cursor (runs kLand items ahead of the decode cursor)
        int g_tile_i = 0, g_kb = 0;
        int g_cb = 0, g_r = 0, g_s = 0;              // uniform-tap mode: channel block / filter row / col
        int g_tap = 0;
        long long g_tapoff = 0;                      // (r*dil_h*W + s*dil_w) * Cp, warp-uniform
        uint32_t g_mask = 0;                         // bit (r*S+s): that filter tap reads inside the image
        const uint8_t* g_xn = p.x;
        int g_hi0 = 0, g_wi0 = 0;
        bool g_rowok = false;
        auto g_setup = [&]() {
            const int tile = (int)blockIdx.x + g_tile_i * (int)gridDim.x;
            const uint32_t m = (uint32_t)(tile / p.n_tiles) * kBlockM + (uint32_t)row;
            g_rowok = m < p.M;
            const uint32_t mm = g_rowok ? m : 0u;
            const uint32_t n = p.div_hw.div(mm);
            const uint32_t rem = mm - n * p.div_hw.d;
            const uint32_t ho = p.div_w.div(rem);
            const uint32_t wo = rem - ho * p.div_w.d;
            g_hi0 = (int)ho * p.sh - p.ph;
            g_wi0 = (int)wo * p.sw - p.pw;
            g_xn = p.x + (size_t)n * p.H * p.W * p.Cp;
            g_cb = g_r = g_s = g_tap = 0;
            g_tapoff = 0;
            if (GRAN == 64) {
                // base of this thread's 16-byte slice at tap (0,0); only dereferenced where the mask allows
                g_xn += ((long long)g_hi0 * p.W + g_wi0) * p.Cp + quarter * 16;
                g_mask = 0;
                if (g_rowok && p.taps <= 32) {
                    for (int r = 0, t = 0; r < p.R; ++r) {
                        const int hi = g_hi0 + r * p.dh;
                        const bool hok = hi >= 0 && hi < p.H;
                        for (int s = 0; s < p.S; ++s, ++t) {
                            const int wi = g_wi0 + s * p.dw;
                            if (hok && wi >= 0 && wi < p.W) g_mask |= 1u << t;
                        }
                    }
                }
            }
        };
        auto g_issue = [&](int slot) {
            const uint32_t dst = land_base + (uint32_t)(slot * kLandBytes + tid * 16);
            if (GRAN == 64) {
                // Cp % 64 == 0: the whole K block is one filter tap (warp-uniform), 64 contiguous channels
                bool ok;
                if (p.taps <= 32) {
                    ok = (g_mask >> g_tap) & 1u;
                } else {
                    const int hi = g_hi0 + g_r * p.dh, wi = g_wi0 + g_s * p.dw;
                    ok = g_rowok && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
                }
                const uint8_t* src = ok ? g_xn + (g_tapoff + g_cb * 64) : p.x;
                ptx::cp_async16(dst, src, ok ? 16u : 0u);
            } else if (GRAN == 16) {
                const uint32_t q = (uint32_t)g_kb * 4u + (uint32_t)quarter;
                const uint32_t tap = p.div_cpt.div(q);
                const uint32_t c16 = q - tap * p.div_cpt.d;
                const uint32_t r = p.div_s.div(tap);
                const uint32_t s = tap - r * p.div_s.d;
                const int hi = g_hi0 + (int)r * p.dh, wi = g_wi0 + (int)s * p.dw;
                const bool ok = g_rowok && (int)tap < p.taps && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
                const uint8_t* src = ok ? g_xn + ((size_t)(hi * p.W + wi) * p.Cp + c16 * 16u) : p.x;
                ptx::cp_async16(dst, src, ok ? 16u : 0u);
            } else {  // GRAN == 4: Cp == 4 (the 3-channel stem), one tap per 32-bit word
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t tap = (uint32_t)g_kb * 16u + (uint32_t)quarter * 4u + (uint32_t)j;
                    const uint32_t r = p.div_s.div(tap);
                    const uint32_t s = tap - r * p.div_s.d;
                    const int hi = g_hi0 + (int)r * p.dh, wi = g_wi0 + (int)s * p.dw;
                    const bool ok = g_rowok && (int)tap < p.taps && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
                    const uint8_t* src = ok ? g_xn + (size_t)(hi * p.W + wi) * 4u : p.x;
                    ptx::cp_async4(dst + 4u * j, src, ok ? 4u : 0u);
                }
            }
        };
        auto g_advance = [&]() {
            if (GRAN == 64) {
                if (++g_cb == p.cblocks) {
                    g_cb = 0;
                    ++g_tap;
                    if (++g_s == p.S) { g_s = 0; ++g_r; }
                    g_tapoff = ((long long)g_r * p.dh * p.W + (long long)g_s * p.dw) * p.Cp;
                }
            }
            if (++g_kb == p.num_kb) { g_kb = 0; ++g_tile_i; if (g_tile_i < my_tiles) g_setup(); }
        };

        if (my_tiles > 0) g_setup();
        for (int i = 0; i < kLand; ++i) {
            if (i < total) { g_issue(i); g_advance(); }
            ptx::cp_async_commit();
        }
        uint32_t stage = 0, phase = 0;
        int slot = 0, it = 0;
        for (int ti = 0; ti < my_tiles; ++ti) {
            for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
                ptx::cp_async_wait<kLand - 1>();             // this thread's copy of item `it` has landed
                const uint4 cw = ptx::lds128_volatile(land_base + (uint32_t)(slot * kLandBytes + tid * 16));
                const uint32_t w[4] = {cw.x, cw.y, cw.z, cw.w};
                // code -> float16 through the per-bank table: entry (code, lane) lives in bank `lane`
                uint32_t h[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t c = w[j];
                    const uint32_t e0 = ptx::lds32_off(((c << 7) & 0x7f80u) | lane4, lut_base);
                    const uint32_t e1 = ptx::lds32_off(((c >> 1) & 0x7f80u) | lane4, lut_base);
                    const uint32_t e2 = ptx::lds32_off(((c >> 9) & 0x7f80u) | lane4, lut_base);
                    const uint32_t e3 = ptx::lds32_off(((c >> 17) & 0x7f80u) | lane4, lut_base);
                    h[2 * j] = __byte_perm(e0, e1, 0x5410);
                    h[2 * j + 1] = __byte_perm(e2, e3, 0x5410);
                }
                // refill the landing slot for item it + kLand (its previous content is in registers now)
                if (it + kLand < total) { g_issue(slot); g_advance(); }
                ptx::cp_async_commit();

                ptx::mbar_wait(ptx::smem_u32(&bar_empty[stage]), phase ^ 1u);   // MMA done with this stage
                const uint32_t a_dst = a_base + stage * kABytes + a_row_off;
                ptx::sts128(a_dst + (uint32_t)(((quarter * 2) ^ (row & 7)) << 4), h[0], h[1], h[2], h[3]);
                ptx::sts128(a_dst + (uint32_t)(((quarter * 2 + 1) ^ (row & 7)) << 4), h[4], h[5], h[6], h[7]);
                ptx::fence_proxy_async_smem();               // generic-proxy writes -> async proxy (UMMA)
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bar_full[stage]));
                if (++stage == (uint32_t)C::kStages) { stage = 0; phase ^= 1u; }
                if (++slot == kLand) slot = 0;
            }
            // epilogue of the previous tile: its MMAs were issued a whole tile ago
            if (ti > 0) {
                const int pt = ti - 1;
                const uint32_t buf = (uint32_t)pt & 1u;
                ptx::mbar_wait(ptx::smem_u32(&bar_tfull[buf]), ((uint32_t)pt >> 1) & 1u);
                ptx::tc_fence_after();
                epilogue_tile<BLOCK_N>(p, (int)blockIdx.x + pt * (int)gridDim.x, tmem_base + buf * BLOCK_N, warp, lane);
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bar_tempty[buf]));
            }
        }
        ptx::cp_async_wait<0>();
        if (my_tiles > 0) {
            const int pt = my_tiles - 1;
            const uint32_t buf = (uint32_t)pt & 1u;
            ptx::mbar_wait(ptx::smem_u32(&bar_tfull[buf]), ((uint32_t)pt >> 1) & 1u);
            ptx::tc_fence_after();
            epilogue_tile<BLOCK_N>(p, (int)blockIdx.x + pt * (int)gridDim.x, tmem_base + buf * BLOCK_N, warp, lane);
            ptx::tc_fence_before();
        }
    } else if (warp == kTmaWarp) {
        // =========================== TMA producer (B operand: weights) =============================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
                const int n0 = (tile % p.n_tiles) * BLOCK_N;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    ptx::mbar_wait_backoff(ptx::smem_u32(&bar_empty[stage]), phase ^ 1u, 64);
                    const uint32_t full = ptx::smem_u32(&bar_full[stage]);
                    ptx::mbar_arrive_expect_tx(full, (uint32_t)C::kBBytes);
                    ptx::tma_load_2d(ptx::smem_u32(s_b + stage * C::kBBytes), &tmap_w, full, kb * kBlockK, n0);
                    if (++stage == (uint32_t)C::kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else {
        // =========================== MMA issuer ===================================================
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc(0u, kBlockM, BLOCK_N);
            uint32_t stage = 0, phase = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const uint32_t buf = (uint32_t)ti & 1u;
                ptx::mbar_wait(ptx::smem_u32(&bar_tempty[buf]), (((uint32_t)ti >> 1) & 1u) ^ 1u);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * BLOCK_N;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    ptx::mbar_wait(ptx::smem_u32(&bar_full[stage]), phase);
                    ptx::tc_fence_after();
                    const uint32_t a_addr = ptx::smem_u32(s_a + stage * kABytes);
                    const uint32_t b_addr = ptx::smem_u32(s_b + stage * C::kBBytes);
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        ptx::mma_f16_ss(d_tmem, ptx::smem_desc_sw128(a_addr + k * 32), ptx::smem_desc_sw128(b_addr + k * 32),
                                        idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    ptx::mma_commit(ptx::smem_u32(&bar_empty[stage]));     // frees the smem stage
                    if (++stage == (uint32_t)C::kStages) { stage = 0; phase ^= 1u; }
                }
                ptx::mma_commit(ptx::smem_u32(&bar_tfull[buf]));            // accumulator ready
            }
        }
        __syncwarp();
    }

    // ---- teardown ---------------------------------------------------------------------------------
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
    }
}

// ---- host side -----------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
        else
            cudaGetLastError();
    }
    return fn;
}

template <int BLOCK_N, int GRAN>
static int launch_igemm(const CUtensorMap& tmap, const IgemmParams& p, cudaStream_t st) {
    using C = Cfg<BLOCK_N>;
    auto kern = conv_igemm_kernel<BLOCK_N, GRAN>;
    static DeviceOnce attr_once;
    bool& attr_done = attr_once.flag();
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
        if (e != cudaSuccess) return set_error((int)e, "conv_igemm: smem attribute (%d B): %s", C::kSmemBytes, cudaGetErrorString(e));
        attr_done = true;
    }
    const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
    kern<<<grid, kThreads, C::kSmemBytes, st>>>(tmap, p);
    return check_launch("conv_igemm_kernel");
}

bool conv2d_fwd_dense_v2_supported(const SlfpConvDesc* d);
int conv2d_fwd_dense_v2(const SlfpConvDesc* d, const uint8_t* x_codes, const void* w_f16, const SlfpEpilogue* epi, cudaStream_t st);

bool conv2d_fwd_stem_direct_supported(const SlfpConvDesc* d, const SlfpEpilogue* e);
int conv2d_fwd_stem_direct(const SlfpConvDesc* d, const uint8_t* x_codes, const void* w_f16, const SlfpEpilogue* e, cudaStream_t st);

int conv2d_fwd_dense(const SlfpConvDesc* d, const uint8_t* x_codes, const void* w_f16, const SlfpEpilogue* epi,
                     cudaStream_t st) {
    // c_phys % 16 == 0: the warp-specialised TMA-im2col kernel (conv_igemm_v2.cu); this file keeps the
    // 4-channel network-input layout (the 7x7 / 3x3 stems), whose taps are narrower than a TMA box row.
    if (conv2d_fwd_dense_v2_supported(d)) return conv2d_fwd_dense_v2(d, x_codes, w_f16, epi, st);
    // 3x3 RGB stems of the fused pipeline: CUDA-core direct convolution (conv_stem_direct.cu)
    if (conv2d_fwd_stem_direct_supported(d, epi)) return conv2d_fwd_stem_direct(d, x_codes, w_f16, epi, st);
    if (d->pad_h_extra || d->pad_w_extra)
        return set_error(SLFP_ERR_UNSUPPORTED, "conv2d_fwd: asymmetric padding needs c_phys %% 16 == 0");
    if (epi->y_codes && epi->next_fmt != SLFP_FMT_SLFP34_ACT && epi->next_fmt != SLFP_FMT_SFP33)
        return set_error(SLFP_ERR_UNSUPPORTED, "conv2d_fwd: the 4-channel-input kernel writes signed code formats only");
    if (d->c_phys != 4 && (d->c_phys % 16) != 0)
        return set_error(SLFP_ERR_BAD_ARG, "conv2d_fwd: c_phys=%d must be 4 or a multiple of 16", d->c_phys);
    if (d->fmt != SLFP_FMT_SLFP34_ACT && d->fmt != SLFP_FMT_SFP33)
        return set_error(SLFP_ERR_BAD_ARG, "conv2d_fwd: activation code format %d", d->fmt);
    if (epi->layerout) return set_error(SLFP_ERR_UNSUPPORTED, "conv2d_fwd: layerout epilogue needs c_phys %% 16 == 0 (the 4-channel stem kernel has none)");
    if (epi->y_codes && (epi->k_phys_out % 16 != 0 || epi->k_phys_out < d->k))
        return set_error(SLFP_ERR_BAD_ARG, "conv2d_fwd: k_phys_out=%d", epi->k_phys_out);
    if ((((uintptr_t)x_codes | (uintptr_t)w_f16 | (uintptr_t)epi->y_f32 | (uintptr_t)epi->y_f16 |
          (uintptr_t)epi->y_codes | (uintptr_t)epi->y_codes2) & 15u) != 0)
        return set_error(SLFP_ERR_BAD_ARG, "conv2d_fwd: tensors must be 16-byte aligned");
    IgemmParams p;
    p.x = x_codes;
    p.H = d->h; p.W = d->w; p.Cp = d->c_phys;
    p.R = d->r; p.S = d->s; p.sh = d->stride_h; p.sw = d->stride_w; p.ph = d->pad_h; p.pw = d->pad_w;
    p.dh = d->dil_h; p.dw = d->dil_w;
    p.Ho = (d->h + 2 * d->pad_h - d->dil_h * (d->r - 1) - 1) / d->stride_h + 1;
    p.Wo = (d->w + 2 * d->pad_w - d->dil_w * (d->s - 1) - 1) / d->stride_w + 1;
    if (p.Ho <= 0 || p.Wo <= 0 || d->n <= 0) return set_error(SLFP_ERR_BAD_ARG, "conv2d_fwd: empty output");
    const unsigned long long M64 = (unsigned long long)d->n * p.Ho * p.Wo;
    if (M64 >= (1ull << 31)) return set_error(SLFP_ERR_UNSUPPORTED, "conv2d_fwd: more than 2^31 output pixels");
    p.M = (uint32_t)M64;
    p.Kout = d->k;
    p.taps = d->r * d->s;
    const size_t pitch = slfp_conv_wpitch(d);
    p.num_kb = (int)(pitch / kBlockK);
    const int bn = d->k > 128 ? 256 : (d->k > 64 ? 128 : 64);
    p.m_tiles = (int)((p.M + kBlockM - 1) / kBlockM);
    p.n_tiles = (d->k + bn - 1) / bn;
    p.num_tiles = p.m_tiles * p.n_tiles;
    p.div_hw = make_fastdiv((uint32_t)(p.Ho * p.Wo));
    p.div_w = make_fastdiv((uint32_t)p.Wo);
    p.div_cpt = make_fastdiv((uint32_t)(d->c_phys >= 16 ? d->c_phys / 16 : 1));
    p.div_s = make_fastdiv((uint32_t)d->s);
    p.cblocks = (d->c_phys % 64 == 0) ? d->c_phys / 64 : 0;
    p.epi = *epi;
    p.next_div = make_divk(epi->y_codes ? epi->next_k_div : 1.0f);
    p.next_div2 = make_divk(epi->y_codes2 ? epi->next_k_div2 : 1.0f);

    auto enc = get_encode_fn();
    if (!enc) return set_error(SLFP_ERR_DRIVER, "conv2d_fwd: cuTensorMapEncodeTiled not available");
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)pitch, (cuuint64_t)d->k};
    const cuuint64_t gstr[1] = {(cuuint64_t)pitch * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)bn};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w_f16), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return set_error(SLFP_ERR_DRIVER, "conv2d_fwd: cuTensorMapEncodeTiled failed (%d)", (int)cr);

    p.sfp33 = d->fmt == SLFP_FMT_SFP33 ? 1 : 0;
    const int gran = d->c_phys == 4 ? 4 : (p.cblocks ? 64 : 16);
#define SLFP_IGEMM_CASE(BN)                                             \
    if (bn == BN) {                                                     \
        if (gran == 64) return launch_igemm<BN, 64>(tmap, p, st);       \
        if (gran == 16) return launch_igemm<BN, 16>(tmap, p, st);       \
        return launch_igemm<BN, 4>(tmap, p, st);                        \
    }
    SLFP_IGEMM_CASE(64)
    SLFP_IGEMM_CASE(128)
    SLFP_IGEMM_CASE(256)
#undef SLFP_IGEMM_CASE
    return set_error(SLFP_ERR_UNSUPPORTED, "conv2d_fwd: no tile for k=%d", d->k);
}

}  // namespace slfp
