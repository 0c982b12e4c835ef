// conv_bwd_sm100.cu -- backward of Conv2d_Q (utils/conv2d_func.py:20-25 under autograd, identity STE of
// utils/sfp_quant.py:50-53) as two implicit GEMMs on tcgen05 tensor cores:
//
//   dgrad  dx[n,h,w,c]   = Kw * sum_{r,s,k} gy[n,(h+p-r)/st,(w+p-s)/st,k] * w_q[k,r,s,c]
//   wgrad  dw[k,c,r,s]   = Ka * sum_{n,ho,wo} gy[n,ho,wo,k] * x_q[n,ho*st+r-p,wo*st+s-p,c]
//
// Operands are float16 (kind::f16, float32 accumulation in tensor memory):
//   G  = gy * 2^e   rounded to float16, 2^e chosen from max|gy| so the largest element sits at 2^14 (gradients are
//        far below the float16 normal range otherwise); the epilogue multiplies 2^-e back in;
//   Xh = float16 image of the saved 8-bit activation codes, NHWC;   Wt = float16 image of the saved weight codes,
//        re-laid out [c][tap][k] so the dgrad reduction dimension is contiguous.
// One persistent kernel, six warps: TMA producer / MMA issuer / four epilogue warps (one per TMEM lane quadrant).
//   dgrad: A = im2col(G) by TMA im2col boxes (128 pixels x 64 k, K-major, 128B swizzle), B = Wt tiles; a strided conv
//          is split into its stride_h*stride_w output-parity classes, each a stride-1 convolution with a subset
//          of the taps whose rows the epilogue scatters to (st*i + a, st*j + b): no zero-dilated tensor, no wasted MACs.
//   wgrad: the reduction runs over output pixels, which is the SLOW dimension of both NHWC operands, so both are
//          MN-major shared-memory operands: A = 64-pixel x 128-k boxes of G, B = im2col boxes of Xh (64 pixels x 64 c)
//          per filter tap; the pixel range is split across CTAs and partial tiles are added with red.global.v4.f32 into
//          a zeroed [k][tap][c] float32 accumulator that a last kernel scales into the caller's (strided) dw.
#include "slfp_common.cuh"
#include "sm100_ptx.cuh"

#include <cudaTypedefs.h>

#include <algorithm>
#include <vector>

namespace slfp {
namespace bwd {

constexpr int kBM = 128, kBK = 64, kMaxStages = 4;
// wgrad stages hold 128 reduction pixels (dgrad: 64 k elements = one 128-byte row): every TMA load costs >= ~190 cycles
// whatever its size (tools/ubench/tma_box.cu), and a wgrad stage needs six of them (2 gy boxes + up to 4 im2col boxes) -
// with 64 pixels per stage that was 1 140 TMA cycles against 512 MMA cycles.
constexpr int kWgradPix = 128;
constexpr int kMaxTaps = 32;
constexpr int kThreads = 192;

struct Params {
    uint32_t M;                  // dgrad: output pixels of the class; wgrad: output pixels (reduction length)
    int HiWj, Wj;                // pixel index -> (n, i, j)
    int n_tiles, num_items;
    int num_kb, kchunks, ntaps;
    int org_w, org_h, sw, sh;    // window origin of pixel (i, j): (j*sw + org_w, i*sh + org_h)
    uint16_t off_w[kMaxTaps], off_h[kMaxTaps];
    // dgrad epilogue
    float* out;
    int C, H, W, osh, osw, oa, ob, vec_ok;
    // wgrad
    int units, cchunks, ugroups, splits, pb_total, pb_per_split;
    float* wacc;
    int K, Cp;
    const float* scale;          // device: max|gy| (grad_scale() derives 2^e and 2^-e from it)
    float post;
};

template <int BN, int MODE> struct Cfg {
    static constexpr int kKB = MODE == 1 ? kWgradPix : kBK;                  // reduction elements per stage
    static constexpr int kStages = MODE == 1 ? (BN > 128 ? 2 : (BN > 64 ? 3 : 4)) : 4;
    static constexpr int kABytes = kBM * kKB * 2;
    static constexpr int kBBytes = BN * kKB * 2;
    static constexpr int kBoxBytes = 64 * kKB * 2;                           // one 64-wide MN-major box (wgrad)
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kTmemCols = BN <= 64 ? 128 : (BN <= 128 ? 256 : 512);
    static constexpr int kBufCols = kTmemCols / 2;
    static constexpr int kSmemBytes = kStages * kStageBytes + 256 + 4 * 4096;   // + barriers + epilogue staging
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// 2^e with max|gy| * 2^e in [2^14, 2^15) and 2^-e, from the abs-max word the reduction kernel left in the workspace;
// (1, 1) when the maximum is 0 / Inf / NaN.  Every kernel that needs the scale recomputes it (two instructions)
// instead of a separate one-thread launch per layer.
__device__ __forceinline__ float2 grad_scale(const float* __restrict__ amax) {
    const float m = __ldg(amax);
    float s = 1.f, inv = 1.f;
    if (m > 0.f && m < INFINITY) {
        int e = 14 - (int)((__float_as_uint(m) >> 23) & 0xffu) + 127;       // exponent that moves m into [2^14, 2^15)
        e = e > 120 ? 120 : (e < -120 ? -120 : e);
        s = __uint_as_float((uint32_t)(e + 127) << 23);
        inv = __uint_as_float((uint32_t)(127 - e) << 23);
    }
    return make_float2(s, inv);
}

// MODE 0: dgrad (both operands K-major), MODE 1: wgrad (both operands MN-major)
template <int BN, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
conv_bwd_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const Params p) {
    using C = Cfg<BN, MODE>;
    constexpr int kStages = C::kStages, kABytes = C::kABytes;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + kStages * C::kStageBytes);
    uint64_t* bar_full = s_bar;                 // [kStages] 1 arrive.expect_tx
    uint64_t* bar_empty = bar_full + kMaxStages;   // [kStages] 1 tcgen05.commit
    uint64_t* bar_tfull = bar_empty + kMaxStages;  // [2]
    uint64_t* bar_tempty = bar_tfull + 2;       // [2] 4 epilogue warps
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_tempty + 2);

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_b);
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(ptx::smem_u32(&bar_full[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&bar_empty[s]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(ptx::smem_u32(&bar_tfull[b]), 1);
            ptx::mbar_init(ptx::smem_u32(&bar_tempty[b]), 4);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc<C::kTmemCols>(ptx::smem_u32(s_tmem));
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    const int my_items = ((int)blockIdx.x < p.num_items) ? (p.num_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    // K blocks of a work item.  dgrad: every tile runs all taps x k chunks.  wgrad: the split's range of 64-pixel blocks.
    auto item_kb = [&](int item, int& kb0) -> int {
        if (MODE == 0) { kb0 = 0; return p.num_kb; }
        const int split = item % p.splits;
        kb0 = split * p.pb_per_split;
        const int kb1 = min(p.pb_total, kb0 + p.pb_per_split);
        return kb1 - kb0;
    };

    if (warp == 0) {
        // =========================== TMA producer (converged warp, one elected lane issues) ===============
        {
            uint32_t stage = 0, phase = 0;
            for (int it = 0; it < my_items; ++it) {
                const int item = (int)blockIdx.x + it * (int)gridDim.x;
                int kb0;
                const int nkb = item_kb(item, kb0);
                if (MODE == 0) {
                    const uint32_t m0 = (uint32_t)(item / p.n_tiles) * kBM;
                    const int n0 = (item % p.n_tiles) * BN;
                    const int n = (int)(m0 / (uint32_t)p.HiWj);
                    const int rem = (int)(m0 - (uint32_t)n * (uint32_t)p.HiWj);
                    const int i = rem / p.Wj, j = rem - i * p.Wj;
                    const int w0 = j + p.org_w, h0 = i + p.org_h;
                    int t = 0, kc = 0;
                    for (int kb = 0; kb < nkb; ++kb) {
                        ptx::mbar_wait(ptx::smem_u32(&bar_empty[stage]), phase ^ 1u, 1u | ((uint32_t)kb << 8));
                        const uint32_t full = ptx::smem_u32(&bar_full[stage]);
                        const uint32_t sa = ptx::smem_u32(smem + stage * C::kStageBytes);
                        if (ptx::elect_one()) {
                            ptx::mbar_arrive_expect_tx(full, (uint32_t)C::kStageBytes);
                            ptx::tma_load_im2col_4d(sa, &tmap_a, full, kc * 64, w0, h0, n, p.off_w[t], p.off_h[t]);
                            ptx::tma_load_2d(sa + kABytes, &tmap_b, full, kb * kBK, n0);
                        }
                        __syncwarp();
                        if (++kc == p.kchunks) { kc = 0; ++t; }
                        if (++stage == (uint32_t)kStages) { stage = 0; phase ^= 1u; }
                    }
                } else {
                    const int tile = item / p.splits;
                    const int ko0 = (tile / p.ugroups) * kBM;
                    const int u0 = (tile % p.ugroups) * (BN / 64);
                    const int nu = min(BN / 64, p.units - u0);
                    for (int kb = 0; kb < nkb; ++kb) {
                        const uint32_t pix0 = (uint32_t)(kb0 + kb) * (uint32_t)C::kKB;
                        const int n = (int)(pix0 / (uint32_t)p.HiWj);
                        const int rem = (int)(pix0 - (uint32_t)n * (uint32_t)p.HiWj);
                        const int ho = rem / p.Wj, wo = rem - ho * p.Wj;
                        const int w0 = wo * p.sw + p.org_w, h0 = ho * p.sh + p.org_h;
                        ptx::mbar_wait(ptx::smem_u32(&bar_empty[stage]), phase ^ 1u, 1u | ((uint32_t)kb << 8));
                        const uint32_t full = ptx::smem_u32(&bar_full[stage]);
                        const uint32_t sa = ptx::smem_u32(smem + stage * C::kStageBytes);
                        if (ptx::elect_one()) {
                            ptx::mbar_arrive_expect_tx(full, (uint32_t)(kABytes + nu * C::kBoxBytes));
                            ptx::tma_load_2d(sa, &tmap_a, full, ko0, (int)pix0);
                            ptx::tma_load_2d(sa + C::kBoxBytes, &tmap_a, full, ko0 + 64, (int)pix0);
                        }
                        __syncwarp();
                        for (int u = 0; u < nu; ++u) {
                            const int unit = u0 + u, tap = unit / p.cchunks, cc = unit - tap * p.cchunks;
                            if (ptx::elect_one())
                                ptx::tma_load_im2col_4d(sa + kABytes + u * C::kBoxBytes, &tmap_b, full, cc * 64, w0, h0, n, p.off_w[tap], p.off_h[tap]);
                            __syncwarp();
                        }
                        if (++stage == (uint32_t)kStages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // =========================== MMA issuer (converged warp, one elected lane) ========================
        constexpr uint32_t idesc = ptx::make_idesc(0u, kBM, BN) | (MODE == 1 ? ((1u << 15) | (1u << 16)) : 0u);
        constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);     // SBO 1024 B, version 1, SWIZZLE_128B
        constexpr uint32_t kLbo = MODE == 1 ? ((uint32_t)C::kBoxBytes >> 4) : 1u;  // MN-major: next 64-element atom along M / N
        constexpr uint32_t kStep = MODE == 1 ? (2048u >> 4) : 2u;                // start-address advance per K = 16
        const uint32_t base_lo = ((ptx::smem_u32(smem) >> 4) & 0x3fffu) | (kLbo << 16);
        uint32_t stage = 0, phase = 0;
        for (int it = 0; it < my_items; ++it) {
            const int item = (int)blockIdx.x + it * (int)gridDim.x;
            int kb0;
            const int nkb = item_kb(item, kb0);
            const uint32_t buf = (uint32_t)it & 1u;
            ptx::mbar_wait(ptx::smem_u32(&bar_tempty[buf]), (((uint32_t)it >> 1) & 1u) ^ 1u, 3u);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * (uint32_t)C::kBufCols;
            for (int kb = 0; kb < nkb; ++kb) {
                ptx::mbar_wait(ptx::smem_u32(&bar_full[stage]), phase, 4u | ((uint32_t)kb << 8));
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    const uint32_t a_lo = base_lo + stage * (uint32_t)(C::kStageBytes >> 4);
                    const uint32_t b_lo = a_lo + (uint32_t)(kABytes >> 4);
#pragma unroll
                    for (int k = 0; k < C::kKB / 16; ++k)
                        ptx::mma_f16_ss(d_tmem, ((uint64_t)kDescHi << 32) | (uint64_t)(a_lo + kStep * k),
                                        ((uint64_t)kDescHi << 32) | (uint64_t)(b_lo + kStep * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    ptx::mma_commit(ptx::smem_u32(&bar_empty[stage]));
                    if (kb == nkb - 1) ptx::mma_commit(ptx::smem_u32(&bar_tfull[buf]));
                }
                __syncwarp();
                if (++stage == (uint32_t)kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        // =========================== epilogue: TMEM -> global ============================================
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const float inv = grad_scale(p.scale).y;
        uint8_t* stg = smem + kStages * C::kStageBytes + 256 + (warp & 3) * 4096;      // 32 rows x 128 B per warp
        const uint32_t stg_u = ptx::smem_u32(stg);
        for (int it = 0; it < my_items; ++it) {
            const int item = (int)blockIdx.x + it * (int)gridDim.x;
            const uint32_t buf = (uint32_t)it & 1u;
            ptx::mbar_wait_backoff(ptx::smem_u32(&bar_tfull[buf]), ((uint32_t)it >> 1) & 1u, 64, 5u);
            ptx::tc_fence_after();
            const uint32_t t_row = tmem_base + buf * (uint32_t)C::kBufCols + ((uint32_t)(quad * 32) << 16);
            if (MODE == 0) {
                const float sc = p.post * inv;
                const uint32_t m = (uint32_t)(item / p.n_tiles) * kBM + (uint32_t)row;
                const int n0 = (item % p.n_tiles) * BN;
                const bool valid = m < p.M;
                const int n = (int)(m / (uint32_t)p.HiWj);
                const int rem = (int)(m - (uint32_t)n * (uint32_t)p.HiWj);
                const int i = rem / p.Wj, j = rem - i * p.Wj;
                float* dst = p.out + (((size_t)n * p.H + (size_t)(i * p.osh + p.oa)) * p.W + (size_t)(j * p.osw + p.ob)) * p.C + n0;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    if (n0 + c0 >= p.C) break;
                    uint32_t v[32];
                    ptx::tmem_ld32(t_row + (uint32_t)c0, v);
                    ptx::tmem_ld_wait();
                    if (p.vec_ok) {
                        // A thread owns one pixel row, so direct stores would touch 32 rows per instruction (half-filled
                        // sectors: 2.7 TB/s measured).  Stage the warp's 32 x 32 block in shared memory (16-byte chunks
                        // XOR-swizzled by row: conflict-free both ways) and write it back 4 rows per instruction, 8 lanes
                        // covering one 128-byte row piece.
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            ptx::sts128(stg_u + (uint32_t)(lane * 128 + ((q ^ (lane & 7)) << 4)),
                                        __float_as_uint(__uint_as_float(v[q * 4]) * sc), __float_as_uint(__uint_as_float(v[q * 4 + 1]) * sc),
                                        __float_as_uint(__uint_as_float(v[q * 4 + 2]) * sc), __float_as_uint(__uint_as_float(v[q * 4 + 3]) * sc));
                        __syncwarp();
                        const unsigned long long my_row = valid ? (unsigned long long)(uintptr_t)dst : 0ull;
                        const int cch = lane & 7;
                        const bool col_ok = n0 + c0 + cch * 4 + 4 <= p.C;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int r = k * 4 + (lane >> 3);
                            const unsigned long long rp = __shfl_sync(0xffffffffu, my_row, r);
                            const float4 val = ptx::lds128_f4(stg_u + (uint32_t)(r * 128 + ((cch ^ (r & 7)) << 4)));
                            if (rp != 0ull && col_ok) *reinterpret_cast<float4*>(reinterpret_cast<float*>(rp) + c0 + cch * 4) = val;
                        }
                        __syncwarp();
                    } else if (valid) {
#pragma unroll
                        for (int q = 0; q < 32; ++q)
                            if (n0 + c0 + q < p.C) dst[c0 + q] = __uint_as_float(v[q]) * sc;
                    }
                }
            } else {
                const int tile = item / p.splits;
                const int ko = (tile / p.ugroups) * kBM + row;
                const int u0 = (tile % p.ugroups) * (BN / 64);
                const int nu = min(BN / 64, p.units - u0);
                for (int u = 0; u < nu; ++u) {
                    const int unit = u0 + u, tap = unit / p.cchunks, cc = unit - tap * p.cchunks;
                    float* dst = p.wacc + ((size_t)ko * p.ntaps + tap) * p.Cp + cc * 64;
#pragma unroll 1
                    for (int h = 0; h < 2; ++h) {
                        uint32_t v[32];
                        ptx::tmem_ld32(t_row + (uint32_t)(u * 64 + h * 32), v);
                        ptx::tmem_ld_wait();
                        if (ko < p.K) {
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                red_add_v4(dst + h * 32 + q * 4, __uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]),
                                           __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
                        }
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bar_tempty[buf]));
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
    }
}

// ---- operand preparation (HBM-bound element-wise passes) ------------------------------------------------------
// gy [rows][K] float32 -> G [rows][Kp] float16 (scaled, zero-padded columns); 8 outputs per thread
__global__ void __launch_bounds__(256) grad_to_f16_kernel(const float* __restrict__ gy, size_t rows, int K, int Kp,
                                                          const float* __restrict__ scale, __half* __restrict__ g) {
    const float s = grad_scale(scale).x;
    const int k8 = Kp >> 3;
    const size_t total = rows * (size_t)k8;
    const bool vec = (K & 3) == 0;
    for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256) {
        const size_t r = idx / k8;
        const int k0 = (int)(idx - r * k8) * 8;
        const float* src = gy + r * K + k0;
        float v[8];
        if (vec && k0 + 8 <= K) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (k0 + i < K) ? __ldg(src + i) : 0.f;
        }
        uint4 o;
        __half2 h;
        h = __floats2half2_rn(v[0] * s, v[1] * s); o.x = *reinterpret_cast<uint32_t*>(&h);
        h = __floats2half2_rn(v[2] * s, v[3] * s); o.y = *reinterpret_cast<uint32_t*>(&h);
        h = __floats2half2_rn(v[4] * s, v[5] * s); o.z = *reinterpret_cast<uint32_t*>(&h);
        h = __floats2half2_rn(v[6] * s, v[7] * s); o.w = *reinterpret_cast<uint32_t*>(&h);
        *reinterpret_cast<uint4*>(g + r * Kp + k0) = o;
    }
}

// activation codes -> float16 (16 codes per thread, table in shared memory)
__global__ void __launch_bounds__(256) codes_to_f16_kernel(const uint8_t* __restrict__ codes, size_t n16, int fmt,
                                                           __half* __restrict__ out) {
    __shared__ uint16_t s_tab[256];
    s_tab[threadIdx.x] = __half_as_ushort(__float2half_rn(decode_act_any(threadIdx.x, fmt, c_pow2frac)));
    __syncthreads();
    for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < n16; idx += (size_t)gridDim.x * 256) {
        const uint4 c = __ldg(reinterpret_cast<const uint4*>(codes) + idx);
        const uint32_t w[4] = {c.x, c.y, c.z, c.w};
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            o[2 * i] = (uint32_t)s_tab[w[i] & 0xffu] | ((uint32_t)s_tab[(w[i] >> 8) & 0xffu] << 16);
            o[2 * i + 1] = (uint32_t)s_tab[(w[i] >> 16) & 0xffu] | ((uint32_t)s_tab[w[i] >> 24] << 16);
        }
        uint4* dst = reinterpret_cast<uint4*>(out) + idx * 2;
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
}

// weight codes KRSC [K][pitch] -> Wt [Crows][T][Kp] float16 for one parity class (taps listed in tapidx)
struct TapList { int n; int idx[kMaxTaps]; };
__global__ void __launch_bounds__(256) wt_prep_kernel(const uint8_t* __restrict__ wc, size_t pitch, int Cw, int K, int Kp,
                                                      int Crows, int wfmt, TapList tl, __half* __restrict__ wt) {
    __shared__ uint16_t s_tab[256];
    s_tab[threadIdx.x] = __half_as_ushort(__float2half_rn(decode_act_any(threadIdx.x, wfmt == SLFP_FMT_SFP33 ? SLFP_FMT_SFP33 : SLFP_FMT_SLFP34_ACT, c_pow2frac)));
    __syncthreads();
    const size_t total = (size_t)Crows * tl.n * Kp;
    for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256) {
        const int k = (int)(idx % Kp);
        const size_t ct = idx / Kp;
        const int t = (int)(ct % tl.n);
        const int c = (int)(ct / tl.n);
        uint16_t v = 0;
        if (k < K && c < Cw) v = s_tab[wc[(size_t)k * pitch + (size_t)tl.idx[t] * Cw + c]];
        reinterpret_cast<uint16_t*>(wt)[idx] = v;
    }
}

// wacc [K][taps][Cp] float32 -> dw (strided OIHW view) * Ka * 2^-e
__global__ void __launch_bounds__(256) wgrad_finalize_kernel(const float* __restrict__ wacc, int K, int C, int Cp, int R, int S,
                                                             const float* __restrict__ scale, float post, float* __restrict__ dw,
                                                             long long so, long long sc, long long sr, long long ss) {
    const float f = post * grad_scale(scale).y;
    const size_t total = (size_t)K * C * R * S;
    for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256) {
        // idx enumerates the destination in (k, c, r, s) order (coalesced writes for a contiguous OIHW tensor)
        const int s = (int)(idx % S);
        const int r = (int)((idx / S) % R);
        const int c = (int)((idx / ((size_t)S * R)) % C);
        const int k = (int)(idx / ((size_t)S * R * C));
        dw[k * so + c * sc + r * sr + s * ss] = wacc[((size_t)k * R * S + r * S + s) * Cp + c] * f;
    }
}

template <typename PFN>
static PFN driver_fn(const char* name) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint(name, &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
        return reinterpret_cast<PFN>(ptr);
    cudaGetLastError();
    return nullptr;
}

template <int BN, int MODE>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, cudaStream_t st) {
    using C = Cfg<BN, MODE>;
    auto kern = conv_bwd_kernel<BN, MODE>;
    static DeviceOnce attr_once;
    bool& attr_done = attr_once.flag();
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
        if (e != cudaSuccess) return set_error((int)e, "conv_bwd: smem attribute (%d B): %s", C::kSmemBytes, cudaGetErrorString(e));
        attr_done = true;
    }
    const int grid = p.num_items < num_sms() ? p.num_items : num_sms();
    kern<<<grid, kThreads, C::kSmemBytes, st>>>(ta, tb, p);
    return check_launch(MODE == 0 ? "conv_bwd_kernel<dgrad>" : "conv_bwd_kernel<wgrad>");
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

struct Layout {
    size_t off_scale, off_g, off_x, off_wt, off_wacc, total;
    int Ho, Wo, Kp;
    bool dx_ok, dw_ok;
};

static Layout plan(const SlfpConvDesc* d, bool need_dx, bool need_dw) {
    Layout L;
    memset(&L, 0, sizeof(L));
    L.Ho = (d->h + 2 * d->pad_h - d->dil_h * (d->r - 1) - 1) / d->stride_h + 1;
    L.Wo = (d->w + 2 * d->pad_w - d->dil_w * (d->s - 1) - 1) / d->stride_w + 1;
    L.Kp = (d->k + 63) / 64 * 64;
    const bool common = d->groups == 1 && !d->pad_h_extra && !d->pad_w_extra && L.Ho > 0 && L.Wo > 0 &&
                        (d->fmt == SLFP_FMT_SFP33 || d->fmt == SLFP_FMT_SLFP34_ACT) &&
                        (unsigned long long)d->n * L.Ho * L.Wo < (1ull << 31) && (unsigned long long)d->n * d->h * d->w < (1ull << 31) &&
                        d->pad_h < 120 && d->pad_w < 120 && (d->r - 1) * d->dil_h < 250 && (d->s - 1) * d->dil_w < 250 &&
                        d->stride_h <= 8 && d->stride_w <= 8 && getenv("SLFP_BWD_DIRECT") == nullptr;
    L.dx_ok = common && need_dx && d->r * d->s <= kMaxTaps;
    L.dw_ok = common && need_dw && d->c_phys % 64 == 0 && d->r * d->s <= kMaxTaps;
    size_t off = 0;
    L.off_scale = off; off += 256;
    const size_t mout = (size_t)d->n * L.Ho * L.Wo;
    L.off_g = off; off += align_up(mout * L.Kp * 2 + 64 * L.Kp * 2, 256);
    if (L.dw_ok) {
        L.off_x = off; off += align_up((size_t)d->n * d->h * d->w * d->c_phys * 2, 256);
        L.off_wacc = off; off += align_up((size_t)d->k * d->r * d->s * d->c_phys * 4, 256);
    }
    if (L.dx_ok) { L.off_wt = off; off += align_up((size_t)d->c_phys * d->r * d->s * L.Kp * 2, 256); }
    L.total = (L.dx_ok || L.dw_ok) ? off : 0;
    return L;
}

}  // namespace bwd

int conv2d_bwd_direct(const SlfpConvDesc* d, const float* gy, const uint8_t* x_codes, const uint8_t* w_codes, int wfmt,
                      float ka, float kw, float* dx, float* dwt, long long so, long long sc, long long sr, long long ss,
                      float* db, cudaStream_t st);

size_t conv2d_bwd_tc_workspace(const SlfpConvDesc* d, int need_dx, int need_dw) {
    return bwd::plan(d, need_dx != 0, need_dw != 0).total;
}

int conv2d_bwd_tc(const SlfpConvDesc* d, const float* gy, const uint8_t* x_codes, const uint8_t* w_codes, int wfmt,
                  float ka, float kw, float* dx, float* dwt, long long so, long long sc, long long sr, long long ss,
                  float* db, void* workspace, size_t ws_bytes, cudaStream_t st, const float* gy_absmax) {
    using namespace bwd;
    const Layout L = plan(d, dx != nullptr, dwt != nullptr);
    if (L.total == 0 || !workspace || ws_bytes < L.total || (((uintptr_t)workspace) & 255u))
        return conv2d_bwd_direct(d, gy, x_codes, w_codes, wfmt, ka, kw, dx, dwt, so, sc, sr, ss, db, st);
    static auto enc_tiled = driver_fn<PFN_cuTensorMapEncodeTiled_v12000>("cuTensorMapEncodeTiled");
    static auto enc_im2col = driver_fn<PFN_cuTensorMapEncodeIm2col_v12000>("cuTensorMapEncodeIm2col");
    if (!enc_tiled || !enc_im2col) return set_error(SLFP_ERR_DRIVER, "conv2d_bwd: cuTensorMapEncode{Tiled,Im2col} not available");

    uint8_t* ws = (uint8_t*)workspace;
    float* scale = (float*)(ws + L.off_scale);
    __half* G = (__half*)(ws + L.off_g);
    const int Ho = L.Ho, Wo = L.Wo, Kp = L.Kp;
    const size_t mout = (size_t)d->n * Ho * Wo;
    int rc;
    // G = float16(gy * 2^e)
    if (gy_absmax) {                                         // max |gy| already known on the device (the kernel that wrote gy tracked it)
        cudaError_t e = cudaMemcpyAsync(scale, gy_absmax, sizeof(float), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return set_error((int)e, "conv2d_bwd: copy of the gradient abs-max: %s", cudaGetErrorString(e));
    } else if ((rc = slfp_absmax_f32(gy, mout * d->k, scale, 1, (slfp_stream_t)st))) return rc;
    {
        const size_t tot = mout * (Kp / 8);
        const int grid = (int)std::min<size_t>((size_t)num_sms() * 16, ceil_div_sz(tot, 256));
        grad_to_f16_kernel<<<grid, 256, 0, st>>>(gy, mout, d->k, Kp, scale, G);
        if ((rc = check_launch("grad_to_f16_kernel"))) return rc;
    }

    // ---------------- dgrad ----------------
    if (dx && !L.dx_ok) {
        if ((rc = conv2d_bwd_direct(d, gy, x_codes, w_codes, wfmt, ka, kw, dx, nullptr, 0, 0, 0, 0, nullptr, st))) return rc;
    } else if (dx) {
        if (!w_codes) return set_error(SLFP_ERR_BAD_ARG, "conv2d_bwd: dx needs w_codes");
        const size_t pitch = slfp_conv_wpitch(d);
        __half* Wt = (__half*)(ws + L.off_wt);
        bool need_zero = false;
        struct Cls { int a, b, Hi, Wj, T, dminh, dminw; TapList tl; int dh[kMaxTaps], dw[kMaxTaps]; };
        std::vector<Cls> classes;
        for (int a = 0; a < d->stride_h; ++a)
            for (int b = 0; b < d->stride_w; ++b) {
                Cls c;
                c.a = a; c.b = b;
                c.Hi = (d->h - a + d->stride_h - 1) / d->stride_h;
                c.Wj = (d->w - b + d->stride_w - 1) / d->stride_w;
                if (c.Hi <= 0 || c.Wj <= 0) continue;
                c.T = 0; c.dminh = 1 << 30; c.dminw = 1 << 30;
                for (int r = 0; r < d->r; ++r) {
                    const int th = a + d->pad_h - r * d->dil_h;
                    if (((th % d->stride_h) + d->stride_h) % d->stride_h) continue;
                    for (int s = 0; s < d->s; ++s) {
                        const int tw = b + d->pad_w - s * d->dil_w;
                        if (((tw % d->stride_w) + d->stride_w) % d->stride_w) continue;
                        c.tl.idx[c.T] = r * d->s + s;
                        c.dh[c.T] = floor_div(th, d->stride_h);
                        c.dw[c.T] = floor_div(tw, d->stride_w);
                        c.dminh = std::min(c.dminh, c.dh[c.T]);
                        c.dminw = std::min(c.dminw, c.dw[c.T]);
                        ++c.T;
                    }
                }
                c.tl.n = c.T;
                if (c.T == 0) { need_zero = true; continue; }
                classes.push_back(c);
            }
        if (need_zero) cudaMemsetAsync(dx, 0, (size_t)d->n * d->h * d->w * d->c * sizeof(float), st);
        size_t wt_off = 0;
        for (const Cls& c : classes) {
            __half* wt = Wt + wt_off;
            wt_off += (size_t)d->c_phys * c.T * Kp;
            {
                const size_t tot = (size_t)d->c_phys * c.T * Kp;
                const int grid = (int)std::min<size_t>((size_t)num_sms() * 8, ceil_div_sz(tot, 256));
                wt_prep_kernel<<<grid, 256, 0, st>>>(w_codes, pitch, d->c_phys, d->k, Kp, d->c_phys, wfmt, c.tl, wt);
                if ((rc = check_launch("wt_prep_kernel"))) return rc;
            }
            Params p;
            memset(&p, 0, sizeof(p));
            p.M = (uint32_t)((size_t)d->n * c.Hi * c.Wj);
            p.HiWj = c.Hi * c.Wj; p.Wj = c.Wj;
            const int bn = d->c > 128 ? 256 : (d->c > 64 ? 128 : 64);
            p.n_tiles = (d->c + bn - 1) / bn;
            p.num_items = (int)((p.M + kBM - 1) / kBM) * p.n_tiles;
            p.kchunks = Kp / 64; p.ntaps = c.T; p.num_kb = c.T * p.kchunks;
            p.org_w = c.dminw; p.org_h = c.dminh; p.sw = p.sh = 1;
            for (int t = 0; t < c.T; ++t) { p.off_w[t] = (uint16_t)(c.dw[t] - c.dminw); p.off_h[t] = (uint16_t)(c.dh[t] - c.dminh); }
            p.out = dx; p.C = d->c; p.H = d->h; p.W = d->w;
            p.osh = d->stride_h; p.osw = d->stride_w; p.oa = c.a; p.ob = c.b;
            p.vec_ok = (d->c % 4 == 0 && (((uintptr_t)dx) & 15u) == 0) ? 1 : 0;
            p.scale = scale; p.post = kw;
            CUtensorMap ta, tb;
            {
                const cuuint64_t gdim[4] = {(cuuint64_t)Kp, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)d->n};
                const cuuint64_t gstr[3] = {(cuuint64_t)Kp * 2, (cuuint64_t)Kp * 2 * Wo, (cuuint64_t)Kp * 2 * Wo * Ho};
                const int lower[2] = {c.dminw, c.dminh};
                const int upper[2] = {c.dminw + c.Wj - Wo, c.dminh + c.Hi - Ho};
                const cuuint32_t estr[4] = {1, 1, 1, 1};
                CUresult cr = enc_im2col(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, G, gdim, gstr, lower, upper, 64u, (cuuint32_t)kBM, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (cr != CUDA_SUCCESS) return set_error(SLFP_ERR_DRIVER, "conv2d_bwd: cuTensorMapEncodeIm2col(G) failed (%d)", (int)cr);
            }
            {
                const cuuint64_t gdim[2] = {(cuuint64_t)c.T * Kp, (cuuint64_t)d->c_phys};
                const cuuint64_t gstr[1] = {(cuuint64_t)c.T * Kp * 2};
                const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)bn};
                const cuuint32_t estr[2] = {1, 1};
                CUresult cr = enc_tiled(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, wt, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (cr != CUDA_SUCCESS) return set_error(SLFP_ERR_DRIVER, "conv2d_bwd: cuTensorMapEncodeTiled(Wt) failed (%d)", (int)cr);
            }
            rc = bn == 256 ? launch<256, 0>(ta, tb, p, st) : (bn == 128 ? launch<128, 0>(ta, tb, p, st) : launch<64, 0>(ta, tb, p, st));
            if (rc) return rc;
        }
    }

    // ---------------- wgrad ----------------
    if (dwt && !L.dw_ok) {
        if ((rc = conv2d_bwd_direct(d, gy, x_codes, w_codes, wfmt, ka, kw, nullptr, dwt, so, sc, sr, ss, nullptr, st))) return rc;
    } else if (dwt) {
        if (!x_codes) return set_error(SLFP_ERR_BAD_ARG, "conv2d_bwd: dw needs x_codes");
        __half* Xh = (__half*)(ws + L.off_x);
        float* wacc = (float*)(ws + L.off_wacc);
        const size_t nx = (size_t)d->n * d->h * d->w * d->c_phys;
        {
            const int grid = (int)std::min<size_t>((size_t)num_sms() * 16, ceil_div_sz(nx / 16, 256));
            codes_to_f16_kernel<<<grid, 256, 0, st>>>(x_codes, nx / 16, d->fmt, Xh);
            if ((rc = check_launch("codes_to_f16_kernel"))) return rc;
        }
        const int taps = d->r * d->s;
        cudaMemsetAsync(wacc, 0, (size_t)d->k * taps * d->c_phys * sizeof(float), st);
        Params p;
        memset(&p, 0, sizeof(p));
        p.M = (uint32_t)mout;
        p.HiWj = Ho * Wo; p.Wj = Wo;
        p.ntaps = taps; p.cchunks = d->c_phys / 64; p.units = taps * p.cchunks;
        int upt = p.units >= 4 ? 4 : p.units;
        if (p.units > 4 && p.units % 4 != 0 && p.units % 3 == 0) upt = 3;
        p.ugroups = (p.units + upt - 1) / upt;
        const int ko_tiles = (d->k + kBM - 1) / kBM;
        const int tiles = ko_tiles * p.ugroups;
        p.pb_total = (int)((mout + kWgradPix - 1) / kWgradPix);
        int splits = std::max(1, (2 * num_sms()) / tiles);
        splits = std::min(splits, p.pb_total);
        p.pb_per_split = (p.pb_total + splits - 1) / splits;
        p.splits = (p.pb_total + p.pb_per_split - 1) / p.pb_per_split;
        p.num_items = tiles * p.splits;
        p.org_w = -d->pad_w; p.org_h = -d->pad_h; p.sw = d->stride_w; p.sh = d->stride_h;
        for (int r = 0; r < d->r; ++r)
            for (int s = 0; s < d->s; ++s) { p.off_w[r * d->s + s] = (uint16_t)(s * d->dil_w); p.off_h[r * d->s + s] = (uint16_t)(r * d->dil_h); }
        p.wacc = wacc; p.K = d->k; p.Cp = d->c_phys; p.scale = scale; p.post = ka;
        CUtensorMap ta, tb;
        {
            // G as [pixels][Kp]; the buffer carries 64 zeroed rows of slack so the last 64-pixel box stays inside it
            const cuuint64_t gdim[2] = {(cuuint64_t)Kp, (cuuint64_t)mout};
            const cuuint64_t gstr[1] = {(cuuint64_t)Kp * 2};
            const cuuint32_t box[2] = {64u, (cuuint32_t)kWgradPix};
            const cuuint32_t estr[2] = {1, 1};
            CUresult cr = enc_tiled(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, G, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (cr != CUDA_SUCCESS) return set_error(SLFP_ERR_DRIVER, "conv2d_bwd: cuTensorMapEncodeTiled(G) failed (%d)", (int)cr);
        }
        {
            const cuuint64_t gdim[4] = {(cuuint64_t)d->c_phys, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->n};
            const cuuint64_t gstr[3] = {(cuuint64_t)d->c_phys * 2, (cuuint64_t)d->c_phys * 2 * d->w, (cuuint64_t)d->c_phys * 2 * d->w * d->h};
            const int lower[2] = {-d->pad_w, -d->pad_h};
            const int upper[2] = {d->pad_w - (d->s - 1) * d->dil_w, d->pad_h - (d->r - 1) * d->dil_h};
            const cuuint32_t estr[4] = {1, (cuuint32_t)d->stride_w, (cuuint32_t)d->stride_h, 1};
            CUresult cr = enc_im2col(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, Xh, gdim, gstr, lower, upper, 64u, (cuuint32_t)kWgradPix, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (cr != CUDA_SUCCESS) return set_error(SLFP_ERR_DRIVER, "conv2d_bwd: cuTensorMapEncodeIm2col(X) failed (%d)", (int)cr);
        }
        rc = upt == 4 ? launch<256, 1>(ta, tb, p, st) : (upt == 3 ? launch<192, 1>(ta, tb, p, st) : (upt == 2 ? launch<128, 1>(ta, tb, p, st) : launch<64, 1>(ta, tb, p, st)));
        if (rc) return rc;
        {
            const size_t tot = (size_t)d->k * d->c * taps;
            const int grid = (int)std::min<size_t>((size_t)num_sms() * 8, ceil_div_sz(tot, 256));
            wgrad_finalize_kernel<<<grid, 256, 0, st>>>(wacc, d->k, d->c, d->c_phys, d->r, d->s, scale, ka, dwt, so, sc, sr, ss);
            if ((rc = check_launch("wgrad_finalize_kernel"))) return rc;
        }
    }
    if (db) return conv2d_bwd_direct(d, gy, nullptr, nullptr, wfmt, ka, kw, nullptr, nullptr, 0, 0, 0, 0, db, st);
    return 0;
}

}  // namespace slfp
