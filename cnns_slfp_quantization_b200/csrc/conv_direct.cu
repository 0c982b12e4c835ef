// conv_direct.cu -- CUDA-core kernels around the tensor-core path:
//   * depthwise / grouped convolution forward on 8-bit codes (MobileNet, ShuffleNet): an HBM-bound
//     stencil, 4 channels per thread, fp32 accumulate (Conv2d_Q.forward with groups > 1,
//     utils/conv2d_func.py:20-25; callers nets_imgnet/mobilenetv1.py:27, nets_cifar/shufflenet_v2.py:68)
//   * convolution backward with the identity straight-through estimator (utils/sfp_quant.py:50-53):
//       dx = dgrad(gy*Ka*Kw, w_q)/Ka   dw = wgrad(gy*Ka*Kw, x_q)/Kw   db = sum(gy)
//   * max-pool on codes and global average pool (glue of the fused eval pipeline).
#include <stdlib.h>

#include "slfp_common.cuh"
#include "sm100_ptx.cuh"

namespace slfp {

struct DirectParams {
    const uint8_t* x;   // NHWC codes, Cp physical channels
    const uint8_t* w;   // [K][R][S][Cg] weight codes (same format family as the activations)
    int N, H, W, C, Cp, K, R, S, sh, sw, ph, pw, dh, dw, Ho, Wo, groups;
    SlfpEpilogue epi;
};

__device__ __forceinline__ void epilogue_store(const SlfpEpilogue& e, float t, size_t pix, int k, int Kout,
                                               const uint32_t* tab) {
    (void)tab;
    if (e.ch_mul) {
        t = fmaf(t, __ldg(e.ch_mul + k), __ldg(e.ch_add + k));        // folded affine (fused eval pipeline)
    } else {
        if (e.bias_q) t += __ldg(e.bias_q + k);
        t = t * e.post_a;
        t = t * e.post_b;
        if (e.ch_scale) t = fmaf(t, __ldg(e.ch_scale + k), __ldg(e.ch_shift + k));
    }
    const size_t off = pix * Kout + k;
    if (e.residual)
        t += e.residual_f16 ? __half2float(reinterpret_cast<const __half*>(e.residual)[off])
                            : reinterpret_cast<const float*>(e.residual)[off];
    if (e.layerout) t = layerout_quantize(t, e.layerout == 2);
    if (e.relu) t = (t != t) ? t : fmaxf(t, 0.0f);             // torch.relu keeps NaN (fmaxf would drop it)
    if (e.y_f32) e.y_f32[off] = t;
    if (e.y_f16) reinterpret_cast<__half*>(e.y_f16)[off] = __float2half_rn(t);
    if (e.y_codes) {
        const float q = div_rn(t, e.next_k_div);
        e.y_codes[pix * e.k_phys_out + k] = (uint8_t)(e.next_fmt == SLFP_FMT_E4M3 ? encode_e4m3(q)
            : (e.next_fmt == SLFP_FMT_SFP33) ? encode<SLFP_FMT_SFP33>(q) : encode<SLFP_FMT_SLFP34_ACT>(q));
    }
    if (e.y_codes2) {
        const float q = div_rn(t, e.next_k_div2);
        e.y_codes2[pix * e.k_phys_out + k] = (uint8_t)(e.next_fmt == SLFP_FMT_E4M3 ? encode_e4m3(q)
            : (e.next_fmt == SLFP_FMT_SFP33) ? encode<SLFP_FMT_SFP33>(q) : encode<SLFP_FMT_SLFP34_ACT>(q));
    }
}

// Depthwise (C == K == groups): thread = (pixel, 4 consecutive channels).  Loads are one 32-bit word
// of codes per tap; a warp covers 128 contiguous channels (or several pixels when C < 128).
// Code -> float32 through a 256-entry shared table and the filter decoded once per CTA into shared memory
// ([C][R*S] float32, dynamic shared memory): the per-tap arithmetic decode of both operands was ~250 instructions
// per output.
template <bool SFP33>
__global__ void __launch_bounds__(256) dwconv_fwd_kernel(DirectParams p) {
    __shared__ uint32_t s_tab[16];
    __shared__ float s_dec[256];
    extern __shared__ float s_wf[];                           // [C][R*S]
    if (threadIdx.x < 16) s_tab[threadIdx.x] = c_pow2frac[threadIdx.x];
    __syncthreads();
    s_dec[threadIdx.x] = decode<SFP33>(threadIdx.x, s_tab);
    const int taps = p.R * p.S;
    for (int i = threadIdx.x; i < p.C * taps; i += 256) s_wf[i] = decode<SFP33>(__ldg(p.w + i), s_tab);
    __syncthreads();
    const int cq = p.Cp >> 2;                                 // channel quads per pixel
    const size_t total = (size_t)p.N * p.Ho * p.Wo * cq;
    for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256) {
        const int c0 = (int)(idx % cq) * 4;
        const size_t pix = idx / cq;
        const int wo = (int)(pix % p.Wo);
        const int ho = (int)((pix / p.Wo) % p.Ho);
        const int n = (int)(pix / ((size_t)p.Wo * p.Ho));
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int r = 0; r < p.R; ++r) {
            const int hi = ho * p.sh - p.ph + r * p.dh;
            if (hi < 0 || hi >= p.H) continue;
            for (int s = 0; s < p.S; ++s) {
                const int wi = wo * p.sw - p.pw + s * p.dw;
                if (wi < 0 || wi >= p.W) continue;
                const uint32_t wd = __ldg(reinterpret_cast<const uint32_t*>(p.x + (((size_t)n * p.H + hi) * p.W + wi) * p.Cp + c0));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (c0 + j < p.C)
                        acc[j] = fmaf(s_dec[(wd >> (8 * j)) & 0xffu], s_wf[(c0 + j) * taps + r * p.S + s], acc[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (c0 + j < p.K) epilogue_store(p.epi, acc[j], pix, c0 + j, p.K, s_tab);
    }
}

// ---- depthwise forward, fused-pipeline form (HBM-bound stencil) ------------------------------------------------
// thread = (output pixel, 16 consecutive channels): one 16-byte code load per filter tap (coalesced across the
// channel groups of a pixel), codes -> float32 through a bank-conflict-free table in shared memory (any of the four
// activation code formats; 256 entries x 32 bank copies), float32 weights of all taps staged in shared memory once
// per CTA, float32 accumulate; folded per-channel affine + ReLU + the 2-instruction post-ReLU encoder
// (slfp_common.cuh encode_relu_fast_raw16), one 16-byte code store per thread.
// Algorithmic bytes per output element: taps/stride^2 in (each input byte is fetched from DRAM once; the window
// overlap hits L1/L2) + 1 out.
struct DwFastParams {
    const uint8_t* x;
    const uint8_t* w;          // [C][R*S] weight codes
    int N, H, W, Cp, C, R, S, sh, sw, ph, pw, dh, dw, Ho, Wo;
    int act_fmt, wgt_sfp33, out_sfp33;
    int out_signed;            // 1: SLFP_FMT_SFP33_SFAST (no ReLU: sign bit + 7-bit magnitude code), 0: post-ReLU codes
    int out_e4m3;              // 1: SLFP_FMT_E4M3 (round-half-even e4m3 bytes; `relu` tells whether a ReLU precedes)
    int relu;
    float rk;                  // 1 / Ka_next
    const float* ch_mul;
    const float* ch_add;
    float sc;                  // 1 / (16 Ka_next)
    uint8_t* y;
    // strip kernel: 32-bit work-item index split by magic-number division (n / d == umulhi(n, mg) >> sh for n < 2^31, d > 1);
    // the three 64-bit divisions + two modulos per work item were ~500 of its ~3 000 instructions
    unsigned total32, cg_mg, cg_sh, ws_mg, ws_sh, ho_mg, ho_sh;
};
static void dw_magic(unsigned d, unsigned& mg, unsigned& sh) {
    mg = 0; sh = 0;
    if (d > 1) {
        unsigned lg = 31 - (unsigned)__builtin_clz(d);
        if (d & (d - 1)) ++lg;
        const unsigned pw = 31 + lg;
        mg = (unsigned)(((1ull << pw) + d - 1) / d);
        sh = pw - 32;
    }
}
__device__ __forceinline__ unsigned dw_div(unsigned n, unsigned d, unsigned mg, unsigned sh) { return d == 1u ? n : (__umulhi(n, mg) >> sh); }

// weight of (tap t, channel c = 16 q + 4 g + e) lives at s_w[t * Cp + (g * (Cp / 16) + q) * 4 + e]: the threads of a
// warp differ in q, so their 16-byte weight loads for a fixed g are consecutive (bank-conflict free)
template <int TAPS>       // R * S when it is 9 (3x3: all nine 16-byte loads are issued before the first use), else 0
__global__ void __launch_bounds__(256) dwconv_fast_kernel(const DwFastParams p) {
    extern __shared__ __align__(128) uint8_t dsm[];
    uint32_t* s_lut = reinterpret_cast<uint32_t*>(dsm);                 // [256 codes][32 banks] float32 bits
    float* s_w = reinterpret_cast<float*>(dsm + 256 * 32 * 4);          // [taps][Cp], permuted as above
    const int taps = p.R * p.S;
    const int cg = p.Cp >> 4;
    for (int i = threadIdx.x; i < 256 * 32; i += 256) s_lut[i] = __float_as_uint(decode_act_any((uint32_t)(i >> 5), p.act_fmt, c_pow2frac));
    for (int i = threadIdx.x; i < taps * p.Cp; i += 256) {
        const int t = i / p.Cp, c = i - t * p.Cp;
        float wv = 0.0f;
        if (c < p.C) {
            const uint32_t code = p.w[(size_t)c * taps + t];
            wv = p.wgt_sfp33 ? decode<true>(code, c_pow2frac) : decode<false>(code, c_pow2frac);
        }
        const int q = c >> 4, g = (c >> 2) & 3, e = c & 3;
        s_w[t * p.Cp + (g * cg + q) * 4 + e] = wv;
    }
    __syncthreads();
    const uint32_t lut_base = ptx::smem_u32(s_lut), w_base = ptx::smem_u32(s_w);
    const uint32_t lane4 = (threadIdx.x & 31u) * 4u;
    const size_t total = (size_t)p.N * p.Ho * p.Wo * cg;
    for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256) {
        const int q = (int)(idx % cg), c0 = q * 16;
        const size_t pix = idx / cg;
        const int wo = (int)(pix % p.Wo), ho = (int)((pix / p.Wo) % p.Ho), n = (int)(pix / ((size_t)p.Wo * p.Ho));
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
        auto tap_fma = [&](const uint4& cw, int t) {
            const uint32_t wds[4] = {cw.x, cw.y, cw.z, cw.w};
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const uint32_t c = wds[g];
                const float4 w4 = ptx::lds128_f4(w_base + (uint32_t)(t * p.Cp + (g * cg + q) * 4) * 4u);
                const float x0 = __uint_as_float(ptx::lds32_off(ptx::and_or(c << 7, 0x7f80u, lane4), lut_base));
                const float x1 = __uint_as_float(ptx::lds32_off(ptx::and_or(c >> 1, 0x7f80u, lane4), lut_base));
                const float x2 = __uint_as_float(ptx::lds32_off(ptx::and_or(c >> 9, 0x7f80u, lane4), lut_base));
                const float x3 = __uint_as_float(ptx::lds32_off(ptx::and_or(c >> 17, 0x7f80u, lane4), lut_base));
                acc[4 * g + 0] = fmaf(x0, w4.x, acc[4 * g + 0]);
                acc[4 * g + 1] = fmaf(x1, w4.y, acc[4 * g + 1]);
                acc[4 * g + 2] = fmaf(x2, w4.z, acc[4 * g + 2]);
                acc[4 * g + 3] = fmaf(x3, w4.w, acc[4 * g + 3]);
            }
        };
        const uint8_t* xn = p.x + (size_t)n * p.H * p.W * p.Cp + c0;
        if (TAPS == 9) {
            uint4 cw[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int hi = ho * p.sh - p.ph + (t / 3) * p.dh, wi = wo * p.sw - p.pw + (t % 3) * p.dw;
                cw[t] = make_uint4(0u, 0u, 0u, 0u);                                 // zero padding: code 0 = 0.0
                if (hi >= 0 && hi < p.H && wi >= 0 && wi < p.W)
                    cw[t] = __ldg(reinterpret_cast<const uint4*>(xn + ((size_t)hi * p.W + wi) * p.Cp));
            }
#pragma unroll
            for (int t = 0; t < 9; ++t) tap_fma(cw[t], t);
        } else {
            int t = 0;
            for (int r = 0; r < p.R; ++r) {
                const int hi = ho * p.sh - p.ph + r * p.dh;
                for (int s = 0; s < p.S; ++s, ++t) {
                    const int wi = wo * p.sw - p.pw + s * p.dw;
                    if (hi < 0 || hi >= p.H || wi < 0 || wi >= p.W) continue;
                    tap_fma(__ldg(reinterpret_cast<const uint4*>(xn + ((size_t)hi * p.W + wi) * p.Cp)), t);
                }
            }
        }
        int32_t tq[16];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const float4 m4 = __ldg(reinterpret_cast<const float4*>(p.ch_mul + c0) + g);
            const float4 a4 = __ldg(reinterpret_cast<const float4*>(p.ch_add + c0) + g);
            const float v0 = fmaf(acc[4 * g + 0], m4.x, a4.x), v1 = fmaf(acc[4 * g + 1], m4.y, a4.y);
            const float v2 = fmaf(acc[4 * g + 2], m4.z, a4.z), v3 = fmaf(acc[4 * g + 3], m4.w, a4.w);
            if (p.out_e4m3) {
                const uint32_t lo = p.relu ? encode_e4m3x2_relu(v0 * p.rk, v1 * p.rk) : encode_e4m3x2(v0 * p.rk, v1 * p.rk);
                const uint32_t hi = p.relu ? encode_e4m3x2_relu(v2 * p.rk, v3 * p.rk) : encode_e4m3x2(v2 * p.rk, v3 * p.rk);
                tq[4 * g + 0] = (int32_t)(lo & 0xffu); tq[4 * g + 1] = (int32_t)(lo >> 8);
                tq[4 * g + 2] = (int32_t)(hi & 0xffu); tq[4 * g + 3] = (int32_t)(hi >> 8);
            } else if (p.out_signed) {
                const float vv[4] = {v0, v1, v2, v3};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int32_t m = encode_relu_fast_raw16<true>(__saturatef(fabsf(vv[e]) * p.sc));
                    tq[4 * g + e] = (m < 0 ? 0 : (m > 127 ? 127 : m)) | (int32_t)((__float_as_uint(vv[e]) >> 24) & 0x80u);
                }
            } else if (p.out_sfp33) {
                tq[4 * g + 0] = encode_relu_fast_raw16<true>(__saturatef(v0 * p.sc)); tq[4 * g + 1] = encode_relu_fast_raw16<true>(__saturatef(v1 * p.sc));
                tq[4 * g + 2] = encode_relu_fast_raw16<true>(__saturatef(v2 * p.sc)); tq[4 * g + 3] = encode_relu_fast_raw16<true>(__saturatef(v3 * p.sc));
            } else {
                tq[4 * g + 0] = encode_relu_fast_raw16<false>(__saturatef(v0 * p.sc)); tq[4 * g + 1] = encode_relu_fast_raw16<false>(__saturatef(v1 * p.sc));
                tq[4 * g + 2] = encode_relu_fast_raw16<false>(__saturatef(v2 * p.sc)); tq[4 * g + 3] = encode_relu_fast_raw16<false>(__saturatef(v3 * p.sc));
            }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) tq[j] = (c0 + j < p.C) ? tq[j] : 0;             // zero code in pad channels
        uint4 pk;
        pk.x = ptx::pack_sat_u8x4(tq[0], tq[1], tq[2], tq[3]);   pk.y = ptx::pack_sat_u8x4(tq[4], tq[5], tq[6], tq[7]);
        pk.z = ptx::pack_sat_u8x4(tq[8], tq[9], tq[10], tq[11]); pk.w = ptx::pack_sat_u8x4(tq[12], tq[13], tq[14], tq[15]);
        *reinterpret_cast<uint4*>(p.y + pix * p.Cp + c0) = pk;
    }
}

// 3x3, stride 1 / 2, dilation 1: a thread computes a strip of 4 output pixels along W for its 4*VEC channels.  Every
// input vector (pixel x 4*VEC channels) of the 3 x (3*stride+3) window is fetched and decoded ONCE for the strip
// (4.5 table look-ups per output at stride 1 instead of 9) and every weight vector is read once per 4 outputs.
// Both forms hold 64 accumulators in 128 registers (two CTAs = 16 warps per SM, one resident wave): VEC = 4 is 4 outputs x 16
// channels (16-byte vectors); VEC = 2 is 8 channels (8-byte vectors) x 8 outputs at stride 1 (3.75 instead of 4.5 table decodes and
// half the weight reads per output) or x 4 outputs at stride 2 (no spills, where the 16-channel form spills its 9-column window).
// OUT: 0 = post-ReLU fast codes, 1 = signed fast SFP<3,3>, 2 = e4m3 bytes (compile-time: the run-time form spilled),
// 3 = e4m3 bytes out AND in: the input bytes are decoded by the hardware converter (cvt.rn.f16x2.e4m3x2: two values per
// instruction + one float16 -> float32 each, exact) instead of the shared-memory table (shift + LOP3 + LDS per value)
template <int STRIDE, int VEC, int OUT>
__global__ void __launch_bounds__(256, 2) dwconv3x3_strip_kernel(const DwFastParams p) {
    extern __shared__ __align__(128) uint8_t dsm[];
    uint32_t* s_lut = reinterpret_cast<uint32_t*>(dsm);
    float* s_w = reinterpret_cast<float*>(dsm + 256 * 32 * 4);
    constexpr int CH = 4 * VEC;
    const int cg = p.Cp / CH;
    {   // lane-replicated decode table: each thread decodes ONE code, the warp broadcasts it into the 32 lane copies
        const uint32_t mine = __float_as_uint(decode_act_any((uint32_t)threadIdx.x, p.act_fmt, c_pow2frac));
        const int lane = threadIdx.x & 31, code0 = threadIdx.x & ~31;
#pragma unroll 8
        for (int j = 0; j < 32; ++j) s_lut[(code0 + j) * 32 + lane] = __shfl_sync(0xffffffffu, mine, j);
    }
    for (int c = threadIdx.x; c < p.Cp; c += 256) {            // a thread decodes the nine contiguous taps of its channels
        const int dst = (((c >> 2) & (VEC - 1)) * cg + c / CH) * 4 + (c & 3);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            float wv = 0.0f;
            if (c < p.C) {
                const uint32_t code = p.w[(size_t)c * 9 + t];
                wv = p.wgt_sfp33 ? decode<true>(code, c_pow2frac) : decode<false>(code, c_pow2frac);
            }
            s_w[t * p.Cp + dst] = wv;
        }
    }
    __syncthreads();
    const uint32_t lut_base = ptx::smem_u32(s_lut), w_base = ptx::smem_u32(s_w);
    const uint32_t lane4 = (threadIdx.x & 31u) * 4u;
    // strip length: 4 outputs of 16 channels, or (VEC = 2) 8 outputs of 8 channels at stride 1 - the same 64 accumulators,
    // 10 instead of 12 decoded input columns per 8 outputs and half the weight reads per output
    constexpr int kOut = (VEC == 2 && STRIDE == 1) ? 8 : 4, kCols = (kOut - 1) * STRIDE + 3;
    constexpr bool kHwDecode = OUT == 3;
    constexpr bool kE4m3Out = OUT == 2 || OUT == 3;
    const int wstrips = (p.Wo + kOut - 1) / kOut;
    const int enc_shift = p.out_sfp33 ? 19 : 18, enc_bias = p.out_sfp33 ? 0x76F : 0xEDF;    // encode_relu_fast_raw16<>
    const bool pad_channels = p.C != p.Cp;
    for (unsigned idx = blockIdx.x * 256u + threadIdx.x; idx < p.total32; idx += gridDim.x * 256u) {
        unsigned rest = dw_div(idx, (unsigned)cg, p.cg_mg, p.cg_sh);
        const int q = (int)(idx - rest * (unsigned)cg), c0 = q * CH;
        const unsigned r1 = dw_div(rest, (unsigned)wstrips, p.ws_mg, p.ws_sh);
        const int ws = (int)(rest - r1 * (unsigned)wstrips);
        const unsigned r2 = dw_div(r1, (unsigned)p.Ho, p.ho_mg, p.ho_sh);
        const int ho = (int)(r1 - r2 * (unsigned)p.Ho), n = (int)r2;
        const int wo0 = ws * kOut;
        float acc[kOut][CH];
#pragma unroll
        for (int o = 0; o < kOut; ++o)
#pragma unroll
            for (int j = 0; j < CH; ++j) acc[o][j] = 0.0f;
        const uint8_t* xn = p.x + (size_t)n * p.H * p.W * p.Cp + c0;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int hi = ho * STRIDE - p.ph + r;
            const bool hok = hi >= 0 && hi < p.H;
            uint32_t cw[kCols][VEC];
#pragma unroll
            for (int cidx = 0; cidx < kCols; ++cidx) {
                const int wi = wo0 * STRIDE - p.pw + cidx;
#pragma unroll
                for (int g = 0; g < VEC; ++g) cw[cidx][g] = 0u;
                if (hok && wi >= 0 && wi < p.W) {
                    const uint8_t* src = xn + ((size_t)hi * p.W + wi) * p.Cp;
                    if (VEC == 4) {
                        const uint4 v = __ldg(reinterpret_cast<const uint4*>(src));
                        cw[cidx][0] = v.x; cw[cidx][1] = v.y; cw[cidx][VEC - 2] = v.z; cw[cidx][VEC - 1] = v.w;
                    } else {
                        const uint2 v = __ldg(reinterpret_cast<const uint2*>(src));
                        cw[cidx][0] = v.x; cw[cidx][1] = v.y;
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < VEC; ++g) {
                float4 w4[3];
#pragma unroll
                for (int s = 0; s < 3; ++s) w4[s] = ptx::lds128_f4(w_base + (uint32_t)((r * 3 + s) * p.Cp + (g * cg + q) * 4) * 4u);
#pragma unroll
                for (int cidx = 0; cidx < kCols; ++cidx) {
                    const uint32_t c = cw[cidx][g];
                    float x0, x1, x2, x3;
                    if (kHwDecode) {
                        uint32_t h01, h23;
                        asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\t"
                            "cvt.rn.f16x2.e4m3x2 %0, lo;\n\tcvt.rn.f16x2.e4m3x2 %1, hi;\n\t}\n" : "=r"(h01), "=r"(h23) : "r"(c));
                        const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&h01));
                        const float2 f23 = __half22float2(*reinterpret_cast<const __half2*>(&h23));
                        x0 = f01.x; x1 = f01.y; x2 = f23.x; x3 = f23.y;
                    } else {
                        x0 = __uint_as_float(ptx::lds32_off(ptx::and_or(c << 7, 0x7f80u, lane4), lut_base));
                        x1 = __uint_as_float(ptx::lds32_off(ptx::and_or(c >> 1, 0x7f80u, lane4), lut_base));
                        x2 = __uint_as_float(ptx::lds32_off(ptx::and_or(c >> 9, 0x7f80u, lane4), lut_base));
                        x3 = __uint_as_float(ptx::lds32_off(ptx::and_or(c >> 17, 0x7f80u, lane4), lut_base));
                    }
#pragma unroll
                    for (int o = 0; o < kOut; ++o) {
                        const int s = cidx - o * STRIDE;                    // filter column this input column has for output o
                        if (s >= 0 && s < 3) {
                            acc[o][4 * g + 0] = fmaf(x0, w4[s].x, acc[o][4 * g + 0]);
                            acc[o][4 * g + 1] = fmaf(x1, w4[s].y, acc[o][4 * g + 1]);
                            acc[o][4 * g + 2] = fmaf(x2, w4[s].z, acc[o][4 * g + 2]);
                            acc[o][4 * g + 3] = fmaf(x3, w4[s].w, acc[o][4 * g + 3]);
                        }
                    }
                }
            }
        }
        float4 m4[VEC], a4[VEC];
#pragma unroll
        for (int g = 0; g < VEC; ++g) {
            m4[g] = __ldg(reinterpret_cast<const float4*>(p.ch_mul + c0) + g);
            a4[g] = __ldg(reinterpret_cast<const float4*>(p.ch_add + c0) + g);
            if (kE4m3Out) {                                   // fold 1 / Ka_next into the affine: one FFMA yields the quotient
                m4[g].x *= p.rk; m4[g].y *= p.rk; m4[g].z *= p.rk; m4[g].w *= p.rk;
                a4[g].x *= p.rk; a4[g].y *= p.rk; a4[g].z *= p.rk; a4[g].w *= p.rk;
            }
        }
        const size_t pix0 = ((size_t)n * p.Ho + ho) * p.Wo + wo0;
#pragma unroll
        for (int o = 0; o < kOut; ++o) {
            if (wo0 + o >= p.Wo) break;
            int32_t tq[CH];
            uint32_t e4w[VEC];
#pragma unroll
            for (int g = 0; g < VEC; ++g) {
                const float mm[4] = {m4[g].x, m4[g].y, m4[g].z, m4[g].w}, aa[4] = {a4[g].x, a4[g].y, a4[g].z, a4[g].w};
                if (kE4m3Out) {
                    float y4[4];                                // (m4 / a4 carry 1 / Ka_next: see above)
#pragma unroll
                    for (int e = 0; e < 4; ++e) y4[e] = fmaf(acc[o][4 * g + e], mm[e], aa[e]);
                    const uint32_t lo = p.relu ? encode_e4m3x2_relu(y4[0], y4[1]) : encode_e4m3x2(y4[0], y4[1]);
                    const uint32_t hi = p.relu ? encode_e4m3x2_relu(y4[2], y4[3]) : encode_e4m3x2(y4[2], y4[3]);
                    e4w[g] = lo | (hi << 16);
                    continue;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float y = fmaf(acc[o][4 * g + e], mm[e], aa[e]);
                    if ((OUT == 1)) {                 // sign bit + min(code(|y|), 127); a negative zero keeps magnitude code 0
                        const int32_t m = ((int32_t)__float_as_uint(__saturatef(fabsf(y) * p.sc)) >> 19) - 0x76F;
                        tq[4 * g + e] = (m < 0 ? 0 : (m > 127 ? 127 : m)) | (int32_t)((__float_as_uint(y) >> 24) & 0x80u);
                    } else {
                        const float v = __saturatef(y * p.sc);
                        tq[4 * g + e] = ((int32_t)__float_as_uint(v) >> enc_shift) - enc_bias;
                    }
                }
            }
            if (pad_channels) {                                                      // zero code in pad channels
#pragma unroll
                for (int j = 0; j < CH; ++j) tq[j] = (c0 + j < p.C) ? tq[j] : 0;
            }
            uint32_t pk[VEC];
            if (kE4m3Out) {
#pragma unroll
                for (int g = 0; g < VEC; ++g) {
                    pk[g] = e4w[g];
                    if (pad_channels) {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (c0 + 4 * g + e >= p.C) pk[g] &= ~(0xffu << (8 * e));
                    }
                }
            } else {
#pragma unroll
            for (int g = 0; g < VEC; ++g) pk[g] = ptx::pack_sat_u8x4(tq[4 * g], tq[4 * g + 1], tq[4 * g + 2], tq[4 * g + 3]);
            }
            uint8_t* dst = p.y + (pix0 + o) * p.Cp + c0;
            if (VEC == 4) *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[VEC - 2], pk[VEC - 1]);
            else *reinterpret_cast<uint2*>(dst) = make_uint2(pk[0], pk[1]);
        }
    }
}

// Generic grouped convolution: thread = (pixel, output channel), loops over Cg x R x S.
template <bool SFP33>
__global__ void __launch_bounds__(256) gconv_fwd_kernel(DirectParams p) {
    __shared__ uint32_t s_tab[16];
    if (threadIdx.x < 16) s_tab[threadIdx.x] = c_pow2frac[threadIdx.x];
    __syncthreads();
    const int Cg = p.C / p.groups, Kg = p.K / p.groups;
    const size_t total = (size_t)p.N * p.Ho * p.Wo * p.K;
    for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256) {
        const int k = (int)(idx % p.K);
        const size_t pix = idx / p.K;
        const int wo = (int)(pix % p.Wo);
        const int ho = (int)((pix / p.Wo) % p.Ho);
        const int n = (int)(pix / ((size_t)p.Wo * p.Ho));
        const int g = k / Kg;
        float acc = 0.f;
        for (int r = 0; r < p.R; ++r) {
            const int hi = ho * p.sh - p.ph + r * p.dh;
            if (hi < 0 || hi >= p.H) continue;
            for (int s = 0; s < p.S; ++s) {
                const int wi = wo * p.sw - p.pw + s * p.dw;
                if (wi < 0 || wi >= p.W) continue;
                const uint8_t* xp = p.x + (((size_t)n * p.H + hi) * p.W + wi) * p.Cp + g * Cg;
                const uint8_t* wp = p.w + (((size_t)k * p.R + r) * p.S + s) * Cg;
                for (int c = 0; c < Cg; ++c) acc = fmaf(decode<SFP33>(xp[c], s_tab), decode<SFP33>(__ldg(wp + c), s_tab), acc);
            }
        }
        epilogue_store(p.epi, acc, pix, k, p.K, s_tab);
    }
}

int conv2d_fwd_grouped(const SlfpConvDesc* d, const uint8_t* x_codes, const void* w_codes, const SlfpEpilogue* epi,
                       cudaStream_t st) {
    DirectParams p;
    p.x = x_codes; p.w = (const uint8_t*)w_codes;
    p.N = d->n; p.H = d->h; p.W = d->w; p.C = d->c; p.Cp = d->c_phys; p.K = d->k; p.R = d->r; p.S = d->s;
    p.sh = d->stride_h; p.sw = d->stride_w; p.ph = d->pad_h; p.pw = d->pad_w; p.dh = d->dil_h; p.dw = d->dil_w;
    p.groups = d->groups;
    p.Ho = (d->h + 2 * d->pad_h - d->dil_h * (d->r - 1) - 1) / d->stride_h + 1;
    p.Wo = (d->w + 2 * d->pad_w - d->dil_w * (d->s - 1) - 1) / d->stride_w + 1;
    p.epi = *epi;
    if (p.Ho <= 0 || p.Wo <= 0 || d->n <= 0) return set_error(SLFP_ERR_BAD_ARG, "conv2d_fwd(grouped): empty output");
    if (d->groups <= 0 || d->c % d->groups || d->k % d->groups)
        return set_error(SLFP_ERR_BAD_ARG, "conv2d_fwd(grouped): channels not divisible by groups");
    if (d->pad_h_extra || d->pad_w_extra)
        return set_error(SLFP_ERR_UNSUPPORTED, "conv2d_fwd(grouped): asymmetric padding is a dense-path feature");
    {
        // fused-pipeline depthwise: folded affine + ReLU + post-ReLU codes only, 16-channel vectors
        const bool e4 = epi->next_fmt == SLFP_FMT_E4M3;
        const bool sfast = (epi->next_fmt == SLFP_FMT_SFP33_SFAST && !epi->relu) || e4;
        const bool relu_out = (epi->next_fmt == SLFP_FMT_SLFP34_RELU || epi->next_fmt == SLFP_FMT_SFP33_RELU) && epi->relu;
        const bool dw = d->groups == d->c && d->k == d->c;
        const size_t smem = 256 * 32 * 4 + (size_t)d->r * d->s * d->c_phys * 4;
        if (dw && (relu_out || sfast) && !epi->layerout && epi->ch_mul && epi->ch_add && epi->y_codes && !epi->y_codes2 && !epi->y_f16 && !epi->y_f32 &&
            !epi->residual && epi->k_phys_out == d->c_phys && d->c_phys % 16 == 0 && epi->next_k_div > 0.f && smem <= 200 * 1024 &&
            ((((uintptr_t)x_codes | (uintptr_t)epi->y_codes | (uintptr_t)epi->ch_mul | (uintptr_t)epi->ch_add) & 15u) == 0)) {
            DwFastParams q;
            q.x = x_codes; q.w = (const uint8_t*)w_codes;
            q.N = d->n; q.H = d->h; q.W = d->w; q.Cp = d->c_phys; q.C = d->c; q.R = d->r; q.S = d->s;
            q.sh = d->stride_h; q.sw = d->stride_w; q.ph = d->pad_h; q.pw = d->pad_w; q.dh = d->dil_h; q.dw = d->dil_w;
            q.Ho = p.Ho; q.Wo = p.Wo;
            q.act_fmt = d->fmt;
            q.wgt_sfp33 = (d->fmt == SLFP_FMT_SFP33 || d->fmt == SLFP_FMT_SFP33_RELU || d->fmt == SLFP_FMT_SFP33_SFAST || d->fmt == SLFP_FMT_E4M3) ? 1 : 0;    // weights: the q_bit's weight format
            q.out_sfp33 = (epi->next_fmt == SLFP_FMT_SFP33_RELU || sfast) ? 1 : 0;
            q.out_signed = (sfast && !e4) ? 1 : 0;
            q.out_e4m3 = e4 ? 1 : 0;
            q.relu = epi->relu;
            q.rk = (float)(1.0 / (double)epi->next_k_div);
            q.ch_mul = epi->ch_mul; q.ch_add = epi->ch_add;
            q.sc = (float)(1.0 / (16.0 * (double)epi->next_k_div));
            q.y = epi->y_codes;
            static DeviceOnce attr_once;
    bool& attr_done = attr_once.flag();
            if (!attr_done) {
                cudaError_t e = cudaFuncSetAttribute(dwconv_fast_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
                if (e == cudaSuccess) e = cudaFuncSetAttribute(dwconv_fast_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
                if (e != cudaSuccess) return set_error((int)e, "dwconv_fast: smem attribute: %s", cudaGetErrorString(e));
                attr_done = true;
            }
            const size_t tot = (size_t)d->n * p.Ho * p.Wo * (d->c_phys / 16);
            int blocks_per_sm = (int)((220 * 1024) / (smem + 1024));
            blocks_per_sm = blocks_per_sm > 6 ? 6 : (blocks_per_sm < 1 ? 1 : blocks_per_sm);
            const int g = (int)min((size_t)num_sms() * blocks_per_sm, ceil_div_sz(tot, 256));
            const bool strip = d->r == 3 && d->s == 3 && d->dil_h == 1 && d->dil_w == 1 && d->stride_h == d->stride_w &&
                               (d->stride_h == 1 || d->stride_h == 2) && p.Wo >= 4 && getenv("SLFP_DW_NO_STRIP") == nullptr;
            if (strip) {
                // VEC = 2 (8 channels per thread, 8-output strips at stride 1) by default: measured faster than the 16-channel
                // form on every MobileNetV1 layer (step 2.08 -> 1.95 ms); SLFP_DW_VEC=4 selects the 16-channel form
                static int vec = 0;
                if (!vec) {
                    const char* ev = getenv("SLFP_DW_VEC");
                    vec = (ev && ev[0] == '4') ? 4 : 2;
                }
                static const bool no_hw = getenv("SLFP_DW_LUT_DECODE") != nullptr;
                const int om = q.out_e4m3 ? ((d->fmt == SLFP_FMT_E4M3 && !no_hw) ? 3 : 2) : (q.out_signed ? 1 : 0);
                const int k_out = (vec == 2 && d->stride_h == 1) ? 8 : 4;                     // kOut of the kernel
                const size_t tot4 = (size_t)d->n * p.Ho * ((p.Wo + k_out - 1) / k_out) * (d->c_phys / (4 * vec));
                if (tot4 >= (1ull << 31)) return set_error(SLFP_ERR_UNSUPPORTED, "dwconv3x3_strip: more than 2^31 work items");
                q.total32 = (unsigned)tot4;
                dw_magic((unsigned)(d->c_phys / (4 * vec)), q.cg_mg, q.cg_sh);
                dw_magic((unsigned)((p.Wo + k_out - 1) / k_out), q.ws_mg, q.ws_sh);
                dw_magic((unsigned)p.Ho, q.ho_mg, q.ho_sh);
                // exactly one resident wave: the occupancy the kernel really gets with this layer's shared memory
#define SLFP_STRIP(S_, V_, O_)                                                                                          \
                if (d->stride_h == S_ && vec == V_ && om == O_) {                                                       \
                    auto kern = dwconv3x3_strip_kernel<S_, V_, O_>;                                                     \
                    static bool attr = false;                                                                           \
                    if (!attr) {                                                                                        \
                        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)); \
                        if (e != cudaSuccess) return set_error((int)e, "dwconv3x3_strip: smem attribute: %s", cudaGetErrorString(e)); \
                        attr = true;                                                                                    \
                    }                                                                                                   \
                    int per_sm = 0;                                                                                     \
                    cudaError_t eo = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem);            \
                    if (eo != cudaSuccess || per_sm < 1) return set_error((int)eo, "dwconv3x3_strip: occupancy query failed"); \
                    const int g4 = (int)min((size_t)num_sms() * per_sm, ceil_div_sz(tot4, 256));                       \
                    kern<<<g4, 256, smem, st>>>(q);                                                                     \
                    return check_launch("dwconv3x3_strip_kernel");                                                      \
                }
                SLFP_STRIP(1, 2, 0) SLFP_STRIP(1, 2, 1) SLFP_STRIP(1, 2, 2) SLFP_STRIP(2, 2, 0) SLFP_STRIP(2, 2, 1) SLFP_STRIP(2, 2, 2)
                SLFP_STRIP(1, 2, 3) SLFP_STRIP(2, 2, 3) SLFP_STRIP(1, 4, 3) SLFP_STRIP(2, 4, 3)
                SLFP_STRIP(1, 4, 0) SLFP_STRIP(1, 4, 1) SLFP_STRIP(1, 4, 2) SLFP_STRIP(2, 4, 0) SLFP_STRIP(2, 4, 1) SLFP_STRIP(2, 4, 2)
#undef SLFP_STRIP
                return check_launch("dwconv3x3_strip_kernel");
            }
            if (d->r == 3 && d->s == 3) dwconv_fast_kernel<9><<<g, 256, smem, st>>>(q);
            else dwconv_fast_kernel<0><<<g, 256, smem, st>>>(q);
            return check_launch("dwconv_fast_kernel");
        }
    }
    if ((d->fmt != SLFP_FMT_SFP33 && d->fmt != SLFP_FMT_SLFP34_ACT) ||
        (epi->y_codes && epi->next_fmt != SLFP_FMT_SFP33 && epi->next_fmt != SLFP_FMT_SLFP34_ACT))
        return set_error(SLFP_ERR_UNSUPPORTED, "conv2d_fwd(grouped): the generic stencil kernels read and write the signed code formats only");
    const bool sfp = d->fmt == SLFP_FMT_SFP33;
    const bool depthwise = d->groups == d->c && d->k == d->c && (d->c_phys % 4) == 0 && (((uintptr_t)x_codes) & 3u) == 0;
    const size_t total = depthwise ? (size_t)d->n * p.Ho * p.Wo * (d->c_phys / 4) : (size_t)d->n * p.Ho * p.Wo * d->k;
    const int grid = (int)min((size_t)num_sms() * 16, ceil_div_sz(total, 256));
    if (depthwise) {
        const size_t wsm = (size_t)d->c * d->r * d->s * sizeof(float);
        if (wsm > 40000) return set_error(SLFP_ERR_UNSUPPORTED, "conv2d_fwd(depthwise): filter of %zu bytes does not fit shared memory", wsm);
        if (sfp) dwconv_fwd_kernel<true><<<grid, 256, wsm, st>>>(p); else dwconv_fwd_kernel<false><<<grid, 256, wsm, st>>>(p);
    } else {
        if (sfp) gconv_fwd_kernel<true><<<grid, 256, 0, st>>>(p); else gconv_fwd_kernel<false><<<grid, 256, 0, st>>>(p);
    }
    return check_launch("grouped conv forward");
}

// ---- backward (CUDA cores; any groups / stride / dilation) ---------------------------------------------
struct BwdParams {
    const float* gy;        // NHWC [N,Ho,Wo,K]
    const uint8_t* x;       // NHWC codes [N,H,W,Cp]
    const uint8_t* wc;      // KRSC codes, row pitch wpitch
    int N, H, W, C, Cp, K, R, S, sh, sw, ph, pw, dh, dw, Ho, Wo, groups;
    size_t wpitch;
    int Cw;                 // channels per tap in the weight operand (Cp dense, Cg grouped)
    float ka, kw;
    float* dx;              // NHWC [N,H,W,C]
    float* dwt; long long so, sc, sr, ss;
    float* db;
    int pix_splits;
};

// dx[n,hi,wi,c] = sum_{k,r,s : hi = ho*sh - ph + r*dh, ...} (gy*Kw*Ka)[n,ho,wo,k] * w_q[k,r,s,c] / Ka
template <bool SFP33W>
__global__ void __launch_bounds__(256) dgrad_kernel(BwdParams p) {
    __shared__ uint32_t s_tab[16];
    if (threadIdx.x < 16) s_tab[threadIdx.x] = c_pow2frac[threadIdx.x];
    __syncthreads();
    const int Cg = p.C / p.groups, Kg = p.K / p.groups;
    const size_t total = (size_t)p.N * p.H * p.W * p.C;
    for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256) {
        const int c = (int)(idx % p.C);
        const size_t pin = idx / p.C;
        const int wi = (int)(pin % p.W);
        const int hi = (int)((pin / p.W) % p.H);
        const int n = (int)(pin / ((size_t)p.W * p.H));
        const int g = c / Cg, cl = c - g * Cg;
        float acc = 0.f;
        for (int r = 0; r < p.R; ++r) {
            const int th = hi + p.ph - r * p.dh;
            if (th < 0 || th % p.sh) continue;
            const int ho = th / p.sh;
            if (ho >= p.Ho) continue;
            for (int s = 0; s < p.S; ++s) {
                const int tw = wi + p.pw - s * p.dw;
                if (tw < 0 || tw % p.sw) continue;
                const int wo = tw / p.sw;
                if (wo >= p.Wo) continue;
                const float* gp = p.gy + (((size_t)n * p.Ho + ho) * p.Wo + wo) * p.K + g * Kg;
                for (int kk = 0; kk < Kg; ++kk) {
                    const float gs = (__ldg(gp + kk) * p.kw) * p.ka;
                    const uint8_t code = p.wc[(size_t)(g * Kg + kk) * p.wpitch + (size_t)(r * p.S + s) * p.Cw + (p.groups > 1 ? cl : c)];
                    acc = fmaf(gs, decode<SFP33W>(code, s_tab), acc);
                }
            }
        }
        p.dx[idx] = div_rn(acc, p.ka);
    }
}

// dw[k,r,s,c] = sum_pixels (gy*Kw*Ka)[pix,k] * x_q[pix@(r,s), c] / Kw ; grid (K*R*S, c-blocks, pixel splits)
template <bool SFP33A>
__global__ void __launch_bounds__(128) wgrad_kernel(BwdParams p) {
    __shared__ uint32_t s_tab[16];
    if (threadIdx.x < 16) s_tab[threadIdx.x] = c_pow2frac[threadIdx.x];
    __syncthreads();
    const int Cg = p.C / p.groups, Kg = p.K / p.groups;
    const int krs = blockIdx.x;
    const int k = krs / (p.R * p.S), rs = krs % (p.R * p.S), r = rs / p.S, s = rs % p.S;
    const int cl = blockIdx.y * 128 + threadIdx.x;
    if (cl >= Cg) return;
    const int g = k / Kg, c = g * Cg + cl;
    const size_t npix = (size_t)p.N * p.Ho * p.Wo;
    const size_t per = ceil_div_sz(npix, (size_t)p.pix_splits);
    const size_t p0 = per * blockIdx.z, p1 = min(npix, p0 + per);
    float acc = 0.f;
    for (size_t pix = p0; pix < p1; ++pix) {
        const int wo = (int)(pix % p.Wo);
        const int ho = (int)((pix / p.Wo) % p.Ho);
        const int n = (int)(pix / ((size_t)p.Wo * p.Ho));
        const int hi = ho * p.sh - p.ph + r * p.dh, wi = wo * p.sw - p.pw + s * p.dw;
        if (hi < 0 || hi >= p.H || wi < 0 || wi >= p.W) continue;
        const float gs = (__ldg(p.gy + pix * p.K + k) * p.kw) * p.ka;
        const uint8_t code = p.x[(((size_t)n * p.H + hi) * p.W + wi) * p.Cp + c];
        acc = fmaf(gs, decode<SFP33A>(code, s_tab), acc);
    }
    float* dst = p.dwt + k * p.so + cl * p.sc + r * p.sr + s * p.ss;
    if (p.pix_splits == 1) *dst = div_rn(acc, p.kw);
    else atomicAdd(dst, div_rn(acc, p.kw));
}

// Dense layers with few input channels (network stems, C = 3): the reduction over output pixels is long and
// the weight tensor tiny, so one thread owns one (r, s, c) weight column with 64 output channels in registers and
// a block streams its range of output pixels, staging 32 pixels x 64 gy values in shared memory (read back as
// broadcast float4s: 16 LDS per 64 FMAs).  Partial sums are added with atomics (dw zeroed by the caller).
template <bool SFP33A>
__global__ void __launch_bounds__(256) wgrad_smallc_kernel(BwdParams p) {
    __shared__ uint32_t s_tab[16];
    __shared__ __align__(16) float s_g[32][64];
    __shared__ int s_pn[32], s_ph[32], s_pw[32];
    if (threadIdx.x < 16) s_tab[threadIdx.x] = c_pow2frac[threadIdx.x];
    const int cols = p.R * p.S * p.C;
    const int col = blockIdx.y * blockDim.x + threadIdx.x;
    const bool active = col < cols;
    const int c = col % p.C, rs = col / p.C, s = rs % p.S, r = rs / p.S;
    const int k0 = blockIdx.z * 64;
    const size_t npix = (size_t)p.N * p.Ho * p.Wo;
    const size_t per = ceil_div_sz(ceil_div_sz(npix, (size_t)gridDim.x), 32) * 32;
    const size_t p0 = per * blockIdx.x, p1 = min(npix, p0 + per);
    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = 0.f;
    for (size_t pix0 = p0; pix0 < p1; pix0 += 32) {
        __syncthreads();
        for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) {
            const int pp = i >> 6, kk = i & 63;
            const size_t pix = pix0 + pp;
            s_g[pp][kk] = (pix < p1 && k0 + kk < p.K) ? __ldg(p.gy + pix * p.K + k0 + kk) : 0.f;
        }
        if (threadIdx.x < 32) {
            const size_t pix = pix0 + threadIdx.x;
            const int wo = (int)(pix % p.Wo), ho = (int)((pix / p.Wo) % p.Ho);
            s_pn[threadIdx.x] = pix < p1 ? (int)(pix / ((size_t)p.Wo * p.Ho)) : -1;
            s_ph[threadIdx.x] = ho * p.sh - p.ph;
            s_pw[threadIdx.x] = wo * p.sw - p.pw;
        }
        __syncthreads();
        if (active) {
#pragma unroll 1
            for (int pp = 0; pp < 32; ++pp) {
                const int n = s_pn[pp];
                const int hi = s_ph[pp] + r * p.dh, wi = s_pw[pp] + s * p.dw;
                if (n < 0 || hi < 0 || hi >= p.H || wi < 0 || wi >= p.W) continue;
                const float xv = decode<SFP33A>(p.x[(((size_t)n * p.H + hi) * p.W + wi) * p.Cp + c], s_tab);
#pragma unroll
                for (int kq = 0; kq < 16; ++kq) {
                    const float4 g = *reinterpret_cast<const float4*>(&s_g[pp][kq * 4]);
                    acc[kq * 4] = fmaf(g.x, xv, acc[kq * 4]);
                    acc[kq * 4 + 1] = fmaf(g.y, xv, acc[kq * 4 + 1]);
                    acc[kq * 4 + 2] = fmaf(g.z, xv, acc[kq * 4 + 2]);
                    acc[kq * 4 + 3] = fmaf(g.w, xv, acc[kq * 4 + 3]);
                }
            }
        }
    }
    if (active) {
#pragma unroll
        for (int kk = 0; kk < 64; ++kk)
            if (k0 + kk < p.K) atomicAdd(p.dwt + (k0 + kk) * p.so + c * p.sc + r * p.sr + s * p.ss, acc[kk] * p.ka);
    }
}

__global__ void __launch_bounds__(256) dbias_kernel(const float* __restrict__ gy, size_t npix, int K, float* __restrict__ db) {
    const int k = blockIdx.x;
    float acc = 0.f;
    for (size_t pix = threadIdx.x; pix < npix; pix += 256) acc += gy[pix * K + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ float s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += s[i];
        db[k] = t;
    }
}

int conv2d_bwd_direct(const SlfpConvDesc* d, const float* gy, const uint8_t* x_codes, const uint8_t* w_codes, int wfmt,
                      float ka, float kw, float* dx, float* dwt, long long so, long long sc, long long sr, long long ss,
                      float* db, cudaStream_t st) {
    BwdParams p;
    p.gy = gy; p.x = x_codes; p.wc = w_codes;
    p.N = d->n; p.H = d->h; p.W = d->w; p.C = d->c; p.Cp = d->c_phys; p.K = d->k; p.R = d->r; p.S = d->s;
    p.sh = d->stride_h; p.sw = d->stride_w; p.ph = d->pad_h; p.pw = d->pad_w; p.dh = d->dil_h; p.dw = d->dil_w;
    p.groups = d->groups;
    p.Ho = (d->h + 2 * d->pad_h - d->dil_h * (d->r - 1) - 1) / d->stride_h + 1;
    p.Wo = (d->w + 2 * d->pad_w - d->dil_w * (d->s - 1) - 1) / d->stride_w + 1;
    p.wpitch = slfp_conv_wpitch(d);
    p.Cw = d->groups > 1 ? d->c / d->groups : d->c_phys;
    p.ka = ka; p.kw = kw; p.dx = dx; p.dwt = dwt; p.so = so; p.sc = sc; p.sr = sr; p.ss = ss; p.db = db;
    const size_t npix = (size_t)d->n * p.Ho * p.Wo;
    if (d->fmt != SLFP_FMT_SFP33 && d->fmt != SLFP_FMT_SLFP34_ACT)
        return set_error(SLFP_ERR_UNSUPPORTED, "conv2d_bwd: activation codes must be in a signed quantizer format");
    if (d->pad_h_extra || d->pad_w_extra) return set_error(SLFP_ERR_UNSUPPORTED, "conv2d_bwd: asymmetric padding");
    const bool sfp_w = wfmt == SLFP_FMT_SFP33, sfp_a = d->fmt == SLFP_FMT_SFP33;
    int rc = 0;
    if (dx) {
        if (!w_codes) return set_error(SLFP_ERR_BAD_ARG, "conv2d_bwd: dx needs w_codes");
        const size_t total = (size_t)d->n * d->h * d->w * d->c;
        const int grid = (int)min((size_t)num_sms() * 16, ceil_div_sz(total, 256));
        if (sfp_w) dgrad_kernel<true><<<grid, 256, 0, st>>>(p); else dgrad_kernel<false><<<grid, 256, 0, st>>>(p);
        if ((rc = check_launch("dgrad_kernel"))) return rc;
    }
    if (dwt) {
        if (!x_codes) return set_error(SLFP_ERR_BAD_ARG, "conv2d_bwd: dw needs x_codes");
        const int Cg = d->c / d->groups;
        const int krs = d->k * d->r * d->s, cb = (Cg + 127) / 128;
        const bool contiguous_dw = ss == 1 && sr == d->s && sc == (long long)d->r * d->s && so == (long long)Cg * d->r * d->s;
        if (d->groups == 1 && d->c <= 16 && contiguous_dw && getenv("SLFP_WGRAD_GENERIC") == nullptr) {
            const int cols = d->r * d->s * d->c;
            const int threads = min(256, (cols + 31) / 32 * 32);
            const int cblocks = (cols + threads - 1) / threads, ktiles = (d->k + 63) / 64;
            const int splits = (int)max((size_t)1, min(ceil_div_sz(npix, 64), (size_t)(num_sms() * 6 / (cblocks * ktiles))));
            cudaMemsetAsync(dwt, 0, (size_t)d->k * Cg * d->r * d->s * sizeof(float), st);
            dim3 grid(splits, cblocks, ktiles);
            if (sfp_a) wgrad_smallc_kernel<true><<<grid, threads, 0, st>>>(p); else wgrad_smallc_kernel<false><<<grid, threads, 0, st>>>(p);
            if ((rc = check_launch("wgrad_smallc_kernel"))) return rc;
            if (db) {
                dbias_kernel<<<d->k, 256, 0, st>>>(gy, npix, d->k, db);
                if ((rc = check_launch("dbias_kernel"))) return rc;
            }
            return 0;
        }
        int splits = 1;
        while ((size_t)krs * cb * splits < (size_t)num_sms() * 8 && (size_t)splits * 256 < npix && splits < 1024) splits *= 2;
        p.pix_splits = splits;
        if (splits > 1) {
            // atomics accumulate: the destination must start from zero (strided views zeroed by caller)
            const bool contiguous = ss == 1 && sr == d->s && sc == (long long)d->r * d->s && so == (long long)Cg * d->r * d->s;
            if (contiguous) cudaMemsetAsync(dwt, 0, (size_t)d->k * Cg * d->r * d->s * sizeof(float), st);
            else p.pix_splits = 1;
        }
        dim3 grid(krs, cb, p.pix_splits);
        if (sfp_a) wgrad_kernel<true><<<grid, 128, 0, st>>>(p); else wgrad_kernel<false><<<grid, 128, 0, st>>>(p);
        if ((rc = check_launch("wgrad_kernel"))) return rc;
    }
    if (db) {
        dbias_kernel<<<d->k, 256, 0, st>>>(gy, npix, d->k, db);
        if ((rc = check_launch("dbias_kernel"))) return rc;
    }
    return 0;
}

// ---- pooling glue -------------------------------------------------------------------------------------
// Order key of a code: monotone in the decoded value (sat literal above the top grid value).
__device__ __forceinline__ int code_key(uint32_t c) {
    uint32_t u = c & 0x7fu;
    u = (u == kCodeSat) ? 128u : u;
    u = (u == kCodeNaN) ? 200u : u;          // NaN wins, like torch max-pool propagating NaN
    return (c & 0x80u) ? -(int)u : (int)u;
}

__global__ void __launch_bounds__(256) maxpool_codes_kernel(const uint8_t* __restrict__ x, int N, int H, int W, int Cp,
                                                            int kh, int kw, int stride, int pad, int Ho, int Wo,
                                                            uint8_t* __restrict__ y) {
    const int cq = Cp >> 2;
    const size_t total = (size_t)N * Ho * Wo * cq;
    for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256) {
        const int c0 = (int)(idx % cq) * 4;
        const size_t pix = idx / cq;
        const int wo = (int)(pix % Wo), ho = (int)((pix / Wo) % Ho), n = (int)(pix / ((size_t)Wo * Ho));
        int best[4] = {-1000, -1000, -1000, -1000};
        uint32_t bc[4] = {0, 0, 0, 0};
        for (int r = 0; r < kh; ++r) {
            const int hi = ho * stride - pad + r;
            if (hi < 0 || hi >= H) continue;
            for (int s = 0; s < kw; ++s) {
                const int wi = wo * stride - pad + s;
                if (wi < 0 || wi >= W) continue;
                const uint32_t wd = __ldg(reinterpret_cast<const uint32_t*>(x + (((size_t)n * H + hi) * W + wi) * Cp + c0));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t c = (wd >> (8 * j)) & 0xffu;
                    const int key = code_key(c);
                    if (key > best[j]) { best[j] = key; bc[j] = c; }
                }
            }
        }
        *reinterpret_cast<uint32_t*>(y + pix * Cp + c0) = bc[0] | (bc[1] << 8) | (bc[2] << 16) | (bc[3] << 24);
    }
}

// Byte-wise unsigned max has no native instruction (__vmaxu4 is a ~10-instruction emulation: the first version of the 3x3 / 2
// kernel spent 560 of its 770 SASS instructions per thread there and was issue-bound at 3.5 TB/s, ncu
// profiles/r04_final.md); 16-bit lanes do (max.u16x2 = VIMNMX.U16x2).  Each loaded word is split once into its even and
// odd bytes (two PRMT), all maxima run on those halves, and one PRMT per word re-interleaves the result.
struct U16x8 { uint32_t e[4], o[4]; };                                  // 16 bytes as 8 + 8 zero-extended 16-bit lanes
__device__ __forceinline__ uint32_t max_u16x2(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("max.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ U16x8 split16(uint4 v) {
    U16x8 r;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { r.e[i] = __byte_perm(w[i], 0u, 0x4240); r.o[i] = __byte_perm(w[i], 0u, 0x4341); }
    return r;
}
__device__ __forceinline__ U16x8 max16(const U16x8& a, const U16x8& b) {
    U16x8 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) { r.e[i] = max_u16x2(a.e[i], b.e[i]); r.o[i] = max_u16x2(a.o[i], b.o[i]); }
    return r;
}
__device__ __forceinline__ uint4 join16(const U16x8& a) {
    return make_uint4(__byte_perm(a.e[0], a.o[0], 0x6240), __byte_perm(a.e[1], a.o[1], 0x6240),
                      __byte_perm(a.e[2], a.o[2], 0x6240), __byte_perm(a.e[3], a.o[3], 0x6240));
}
// Post-ReLU (unsigned, monotone) codes: byte-wise unsigned max, 16 channels per thread.  HBM-bound: each
// input byte is read once from DRAM (window overlap hits L1/L2), one 16-byte store per thread.
__global__ void __launch_bounds__(256) maxpool_ucodes_kernel(const uint8_t* __restrict__ x, int N, int H, int W, int Cp,
                                                             int kh, int kw, int stride, int pad, int Ho, int Wo,
                                                             uint8_t* __restrict__ y) {
    const int cq = Cp >> 4;
    const size_t total = (size_t)N * Ho * Wo * cq;
    for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256) {
        const int c0 = (int)(idx % cq) * 16;
        const size_t pix = idx / cq;
        const int wo = (int)(pix % Wo), ho = (int)((pix / Wo) % Ho), n = (int)(pix / ((size_t)Wo * Ho));
        U16x8 best = split16(make_uint4(0u, 0u, 0u, 0u));               // maxima on 16-bit lanes (max.u16x2), see above
        for (int r = 0; r < kh; ++r) {
            const int hi = ho * stride - pad + r;
            if (hi < 0 || hi >= H) continue;
            for (int s = 0; s < kw; ++s) {
                const int wi = wo * stride - pad + s;
                if (wi < 0 || wi >= W) continue;
                best = max16(best, split16(__ldg(reinterpret_cast<const uint4*>(x + (((size_t)n * H + hi) * W + wi) * Cp + c0))));
            }
        }
        *reinterpret_cast<uint4*>(y + pix * Cp + c0) = join16(best);
    }
}

// The ResNet stem pool (3x3, stride 2, padding 1) on post-ReLU codes: a thread produces TWO horizontally adjacent
// outputs of 16 channels from one 3 x 5 window - 15 independent 16-byte loads in flight per thread (the generic
// kernel's runtime-bounded tap loop kept one or two) and 7.5 instead of 9 loads per output.
__global__ void __launch_bounds__(256) maxpool3x3s2_ucodes_kernel(const uint8_t* __restrict__ x, int N, int H, int W, int Cp,
                                                                  int Ho, int Wo, uint8_t* __restrict__ y) {
    // work items < 2^31 (host): 32-bit index arithmetic
    const uint32_t cq = (uint32_t)Cp >> 4, Wp = ((uint32_t)Wo + 1u) >> 1;
    const uint32_t total = (uint32_t)N * Ho * Wp * cq;
    for (uint32_t idx = blockIdx.x * 256u + threadIdx.x; idx < total; idx += gridDim.x * 256u) {
        uint32_t t = idx / cq;
        const int c0 = (int)(idx - t * cq) * 16;
        const uint32_t t2 = t / Wp;
        const int wp = (int)(t - t2 * Wp);
        const int n = (int)(t2 / (uint32_t)Ho);
        const int ho = (int)(t2 - (uint32_t)n * Ho);
        const int wo = 2 * wp, wi0 = 2 * wo - 1, hi0 = 2 * ho - 1;
        uint4 v[3][5];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int hi = hi0 + r;
            const bool rok = hi >= 0 && hi < H;
            const uint8_t* row = x + (((size_t)n * H + (rok ? hi : 0)) * W) * Cp + c0;
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                const int wi = wi0 + s;
                v[r][s] = (rok && wi >= 0 && wi < W) ? __ldg(reinterpret_cast<const uint4*>(row + (size_t)wi * Cp)) : make_uint4(0u, 0u, 0u, 0u);
            }
        }
        U16x8 col[5];
#pragma unroll
        for (int s = 0; s < 5; ++s) col[s] = max16(max16(split16(v[0][s]), split16(v[1][s])), split16(v[2][s]));
        uint8_t* dst = y + (((size_t)n * Ho + ho) * Wo + wo) * Cp + c0;
        *reinterpret_cast<uint4*>(dst) = join16(max16(max16(col[0], col[1]), col[2]));
        if (wo + 1 < Wo) *reinterpret_cast<uint4*>(dst + Cp) = join16(max16(max16(col[2], col[3]), col[4]));
    }
}

template <typename T>
__global__ void __launch_bounds__(256) avgpool_kernel(const T* __restrict__ x, int hw, int C, float* __restrict__ y) {
    const int n = blockIdx.y;
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= C) return;
    float acc = 0.f;
    const T* xp = x + (size_t)n * hw * C + c;
    for (int i = 0; i < hw; ++i) acc += (float)xp[(size_t)i * C];
    y[(size_t)n * C + c] = acc / (float)hw;
}

// Global average pool of float16 images, 8 channels per thread (one 16-byte load per pixel, seven in flight), with the
// classifier's activation quantizer fused: mean -> float32 [n, c] (optional) and -> encode(mean / K) code bytes.  The
// sum runs over the pixels in order, like avgpool_kernel (same float32 bits); the codes are encode<FMT>(div_k(mean)),
// i.e. what slfp_quantize_nhwc_f32 produces from the float32 means.  FMT < 0: no codes.
template <int FMT>
__global__ void __launch_bounds__(128) avgpool8_quantize_kernel(const __half* __restrict__ x, int hw, int C, float* __restrict__ y,
                                                                DivK dk, uint8_t* __restrict__ codes) {
    const int n = blockIdx.y;
    const int c0 = (blockIdx.x * 128 + threadIdx.x) * 8;
    if (c0 >= C) return;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const uint4* xp = reinterpret_cast<const uint4*>(x + (size_t)n * hw * C + c0);
    const size_t pitch = (size_t)(C >> 3);
    // seven pixels per round, all seven loads issued before the adds (an unrolled loop with its exit test between the
    // iterations keeps ONE load in flight); the sum still runs over the pixels in order
    for (int i0 = 0; i0 < hw; i0 += 7) {
        uint4 v[7];
#pragma unroll
        for (int u = 0; u < 7; ++u) v[u] = i0 + u < hw ? __ldg(xp + (size_t)(i0 + u) * pitch) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int u = 0; u < 7; ++u) {
            if (i0 + u < hw) {
                const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
                    acc[2 * j] += f.x;
                    acc[2 * j + 1] += f.y;
                }
            }
        }
    }
    const float fhw = (float)hw;
    uint32_t cw[2] = {0u, 0u};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        acc[j] = acc[j] / fhw;
        if (FMT >= 0) cw[j >> 2] |= encode<FMT < 0 ? 0 : FMT>(div_k(acc[j], dk)) << (8 * (j & 3));
    }
    if (y) {
        float4* yp = reinterpret_cast<float4*>(y + (size_t)n * C + c0);
        yp[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        yp[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    if (FMT >= 0) *reinterpret_cast<uint2*>(codes + (size_t)n * C + c0) = make_uint2(cw[0], cw[1]);
}

}  // namespace slfp

using namespace slfp;

extern "C" int slfp_maxpool_codes(const uint8_t* x, int n, int h, int w, int c_phys, int fmt, int kh, int kw_, int stride,
                                  int pad, uint8_t* y, slfp_stream_t stream) {
    if (!x || !y || (c_phys & 3) || (((uintptr_t)x | (uintptr_t)y) & 3u))
        return set_error(SLFP_ERR_BAD_ARG, "slfp_maxpool_codes: bad arguments");
    const int Ho = (h + 2 * pad - kh) / stride + 1, Wo = (w + 2 * pad - kw_) / stride + 1;
    if (fmt == SLFP_FMT_SLFP34_RELU || fmt == SLFP_FMT_SFP33_RELU) {
        if ((c_phys & 15) || (((uintptr_t)x | (uintptr_t)y) & 15u))
            return set_error(SLFP_ERR_BAD_ARG, "slfp_maxpool_codes: post-ReLU codes need c_phys %% 16 == 0 and 16-byte alignment");
        const size_t tot = (size_t)n * Ho * Wo * (c_phys / 16);
        if (tot == 0) return 0;
        const size_t tot2 = (size_t)n * Ho * ((Wo + 1) / 2) * (c_phys / 16);
        if (kh == 3 && kw_ == 3 && stride == 2 && pad == 1 && tot2 < (1ull << 31) && getenv("SLFP_POOL_GENERIC") == nullptr) {
            const int g2 = (int)min((size_t)num_sms() * 8, ceil_div_sz(tot2, 256));
            maxpool3x3s2_ucodes_kernel<<<g2, 256, 0, (cudaStream_t)stream>>>(x, n, h, w, c_phys, Ho, Wo, y);
            return check_launch("maxpool3x3s2_ucodes_kernel");
        }
        const int g = (int)min((size_t)num_sms() * 32, ceil_div_sz(tot, 256));
        maxpool_ucodes_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(x, n, h, w, c_phys, kh, kw_, stride, pad, Ho, Wo, y);
        return check_launch("maxpool_ucodes_kernel");
    }
    const size_t total = (size_t)n * Ho * Wo * (c_phys / 4);
    if (total == 0) return 0;
    const int grid = (int)min((size_t)num_sms() * 16, ceil_div_sz(total, 256));
    maxpool_codes_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, h, w, c_phys, kh, kw_, stride, pad, Ho, Wo, y);
    return check_launch("maxpool_codes_kernel");
}

extern "C" int slfp_avgpool_nhwc(const void* x, int is_f16, int n, int hw, int c, float* y, slfp_stream_t stream) {
    if (!x || !y) return set_error(SLFP_ERR_BAD_ARG, "slfp_avgpool_nhwc: null pointer");
    if (n <= 0 || c <= 0) return 0;
    dim3 grid((c + 255) / 256, n);
    if (is_f16) avgpool_kernel<__half><<<grid, 256, 0, (cudaStream_t)stream>>>((const __half*)x, hw, c, y);
    else avgpool_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, hw, c, y);
    return check_launch("avgpool_kernel");
}

extern "C" int slfp_avgpool_quantize_nhwc_f16(const void* x, int n, int hw, int c, float* y, float k_div, int fmt, uint8_t* codes,
                                              slfp_stream_t stream) {
    if (!x || (!y && !codes)) return set_error(SLFP_ERR_BAD_ARG, "slfp_avgpool_quantize_nhwc_f16: null pointer");
    if (n <= 0 || c <= 0) return 0;
    if (hw <= 0 || (c & 7) || (((uintptr_t)x | (uintptr_t)y) & 15u) || (((uintptr_t)codes) & 7u))
        return set_error(SLFP_ERR_BAD_ARG, "slfp_avgpool_quantize_nhwc_f16: needs c %% 8 == 0 and 16-byte aligned tensors");
    dim3 grid((c / 8 + 127) / 128, n);
    const DivK dk = make_divk(codes ? k_div : 1.0f);
    cudaStream_t st = (cudaStream_t)stream;
    if (!codes) avgpool8_quantize_kernel<-1><<<grid, 128, 0, st>>>((const __half*)x, hw, c, y, dk, nullptr);
    else if (fmt == SLFP_FMT_SFP33) avgpool8_quantize_kernel<SLFP_FMT_SFP33><<<grid, 128, 0, st>>>((const __half*)x, hw, c, y, dk, codes);
    else if (fmt == SLFP_FMT_SLFP34_ACT) avgpool8_quantize_kernel<SLFP_FMT_SLFP34_ACT><<<grid, 128, 0, st>>>((const __half*)x, hw, c, y, dk, codes);
    else return set_error(SLFP_ERR_BAD_ARG, "slfp_avgpool_quantize_nhwc_f16: format %d (SFP<3,3> / SLFP<3,4> activation codes only)", fmt);
    return check_launch("avgpool8_quantize_kernel");
}
