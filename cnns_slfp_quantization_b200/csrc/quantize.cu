// quantize.cu -- fused SLFP / SFP quantizer, de-quantizer, abs-max and weight preparation.
//
// Replaces the ~25-kernel ATen chains of utils/sfp_quant.py:14-47, 63-96, 111-126 plus the
// pre-scale division of utils/conv2d_func.py:21-22 with ONE pass over HBM:
//   4 B read + 1 B (codes) / 4 B (fake-quant fp32) / 2 B (fp16) written per element.
// The kernel is HBM-bound: every warp-level load is one contiguous 512 B request (128-bit per lane,
// L1 no-allocate), a thread keeps four of them in flight, and the grid is a multiple of the SM count.
#include <stdarg.h>
#include <stdio.h>

#include "slfp_common.cuh"

namespace slfp {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error((int)e, "%s: %s", what, cudaGetErrorString(e));
    return 0;
}

int num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
            cudaGetLastError();
            sms = 148;
        }
    }
    return sms;
}

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w) : "memory");
}

constexpr int kQThreads = 256;
constexpr int kQVec = 4;                       // float4 loads in flight per thread
constexpr int kQTile = kQThreads * 4 * kQVec;  // 4096 elements per CTA iteration

struct QuantArgs {
    const float* x;
    size_t n;
    DivK k_div;
    uint8_t* codes;
    float* fakeq;
    __half* f16;
    int zero_is_zero;
    // dynamic max-scaling (slfp_quantize_dyn_f32): K = float32(double(*k_src) / k_src_div) read from DEVICE memory when the
    // kernel starts (abs-max kernel -> [allreduce(MAX)] -> this kernel, no host round trip, CUDA-graph capturable)
    const float* k_src;
    double k_src_div;
    float* k_out;
};

__device__ __forceinline__ DivK resolve_divk(const QuantArgs& a) {
    if (a.k_src == nullptr) return a.k_div;
    // the reference computes the scale in Python float64 (np.array(max) / 15.5) and its arithmetic then uses float32(K)
    const float k = (float)((double)__ldg(a.k_src) / a.k_src_div);
    if (a.k_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *a.k_out = k;
    return make_divk(k);
}

template <int FMT>
__device__ __forceinline__ void quant_elem(float x, const DivK& k_div, bool zz, const uint32_t* tab,
                                           uint32_t& code, float& fq) {
    const float v = div_k(x, k_div);           // == IEEE x / K, like `input / self.Ka` on the CPU
    if (FMT == SLFP_FMT_SFP44_OUT) {
        code = 0;
        fq = layerout_quantize(v, zz);
    } else {
        code = encode<FMT>(v);
        fq = decode<FMT == SLFP_FMT_SFP33>(code, tab);
    }
}

// Fast element: reciprocal division (3 FMA-pipe ops, exact inside the normal range) + encode_q.  The
// caller checks `needs_exact` (dividend so small that the quotient could round to exactly 0, where only
// the true division classifies "zero" vs "tiny" correctly) and redoes those rare elements with div_k.
template <int FMT>
__device__ __forceinline__ uint32_t quant_code_fast(float x, const DivK& k) {
    return encode_q<FMT>(div_k_fused(x, k), x);
}
__device__ __forceinline__ bool needs_exact(float x) {
    const uint32_t ax = __float_as_uint(x) & 0x7fffffffu;
    return (ax - 1u) < 0x04000000u - 1u;                 // 0 < |x| < 2^-119
}

template <int FMT, bool CODES, bool FAKEQ, bool F16>
__global__ void __launch_bounds__(kQThreads) quantize_kernel(QuantArgs a) {
    __shared__ uint32_t s_tab[16];
    constexpr bool kLut = FMT == SLFP_FMT_SFP33 || FMT == SLFP_FMT_SLFP34_ACT;
    constexpr int FL = FMT == SLFP_FMT_SFP33 ? SLFP_FMT_SFP33 : SLFP_FMT_SLFP34_ACT;
    __shared__ uint8_t s_enc[kLut && CODES ? kEncLutBytes : 16];
    __shared__ uint32_t s_encf[kLut && (FAKEQ || F16) ? kEncLutBytes : 4];
    if (threadIdx.x < 16) s_tab[threadIdx.x] = c_pow2frac[threadIdx.x];
    __syncthreads();
    if (kLut) {
        for (int i = threadIdx.x; i < kEncLutBytes; i += kQThreads) {
            const uint32_t u = enc_lut_entry<FL>((uint32_t)i);
            if (CODES) s_enc[i] = (uint8_t)u;
            if (FAKEQ || F16) s_encf[i] = __float_as_uint(decode<FL == SLFP_FMT_SFP33>(u, s_tab));
        }
        __syncthreads();
    }
    const bool zz = a.zero_is_zero != 0;
    const DivK kdiv = resolve_divk(a);
    const size_t n_tiles = a.n / kQTile;
    for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const size_t base = tile * kQTile;
        float4 v[kQVec];
#pragma unroll
        for (int j = 0; j < kQVec; ++j)
            v[j] = ldg_stream(reinterpret_cast<const float4*>(a.x + base) + j * kQThreads + threadIdx.x);
#pragma unroll
        for (int j = 0; j < kQVec; ++j) {
            const size_t off = base + (size_t)(j * kQThreads + threadIdx.x) * 4;
            const float xs[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
            uint32_t c[4];
            float q[4];
            if (FMT == SLFP_FMT_SFP44_OUT || !kdiv.fast) {
#pragma unroll
                for (int i = 0; i < 4; ++i) quant_elem<FMT>(xs[i], kdiv, zz, s_tab, c[i], q[i]);
            } else {
                constexpr int F = FMT == SLFP_FMT_SFP44_OUT ? SLFP_FMT_SFP33 : FMT;
                // common path: reciprocal division (exact inside the normal range) + the in-range encoder.  One
                // group test sends the rare 4-element group to the general path: a NaN / Inf quotient, or a dividend so
                // small that only the true division classifies "zero" vs "tiny" correctly.
                float qv[4];
                float nan_probe = 0.0f;                       // q * 0 accumulates to NaN iff a quotient is NaN or Inf (FMA pipe)
                uint32_t xmin = 0xffffffffu;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    qv[i] = div_k_fused(xs[i], kdiv);
                    nan_probe = fmaf(qv[i], 0.0f, nan_probe);
                    // rotate the sign to bit 0: 2|x| + s.  +0 -> 0 (wraps to 0xffffffff: stays on the fast path), -0 -> 1 and
                    // 0 < |x| < 2^-119 -> below 0x08000000: both take the general path (a zero code carries no sign)
                    const uint32_t ax1 = __funnelshift_l(__float_as_uint(xs[i]), __float_as_uint(xs[i]), 1) - 1u;
                    xmin = ax1 < xmin ? ax1 : xmin;
                }
                if (kLut && !(nan_probe != nan_probe || xmin < 0x08000000u - 1u)) {
                    uint32_t idx[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) idx[i] = enc_lut_index<FL>(qv[i], xs[i]);
                    if (CODES) {
                        uint32_t u[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) u[i] = s_enc[idx[i]];
                        const uint32_t sg = __byte_perm(__byte_perm(__float_as_uint(qv[0]), __float_as_uint(qv[1]), 0x0073),
                                                        __byte_perm(__float_as_uint(qv[2]), __float_as_uint(qv[3]), 0x0073), 0x5410);
                        const uint32_t pk = __byte_perm(__byte_perm(u[0], u[1], 0x0040), __byte_perm(u[2], u[3], 0x0040), 0x5410);
                        *reinterpret_cast<uint32_t*>(a.codes + off) = (sg & 0x80808080u) | pk;
                    }
                    if (FAKEQ || F16) {
                        // |fake-quant value| straight from the float table, sign of the quotient OR-ed in (a zero stays +0)
                        float fq[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) fq[i] = __uint_as_float(s_encf[idx[i]] | (__float_as_uint(qv[i]) & 0x80000000u));
                        if (FAKEQ) stg_stream(reinterpret_cast<float4*>(a.fakeq + off), make_float4(fq[0], fq[1], fq[2], fq[3]));
                        if (F16) {
                            __half2 h0 = __floats2half2_rn(fq[0], fq[1]), h1 = __floats2half2_rn(fq[2], fq[3]);
                            uint2 pk;
                            pk.x = *reinterpret_cast<uint32_t*>(&h0);
                            pk.y = *reinterpret_cast<uint32_t*>(&h1);
                            *reinterpret_cast<uint2*>(a.f16 + off) = pk;
                        }
                    }
                    continue;
                }
                if (nan_probe != nan_probe || xmin < 0x08000000u - 1u) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) c[i] = encode<F>(div_k(xs[i], kdiv));
                } else if (F == SLFP_FMT_SLFP34_WGT) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) c[i] = encode_inrange<F>(qv[i]);
                } else {
                    constexpr int FA = F == SLFP_FMT_SLFP34_WGT ? SLFP_FMT_SLFP34_ACT : F;
#pragma unroll
                    for (int i = 0; i < 4; ++i) c[i] = encode_balanced<FA>(qv[i]);
                }
                if (FAKEQ || F16) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) q[i] = decode<F == SLFP_FMT_SFP33>(c[i], s_tab);
                }
            }
            if (CODES)
                *reinterpret_cast<uint32_t*>(a.codes + off) =
                    __byte_perm(__byte_perm(c[0], c[1], 0x0040), __byte_perm(c[2], c[3], 0x0040), 0x5410);
            if (FAKEQ) stg_stream(reinterpret_cast<float4*>(a.fakeq + off), make_float4(q[0], q[1], q[2], q[3]));
            if (F16) {
                __half2 h0 = __floats2half2_rn(q[0], q[1]), h1 = __floats2half2_rn(q[2], q[3]);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&h0);
                pk.y = *reinterpret_cast<uint32_t*>(&h1);
                *reinterpret_cast<uint2*>(a.f16 + off) = pk;
            }
        }
    }
    // ragged tail (< one tile): scalar, first CTA only
    if (blockIdx.x == 0) {
        for (size_t i = n_tiles * kQTile + threadIdx.x; i < a.n; i += kQThreads) {
            uint32_t c;
            float q;
            quant_elem<FMT>(a.x[i], kdiv, zz, s_tab, c, q);
            if (CODES) a.codes[i] = (uint8_t)c;
            if (FAKEQ) a.fakeq[i] = q;
            if (F16) a.f16[i] = __float2half_rn(q);
        }
    }
}

// fully scalar variant for mis-aligned pointers (sliced tensors); same arithmetic
template <int FMT>
__global__ void __launch_bounds__(kQThreads) quantize_scalar_kernel(QuantArgs a) {
    __shared__ uint32_t s_tab[16];
    if (threadIdx.x < 16) s_tab[threadIdx.x] = c_pow2frac[threadIdx.x];
    __syncthreads();
    const DivK kdiv = resolve_divk(a);
    for (size_t i = (size_t)blockIdx.x * kQThreads + threadIdx.x; i < a.n; i += (size_t)gridDim.x * kQThreads) {
        uint32_t c;
        float q;
        quant_elem<FMT>(a.x[i], kdiv, a.zero_is_zero != 0, s_tab, c, q);
        if (a.codes) a.codes[i] = (uint8_t)c;
        if (a.fakeq) a.fakeq[i] = q;
        if (a.f16) a.f16[i] = __float2half_rn(q);
    }
}

template <int FMT>
static int launch_quantize(const QuantArgs& a, cudaStream_t st) {
    const bool aligned = (((uintptr_t)a.x | (uintptr_t)a.fakeq) & 15u) == 0 && ((uintptr_t)a.codes & 3u) == 0 &&
                         ((uintptr_t)a.f16 & 7u) == 0;
    const int sms = num_sms();
    if (!aligned) {
        int grid = (int)min((size_t)sms * 8, ceil_div_sz(a.n, kQThreads));
        quantize_scalar_kernel<FMT><<<grid, kQThreads, 0, st>>>(a);
        return check_launch("quantize_scalar_kernel");
    }
    const size_t tiles = a.n / kQTile;
    const int sel = (a.codes ? 1 : 0) | (a.fakeq ? 2 : 0) | (a.f16 ? 4 : 0);
    // persistent grid-stride CTAs: exactly one resident wave (a grid larger than what fits leaves a second, thin wave
    // running alone at a fraction of the memory-level parallelism)
#define SLFP_QLAUNCH(C_, F_, H_)                                                                                   \
    {                                                                                                              \
        static int per_sm = 0;                                                                                     \
        if (!per_sm) {                                                                                             \
            int b = 0;                                                                                             \
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, quantize_kernel<FMT, C_, F_, H_>, kQThreads, 0) != cudaSuccess) { \
                cudaGetLastError();                                                                                \
                b = 0;                                                                                             \
            }                                                                                                      \
            per_sm = b > 0 ? b : 4;                                                                                \
        }                                                                                                          \
        const int grid = (int)max((size_t)1, min(tiles, (size_t)sms * per_sm));                                    \
        quantize_kernel<FMT, C_, F_, H_><<<grid, kQThreads, 0, st>>>(a);                                           \
    }                                                                                                              \
    break;
    switch (sel) {
        case 1: SLFP_QLAUNCH(true, false, false)
        case 2: SLFP_QLAUNCH(false, true, false)
        case 3: SLFP_QLAUNCH(true, true, false)
        case 4: SLFP_QLAUNCH(false, false, true)
        case 5: SLFP_QLAUNCH(true, false, true)
        case 6: SLFP_QLAUNCH(false, true, true)
        default: SLFP_QLAUNCH(true, true, true)
    }
#undef SLFP_QLAUNCH
    return check_launch("quantize_kernel");
}

// One element through the table encoder (enc_lut_index + shared-memory table), the generic encoder for the values
// outside its domain (NaN / Inf quotient, -0, dividends below 2^-119) or when K is outside the reciprocal's range.
template <int FMT>
__device__ __forceinline__ uint32_t encode_elem_lut(float x, const DivK& k, const uint8_t* s_enc) {
    const float q = div_k_fused(x, k);
    const uint32_t xb = __float_as_uint(x);
    const uint32_t probe = __funnelshift_l(xb, xb, 1) - 1u;
    if (!k.fast || probe < 0x08000000u - 1u || !(fabsf(q) < INFINITY)) return encode<FMT>(div_k(x, k));
    return (uint32_t)s_enc[enc_lut_index<FMT>(q, x)] | ((__float_as_uint(q) >> 24) & 0x80u);
}

// NHWC tensor whose channel count is not the physical (padded) one (ShuffleNetV2's 24 / 58 / 116 / 232 channels, C = 3
// inputs): thread = four consecutive padded channels of a pixel (c_phys % 4 == 0 and 4-byte aligned codes, else one),
// table encoder for the activation formats, one 32-bit store.  Was one generic encode<> and one byte store per thread.
template <int FMT>
__global__ void __launch_bounds__(256) quantize_pad_kernel(const float* __restrict__ x, size_t npix, int C, int Cp,
                                                           DivK k_div, uint8_t* __restrict__ codes, int vec4) {
    constexpr bool kLut = FMT == SLFP_FMT_SFP33 || FMT == SLFP_FMT_SLFP34_ACT;
    constexpr int FL = FMT == SLFP_FMT_SFP33 ? SLFP_FMT_SFP33 : SLFP_FMT_SLFP34_ACT;
    __shared__ uint8_t s_enc[kLut ? kEncLutBytes : 16];
    if (kLut) {
        for (int i = threadIdx.x; i < kEncLutBytes; i += 256) s_enc[i] = (uint8_t)enc_lut_entry<FL>((uint32_t)i);
        __syncthreads();
    }
    auto enc1 = [&](float v) -> uint32_t {
        if (kLut) return encode_elem_lut<FL>(v, k_div, s_enc);
        return encode<FMT>(div_k(v, k_div));
    };
    if (vec4) {
        const int q = Cp >> 2;
        const size_t total = npix * (size_t)q;
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
            const size_t pix = i / (size_t)q;
            const int c0 = (int)(i - pix * (size_t)q) * 4;
            const float* src = x + pix * (size_t)C + c0;
            uint32_t w = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c0 + j < C) w |= enc1(__ldg(src + j)) << (8 * j);
            *reinterpret_cast<uint32_t*>(codes + pix * (size_t)Cp + c0) = w;
        }
        return;
    }
    const size_t total = npix * (size_t)Cp;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const size_t pix = i / (size_t)Cp;
        const int c = (int)(i - pix * (size_t)Cp);
        codes[i] = (c < C) ? (uint8_t)enc1(x[pix * (size_t)C + c]) : (uint8_t)0;
    }
}

// NCHW float32 image -> NHWC codes with padded channels (the network input: C = 3 -> c_phys = 4).
// Thread = (pixel, group of 4 channels): plane reads are coalesced across the warp, the store is one
// 32-bit word per thread.
template <int FMT>
__global__ void __launch_bounds__(256) quantize_nchw_kernel(const float* __restrict__ x, int N, int C, size_t HW, int Cp,
                                                            DivK k_div, uint8_t* __restrict__ codes) {
    const int groups = Cp >> 2;
    const size_t total = (size_t)N * HW * groups;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const size_t pix = i % HW;
        const size_t ng = i / HW;
        const int g = (int)(ng % groups);
        const size_t n = ng / groups;
        uint32_t w = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = g * 4 + j;
            if (c < C) w |= encode<FMT>(div_k(__ldg(x + (n * C + c) * HW + pix), k_div)) << (8 * j);
        }
        *reinterpret_cast<uint32_t*>(codes + (n * HW + pix) * Cp + g * 4) = w;
    }
}

// The network-input case of the above (c_phys = 4, HW a multiple of 4, fewer than 2^31 pixel quads): a thread takes FOUR
// consecutive pixels - one float4 per plane (all C loads issued first), 32-bit index arithmetic (the generic kernel
// pays three 64-bit divisions per pixel), one 16-byte store of 4 pixels x 4 codes.
template <int FMT>
__global__ void __launch_bounds__(256) quantize_nchw_c4_kernel(const float* __restrict__ x, uint32_t total, int C, uint32_t HW4,
                                                               DivK k_div, uint8_t* __restrict__ codes) {
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < total; i += gridDim.x * 256u) {
        const uint32_t n = i / HW4, q = i - n * HW4;
        const float4* src = reinterpret_cast<const float4*>(x) + (size_t)n * C * HW4 + q;
        float4 v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = c < C ? __ldg(src + (size_t)c * HW4) : make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c < C) {
                w[0] |= encode<FMT>(div_k(v[c].x, k_div)) << (8 * c);
                w[1] |= encode<FMT>(div_k(v[c].y, k_div)) << (8 * c);
                w[2] |= encode<FMT>(div_k(v[c].z, k_div)) << (8 * c);
                w[3] |= encode<FMT>(div_k(v[c].w, k_div)) << (8 * c);
            }
        }
        reinterpret_cast<uint4*>(codes)[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// NCHW float32 image -> space-to-depth NHWC codes [N, H/2, W/2, Cp], channel (dy*2+dx)*C + c.  Thread = one
// folded pixel: 2 x float2 reads per plane (coalesced across the warp), one 16-byte store per 16 channels.
template <int FMT>
__global__ void __launch_bounds__(256) quantize_nchw_s2d_kernel(const float* __restrict__ x, int N, int C, int H, int W, int Cp,
                                                                DivK k_div, uint8_t* __restrict__ codes) {
    constexpr bool kLut = FMT == SLFP_FMT_SFP33 || FMT == SLFP_FMT_SLFP34_ACT;
    constexpr int FL = FMT == SLFP_FMT_SFP33 ? SLFP_FMT_SFP33 : SLFP_FMT_SLFP34_ACT;
    __shared__ uint8_t s_enc[kLut ? kEncLutBytes : 16];
    if (kLut) {
        for (int i = threadIdx.x; i < kEncLutBytes; i += 256) s_enc[i] = (uint8_t)enc_lut_entry<FL>((uint32_t)i);
        __syncthreads();
    }
    const int H2 = H >> 1, W2 = W >> 1;
    const size_t total = (size_t)N * H2 * W2;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const int x2 = (int)(i % W2), y2 = (int)((i / W2) % H2);
        const size_t n = i / ((size_t)W2 * H2);
        uint8_t* dst = codes + i * (size_t)Cp;
        if (Cp == 16 && 4 * C <= 16) {
            // the stem case (C <= 4): one float2 (both dx) per (channel, dy): consecutive threads read consecutive
            // float2 of an image row (coalesced), one 16-byte store per thread
            uint32_t wds[4] = {0u, 0u, 0u, 0u};
            if (C == 3) {
                // RGB input (every net's stem): all six float2 loads issued first, channel positions known at compile time
                // (the generic loop below indexes wds[] with run-time values: local memory, two loads in flight)
                float2 v[3][2];
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int dy = 0; dy < 2; ++dy)
                        v[c][dy] = __ldg(reinterpret_cast<const float2*>(x + ((n * 3 + c) * H + (2 * y2 + dy)) * (size_t)W) + x2);
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int dy = 0; dy < 2; ++dy) {
                        constexpr int kC = 3;
                        const int ch0 = (dy * 2) * kC + c, ch1 = (dy * 2 + 1) * kC + c;
                        uint32_t e0, e1;
                        if (kLut) {
                            e0 = encode_elem_lut<FL>(v[c][dy].x, k_div, s_enc);
                            e1 = encode_elem_lut<FL>(v[c][dy].y, k_div, s_enc);
                        } else {
                            e0 = encode<FMT>(div_k(v[c][dy].x, k_div));
                            e1 = encode<FMT>(div_k(v[c][dy].y, k_div));
                        }
                        wds[ch0 >> 2] |= e0 << (8 * (ch0 & 3));
                        wds[ch1 >> 2] |= e1 << (8 * (ch1 & 3));
                    }
                *reinterpret_cast<uint4*>(dst) = make_uint4(wds[0], wds[1], wds[2], wds[3]);
                continue;
            }
            for (int c = 0; c < C; ++c) {
#pragma unroll
                for (int dy = 0; dy < 2; ++dy) {
                    const float2 v = __ldg(reinterpret_cast<const float2*>(x + ((n * C + c) * H + (2 * y2 + dy)) * (size_t)W) + x2);
                    const int ch0 = (dy * 2) * C + c, ch1 = (dy * 2 + 1) * C + c;
                    if (kLut) {
                        wds[ch0 >> 2] |= encode_elem_lut<FL>(v.x, k_div, s_enc) << (8 * (ch0 & 3));
                        wds[ch1 >> 2] |= encode_elem_lut<FL>(v.y, k_div, s_enc) << (8 * (ch1 & 3));
                    } else {
                        wds[ch0 >> 2] |= encode<FMT>(div_k(v.x, k_div)) << (8 * (ch0 & 3));
                        wds[ch1 >> 2] |= encode<FMT>(div_k(v.y, k_div)) << (8 * (ch1 & 3));
                    }
                }
            }
            *reinterpret_cast<uint4*>(dst) = make_uint4(wds[0], wds[1], wds[2], wds[3]);
            continue;
        }
        for (int c0 = 0; c0 < Cp; c0 += 16) {
            uint32_t wds[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int ch = c0 + j;                       // folded channel = q*C + c, q = dy*2 + dx
                if (ch < 4 * C) {
                    const int q = ch / C, c = ch - q * C;
                    const float v = __ldg(x + ((n * C + c) * H + (2 * y2 + (q >> 1))) * (size_t)W + 2 * x2 + (q & 1));
                    wds[j >> 2] |= encode<FMT>(div_k(v, k_div)) << (8 * (j & 3));
                }
            }
            *reinterpret_cast<uint4*>(dst + c0) = make_uint4(wds[0], wds[1], wds[2], wds[3]);
        }
    }
}

// n / d == umulhi(n, mg) >> sh for n < 2^31 (d > 1): magic-number division by a run-time constant
static void magic_u32(uint32_t d, uint32_t& mg, uint32_t& sh) {
    mg = 0; sh = 0;
    if (d > 1) {
        uint32_t lg = 31 - __builtin_clz(d);
        if (d & (d - 1)) ++lg;
        const uint32_t p = 31 + lg;
        mg = (uint32_t)(((1ull << p) + d - 1) / d);
        sh = p - 32;
    }
}
__device__ __forceinline__ uint32_t div_magic(uint32_t n, uint32_t d, uint32_t mg, uint32_t sh) {
    return d == 1 ? n : (__umulhi(n, mg) >> sh);
}

// The stem case of the above (RGB image, c_phys = 16, w % 4 == 0, 16-byte aligned rows), HBM-bound form: a thread owns
// TWO horizontally adjacent folded pixels = one float4 per (channel, dy) image row piece (six independent 16-byte
// streaming loads in flight, 512 contiguous bytes per warp and load) and writes their 2 x 16 bytes of codes side by
// side; 32-bit index arithmetic with magic-number division (the generic kernel's three 64-bit divisions per pixel made
// it issue-bound at 0.34-0.46 of the HBM peak), table encoder with ONE group probe per thread for the rare general path.
// OUT16 (slfp_quantize_nchw_s2d_f16q): instead of the code bytes the kernel stores the float16 image of each code's value
// (SLFP_FMT_F16Q, 32 bytes per folded pixel) into a physically zero-padded [n, hp, wp, 16] tensor at (pad_top, pad_left).
template <int FL, bool OUT16 = false>
__global__ void __launch_bounds__(256) quantize_nchw_s2d_c3_kernel(const float* __restrict__ x, uint32_t total_pairs, int H, int W,
                                                                   uint32_t Wp, uint32_t mg_wp, uint32_t sh_wp, uint32_t H2,
                                                                   uint32_t mg_h2, uint32_t sh_h2, DivK k_div,
                                                                   uint8_t* __restrict__ codes, int pad_top = 0, int pad_left = 0,
                                                                   int hp = 0, int wp = 0) {
    __shared__ uint8_t s_enc[kEncLutBytes];
    __shared__ unsigned short s_dec[OUT16 ? 256 : 1];
    for (int i = threadIdx.x; i < kEncLutBytes; i += 256) s_enc[i] = (uint8_t)enc_lut_entry<FL>((uint32_t)i);
    if (OUT16) s_dec[threadIdx.x] = __half_as_ushort(__float2half_rn(decode<FL == SLFP_FMT_SFP33>((uint32_t)threadIdx.x, c_pow2frac)));
    __syncthreads();
    const size_t plane = (size_t)H * W;
    for (uint32_t j = blockIdx.x * 256u + threadIdx.x; j < total_pairs; j += gridDim.x * 256u) {
        const uint32_t row = div_magic(j, Wp, mg_wp, sh_wp), px = j - row * Wp;     // row = n * H2 + y2
        const uint32_t n = div_magic(row, H2, mg_h2, sh_h2), y2 = row - n * H2;
        const float* src = x + (size_t)n * 3 * plane + (size_t)(2 * y2) * W + 4 * px;
        float4 v[3][2];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) v[c][dy] = ldg_stream(reinterpret_cast<const float4*>(src + c * plane + dy * W));
        // element e of pixel p: channel ch = (dy * 2 + dx) * 3 + c  <-  v[c][dy].{x,y | z,w}[dx]
        float xs[2][12];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                xs[0][(dy * 2 + 0) * 3 + c] = v[c][dy].x; xs[0][(dy * 2 + 1) * 3 + c] = v[c][dy].y;
                xs[1][(dy * 2 + 0) * 3 + c] = v[c][dy].z; xs[1][(dy * 2 + 1) * 3 + c] = v[c][dy].w;
            }
        float qv[2][12];
        float nan_probe = 0.0f;
        uint32_t xmin = 0xffffffffu;
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int e = 0; e < 12; ++e) {
                qv[p][e] = div_k_fused(xs[p][e], k_div);
                nan_probe = fmaf(qv[p][e], 0.0f, nan_probe);
                const uint32_t ax1 = __funnelshift_l(__float_as_uint(xs[p][e]), __float_as_uint(xs[p][e]), 1) - 1u;
                xmin = ax1 < xmin ? ax1 : xmin;
            }
        const bool general = !k_div.fast || nan_probe != nan_probe || xmin < 0x08000000u - 1u;
        uint32_t w[2][4];
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            uint32_t c8[12];
            if (!general) {
#pragma unroll
                for (int e = 0; e < 12; ++e)
                    c8[e] = (uint32_t)s_enc[enc_lut_index<FL>(qv[p][e], xs[p][e])] | ((__float_as_uint(qv[p][e]) >> 24) & 0x80u);
            } else {
#pragma unroll
                for (int e = 0; e < 12; ++e) c8[e] = encode<FL>(div_k(xs[p][e], k_div));
            }
            if (OUT16) {
                uint32_t hw[6];
#pragma unroll
                for (int e = 0; e < 6; ++e) hw[e] = (uint32_t)s_dec[c8[2 * e]] | ((uint32_t)s_dec[c8[2 * e + 1]] << 16);
                uint4* d16 = reinterpret_cast<uint4*>(codes + (((size_t)n * hp + y2 + pad_top) * wp + 2 * px + p + pad_left) * 32);
                d16[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                d16[1] = make_uint4(hw[4], hw[5], 0u, 0u);           // pad channels 12..15: 0.0
                continue;
            }
#pragma unroll
            for (int g = 0; g < 3; ++g)
                w[p][g] = __byte_perm(__byte_perm(c8[4 * g], c8[4 * g + 1], 0x0040), __byte_perm(c8[4 * g + 2], c8[4 * g + 3], 0x0040), 0x5410);
            w[p][3] = 0u;                                        // pad channels 12..15: code 0 = exact zero
        }
        if (OUT16) continue;
        uint4* dst = reinterpret_cast<uint4*>(codes + (size_t)j * 32);
        dst[0] = make_uint4(w[0][0], w[0][1], w[0][2], w[0][3]);
        dst[1] = make_uint4(w[1][0], w[1][1], w[1][2], w[1][3]);
    }
}

// ---- gather + quantize (split / cat / channel_shuffle as an index map) -------------------------------------------
// thread = (four consecutive output channels, a lane of pixels): its four table entries (source pointer, pixel
// stride, channel) stay in registers while it walks down the pixels of its CTA, so the inner loop is four 2-byte
// gathers (consecutive channels of a run come from consecutive addresses of one source tensor: a warp's loads fall into
// a few 128-byte lines) and one 32-bit store per pixel; a warp's stores cover 128 contiguous bytes.
// The quantizer itself is ONE byte look-up per element: the sources are float16 tensors of post-layerout, post-ReLU
// values (non-negative, <= 248), so every possible input bit pattern below 0x5C00 gets its code from a table the CTA
// builds once with the exact encoder (IEEE division by K, round-half-even) - 23.5 KB of shared memory, persistent
// CTAs.  Anything else (negative, > 248, Inf / NaN) takes the exact encoder directly.
// Pad channels (>= c) get code 0.  c_phys <= 1024.
constexpr int kGatherPix = 64;                                   // pixels per thread and CTA pass
constexpr int kGatherLut = 0x5C00;                               // float16 bit patterns of [0, 256)
template <int FMT, bool E4M3 = false>
__global__ void __launch_bounds__(256) gather_quantize_kernel(const SlfpGatherChan* __restrict__ table, size_t npix, int C, int Cp,
                                                              DivK k_div, uint8_t* __restrict__ codes) {
    __shared__ uint8_t s_code[kGatherLut];
    for (int i = threadIdx.x; i < kGatherLut; i += 256) {
        const float v = __half2float(__ushort_as_half((unsigned short)i));
        uint32_t code = encode<FMT>(div_k(v, k_div));
        if (E4M3) code = sfp33_code_to_e4m3(code);                // the exact SFP<3,3> code, re-spelled as the e4m3 byte
        s_code[i] = (uint8_t)code;
    }
    __syncthreads();
    const int cq = Cp >> 2, lanes = 256 / cq;                     // channel quads per pixel; pixels a CTA handles at once
    const int q = (int)threadIdx.x % cq, pl = (int)threadIdx.x / cq;
    if (pl >= lanes) return;                                      // c_phys / 4 does not divide 256: the last threads idle
    const unsigned short* src[4];
    size_t stride[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = q * 4 + j;
        src[j] = nullptr; stride[j] = 0;
        if (c < C) {
            const SlfpGatherChan t = table[c];
            src[j] = reinterpret_cast<const unsigned short*>(t.src) + t.ch;
            stride[j] = (size_t)t.stride;
        }
    }
    const size_t per_cta = (size_t)lanes * kGatherPix;
    for (size_t base = (size_t)blockIdx.x * per_cta; base < npix; base += (size_t)gridDim.x * per_cta) {
#pragma unroll 4
        for (int i = 0; i < kGatherPix; ++i) {
            const size_t pix = base + (size_t)i * lanes + pl;
            if (pix >= npix) break;
            uint32_t word = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (src[j] != nullptr) {
                    const uint32_t hb = __ldg(src[j] + pix * stride[j]);
                    uint32_t code;
                    if (hb < (uint32_t)kGatherLut) {
                        code = s_code[hb];
                    } else {
                        code = encode<FMT>(div_k(__half2float(__ushort_as_half((unsigned short)hb)), k_div));
                        if (E4M3) code = sfp33_code_to_e4m3(code);
                    }
                    word |= code << (8 * j);
                }
            }
            *reinterpret_cast<uint32_t*>(codes + pix * (size_t)Cp + q * 4) = word;
        }
    }
}

// ---- run-based gather + quantize (see include/slfp_b200.h) ----------------------------------------------------------
// Two phases per CTA tile, both free of shared-memory bank conflicts (the byte-scatter form spent 80 % of its
// shared-memory wavefronts on conflict replays, profiles/r02_config5.md):
//   1. per run, the 16-byte aligned chunks (8 channels) that cover it are copied for every pixel of the tile into a
//      staging region [pixel][chunks] - sector-aligned 128-bit loads, four in flight per thread, 128-bit stores;
//   2. a thread owns one 32-bit word of the output (four consecutive logical channels of a pixel), reads its four
//      float16 inputs from the staging regions through a per-channel (offset, pitch) table - lanes of a warp walk
//      consecutive channels of a run, i.e. consecutive banks - encodes them and writes the word straight to global
//      memory (a warp writes 128 contiguous bytes).
// e4m3 output: clamps + cvt on the FMA / ALU pipes (reciprocal multiply like the other fused encoders); the other
// formats: one byte look-up in a value -> code table the CTA builds with the exact encoder (sources are non-negative
// float16 <= 248: post-layerout, post-ReLU tensors; anything else takes the exact encoder directly).
constexpr int kRunStageBytes = 32768;
constexpr int kRunMaxC = 1024;
template <int FMT, bool E4M3>
__global__ void __launch_bounds__(256) gather_runs_kernel(const SlfpGatherRun* __restrict__ runs, int n_runs, size_t npix, int C, int Cp,
                                                          int tile_pix, DivK k_div, uint8_t* __restrict__ codes) {
    __shared__ uint8_t s_code[E4M3 ? 16 : kGatherLut];
    __shared__ __align__(16) uint8_t s_stage[E4M3 ? kRunStageBytes : kRunStageBytes / 2];     // (the table variants: 48 KB static limit)
    __shared__ uint32_t s_chan[kRunMaxC];                        // per logical channel: staging offset (low 20 bits) | row pitch / 16 (high 12)
    if (!E4M3) {
        for (int i = threadIdx.x; i < kGatherLut; i += 256)
            s_code[i] = (uint8_t)encode<FMT>(div_k(__half2float(__ushort_as_half((unsigned short)i)), k_div));
    }
    {   // staging layout: run r occupies [tile_pix][nck_r * 16 B] starting at off_r; channel table
        int off = 0;
        for (int r = 0; r < n_runs; ++r) {
            const SlfpGatherRun run = runs[r];
            const int c_lo = run.ch0 & ~7, nck = (((run.ch0 + run.len + 7) & ~7) - c_lo) >> 3;
            for (int i = threadIdx.x; i < run.len; i += 256)
                s_chan[run.dst_start + i * run.dst_step] = (uint32_t)(off + (run.ch0 - c_lo + i) * 2) | ((uint32_t)nck << 20);
            off += tile_pix * nck * 16;
        }
    }
    __syncthreads();
    const float rk = k_div.rk;
    const int cq = Cp >> 2;
    const bool fixed_quad = cq <= 256 && (256 % cq) == 0;
    const uint32_t my_lanes = fixed_quad ? 256u / (uint32_t)cq : 1u, my_w = threadIdx.x % (uint32_t)cq, my_pl = threadIdx.x / (uint32_t)cq;
    uint32_t my_off[4] = {0u, 0u, 0u, 0u}, my_pitch[4] = {0u, 0u, 0u, 0u};        // pitch 0 = pad channel
    if (fixed_quad) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int c = (int)my_w * 4 + jj;
            if (c < C) { const uint32_t e = s_chan[c]; my_off[jj] = e & 0xfffffu; my_pitch[jj] = (e >> 20) << 4; }
        }
    }
    for (size_t p0 = (size_t)blockIdx.x * tile_pix; p0 < npix; p0 += (size_t)gridDim.x * tile_pix) {
        const int np = (int)(npix - p0 < (size_t)tile_pix ? npix - p0 : (size_t)tile_pix);
        // ---- phase 1: stage the covering chunks of every run ----------------------------------------------------
        int off = 0;
        for (int r = 0; r < n_runs; ++r) {
            const SlfpGatherRun run = runs[r];
            const int c_lo = run.ch0 & ~7, nck = (((run.ch0 + run.len + 7) & ~7) - c_lo) >> 3;
            const unsigned short* src = reinterpret_cast<const unsigned short*>(run.src) + c_lo;
            const uint32_t total = (uint32_t)np * (uint32_t)nck;
            constexpr int kFly = 4;
            for (uint32_t base = threadIdx.x; base < total; base += 256 * kFly) {
                uint4 v[kFly];
#pragma unroll
                for (int u = 0; u < kFly; ++u) {
                    const uint32_t idx = base + (uint32_t)u * 256u;
                    const uint32_t pp = nck == 1 ? idx : (__umulhi(idx, run.magic) >> run.shift), ck = idx - pp * (uint32_t)nck;
                    v[u] = make_uint4(0u, 0u, 0u, 0u);
                    if (idx < total) v[u] = ldg_stream_u4(reinterpret_cast<const uint4*>(src + (p0 + pp) * (size_t)run.stride) + ck);
                }
#pragma unroll
                for (int u = 0; u < kFly; ++u) {
                    const uint32_t idx = base + (uint32_t)u * 256u;
                    if (idx < total) *reinterpret_cast<uint4*>(s_stage + off + idx * 16u) = v[u];      // [pixel][chunk]: idx itself
                }
            }
            off += tile_pix * nck * 16;
        }
        __syncthreads();
        // ---- phase 2: one output word per thread ---------------------------------------------------------------------
        // (256 % cq == 0: a thread keeps its channel quad for the whole kernel - its four staging offsets and pitches
        // were loaded once, the loop is four LDS.U16 + the encoder + one store per pixel)
        auto emit = [&](uint32_t pp, uint32_t w, const uint32_t (&eo)[4], const uint32_t (&ep)[4]) {
            float f[4];
            uint32_t hb[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                hb[jj] = ep[jj] ? (uint32_t)*reinterpret_cast<const unsigned short*>(s_stage + eo[jj] + pp * ep[jj]) : 0u;
                f[jj] = __half2float(__ushort_as_half((unsigned short)hb[jj]));
            }
            uint32_t word;
            if (E4M3) {
                word = encode_e4m3x2_relu(f[0] * rk, f[1] * rk) | (encode_e4m3x2_relu(f[2] * rk, f[3] * rk) << 16);
            } else {
                word = 0u;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const uint32_t code = hb[jj] < (uint32_t)kGatherLut ? (uint32_t)s_code[hb[jj]] : encode<FMT>(div_k(f[jj], k_div));
                    word |= (ep[jj] ? code : 0u) << (8 * jj);
                }
            }
            reinterpret_cast<uint32_t*>(codes + (p0 + pp) * (size_t)Cp)[w] = word;
        };
        if (fixed_quad) {
            for (uint32_t pp = my_pl; pp < (uint32_t)np; pp += my_lanes) emit(pp, my_w, my_off, my_pitch);
        } else {
            const uint32_t words = (uint32_t)np * (uint32_t)cq;
            for (uint32_t wi = threadIdx.x; wi < words; wi += 256) {
                const uint32_t pp = wi / (uint32_t)cq, w = wi - pp * (uint32_t)cq;
                uint32_t eo[4], ep[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int c = (int)w * 4 + jj;
                    const uint32_t e = c < C ? s_chan[c] : 0u;
                    eo[jj] = e & 0xfffffu; ep[jj] = (e >> 20) << 4;
                }
                emit(pp, w, eo, ep);
            }
        }
        __syncthreads();
    }
}

// ---- de-quantize ------------------------------------------------------------------------------
template <bool SFP33>
__global__ void __launch_bounds__(256) dequantize_kernel(const uint8_t* __restrict__ codes, size_t n,
                                                         float* __restrict__ out) {
    __shared__ uint32_t s_tab[16];
    if (threadIdx.x < 16) s_tab[threadIdx.x] = c_pow2frac[threadIdx.x];
    __syncthreads();
    const bool vec = (((uintptr_t)codes & 3u) | ((uintptr_t)out & 15u)) == 0;
    const size_t n4 = vec ? n / 4 : 0;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
        const uint32_t w = reinterpret_cast<const uint32_t*>(codes)[i];
        float4 o;
        o.x = decode<SFP33>(w & 0xffu, s_tab);
        o.y = decode<SFP33>((w >> 8) & 0xffu, s_tab);
        o.z = decode<SFP33>((w >> 16) & 0xffu, s_tab);
        o.w = decode<SFP33>(w >> 24, s_tab);
        reinterpret_cast<float4*>(out)[i] = o;
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
        out[i] = decode<SFP33>(codes[i], s_tab);
}

// ---- abs-max ------------------------------------------------------------------------------------
// warp shuffles, then a block reduction in shared memory, then one atomicMax per CTA on the bit
// pattern (non-negative floats order like unsigned integers).
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, size_t n,
                                                     unsigned int* __restrict__ out) {
    float m = 0.0f;
    const bool vec = ((uintptr_t)x & 15u) == 0;
    const size_t n4 = vec ? n / 4 : 0;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
        const float4 v = ldg_stream(reinterpret_cast<const float4*>(x) + i);
        m = fmaxf(fmaxf(m, fabsf(v.x)), fmaxf(fabsf(v.y), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
        m = fmaxf(m, fabsf(x[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ float s_m[8];
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < 8 ? s_m[threadIdx.x] : 0.0f;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) atomicMax(out, __float_as_uint(m));
    }
}

// ---- weight preparation ---------------------------------------------------------------------------
// One thread per KRSC destination element (k, r, s, c_phys); weights are small (<= 2.4 M / layer).
struct WPrepArgs {
    const float* w;
    long long so, sc, sr, ss;
    int K, C, Cp, R, S;
    size_t pitch;  // elements per output-channel row of the KRSC operand (>= R*S*Cp, multiple of 64)
    float kw;
    __half* w_f16;
    uint8_t* w_codes;
    float* w_fakeq;  // OIHW contiguous
    // placement of the row inside a wider (concatenated-K) operand and an optional per-output-channel factor
    // applied to the float16 operand only (fused block tails, see slfp_conv2d_fwd_dual)
    size_t out_pitch, out_off;
    size_t lo_off;                                           // != 0: the lo half of the split operand goes to o + lo_off
    int e4m3;                                                // SLFP_CONV_E4M3_OPERANDS: w_f16 receives e4m3 BYTES (SFP<3,3> only)
    const float* row_scale;
    uint32_t mg_pitch, sh_pitch, mg_cp, sh_cp, mg_s, sh_s;   // n / d == umulhi(n, mg) >> sh for n < 2^31 (d > 1)
    DivK dk;                                                 // kw with its reciprocal: exact x / kw without the division sequence
    int vec8_ok;                                             // rows, offsets and pointers allow the 8-element vector path
};

// shared-memory tables of the weight-preparation kernels: bucket table of the SLFP weight encoder, code -> float32
struct WPrepTables {
    uint2 bucket[32];
    float dec[256];
};
template <int FMT>
__device__ __forceinline__ void wprep_tables_init(WPrepTables& t, const uint32_t* s_tab) {
    if (FMT < 0) return;
    if (threadIdx.x < 32) {
        uint32_t cnt, thr;
        wgt_bucket_entry(threadIdx.x, cnt, thr);
        t.bucket[threadIdx.x] = make_uint2(cnt, thr);
    }
    t.dec[threadIdx.x] = decode<FMT == SLFP_FMT_SFP33>(threadIdx.x, s_tab);      // 256 threads
}

template <int FMT>
__device__ __forceinline__ void wprep_element(const WPrepArgs& a, size_t i, const WPrepTables& tb) {
    // a prepared tensor has < 2^32 elements (checked on the host): 32-bit index arithmetic
    const uint32_t i32 = (uint32_t)i, pitch = (uint32_t)a.pitch;
    const int k = (int)div_magic(i32, pitch, a.mg_pitch, a.sh_pitch);
    const uint32_t j = i32 - (uint32_t)k * pitch;
    const int rs = (int)div_magic(j, (uint32_t)a.Cp, a.mg_cp, a.sh_cp);
    const int c = (int)(j - (uint32_t)rs * (uint32_t)a.Cp);
    uint32_t code = 0;
    float fq = 0.0f;
    if (c < a.C && rs < a.R * a.S) {
        const int r = (int)div_magic((uint32_t)rs, (uint32_t)a.S, a.mg_s, a.sh_s), s = rs - r * a.S;
        const float x = a.w[k * a.so + c * a.sc + r * a.sr + s * a.ss];
        const float v = div_k(x, a.dk);                      // == IEEE x / kw
        if (FMT < 0) {
            fq = v;
        } else {
            code = FMT == SLFP_FMT_SLFP34_WGT ? encode_wgt_bucket(v, tb.bucket) : encode<FMT < 0 ? 0 : FMT>(v);
            fq = tb.dec[code];
        }
        if (a.w_fakeq) a.w_fakeq[(((size_t)k * a.C + c) * a.R + r) * a.S + s] = fq;
    }
    const size_t o = (size_t)k * a.out_pitch + a.out_off + j;
    if (a.w_f16 && a.e4m3) {
        reinterpret_cast<uint8_t*>(a.w_f16)[o] = (uint8_t)sfp33_code_to_e4m3(code);
    } else if (a.w_f16) {
        const float val = a.row_scale ? fq * __ldg(a.row_scale + k) : fq;
        const __half hi = __float2half_rn(val);
        a.w_f16[o] = hi;
        if (a.lo_off) a.w_f16[o + a.lo_off] = __float2half_rn(val - __half2float(hi));
    }
    if (a.w_codes) a.w_codes[o] = (uint8_t)code;
}

// Eight consecutive destination elements (same k and tap, eight channels: c_phys % 8 == 0) per thread: the index
// arithmetic - three magic divisions and the 64-bit strided source address, ~150 of the ~175 instructions the
// per-element form spent (ncu: the batched kernel was issue-bound at 85 % SM throughput, not memory-bound) - is done
// once, the eight gathers are independent, and the float16 / code rows leave as one 16-byte / 8-byte store.  Padding
// channels and K-padding taps read as 0, which every format encodes as code 0 / value 0.
template <int FMT>
__device__ __forceinline__ void wprep_vec8(const WPrepArgs& a, size_t i, const WPrepTables& tb) {
    const uint32_t i32 = (uint32_t)i, pitch = (uint32_t)a.pitch;
    const int k = (int)div_magic(i32, pitch, a.mg_pitch, a.sh_pitch);
    const uint32_t j = i32 - (uint32_t)k * pitch;
    const int rs = (int)div_magic(j, (uint32_t)a.Cp, a.mg_cp, a.sh_cp);
    const int c0 = (int)(j - (uint32_t)rs * (uint32_t)a.Cp);
    const bool tap_in = rs < a.R * a.S;
    const int r = (int)div_magic((uint32_t)rs, (uint32_t)a.S, a.mg_s, a.sh_s), s_ = rs - r * a.S;
    const float* src = a.w + (k * a.so + r * a.sr + s_ * a.ss);
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = (tap_in && c0 + e < a.C) ? __ldg(src + (c0 + e) * a.sc) : 0.0f;
    const float rsk = a.row_scale ? __ldg(a.row_scale + k) : 1.0f;
    uint32_t hw[4], lw[4], cw[2] = {0u, 0u};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float v = div_k(x[e], a.dk);                   // == IEEE x / kw
        uint32_t code = 0;
        float fq = v;
        if (FMT >= 0) {
            code = FMT == SLFP_FMT_SLFP34_WGT ? encode_wgt_bucket(v, tb.bucket) : encode<FMT < 0 ? 0 : FMT>(v);
            fq = tb.dec[code];
        }
        if (a.w_fakeq && tap_in && c0 + e < a.C) a.w_fakeq[(((size_t)k * a.C + c0 + e) * a.R + r) * a.S + s_] = fq;
        const float val = a.row_scale ? fq * rsk : fq;
        const __half hh = __float2half_rn(val);
        const uint32_t h = (uint32_t)__half_as_ushort(hh);
        const uint32_t l = (uint32_t)__half_as_ushort(__float2half_rn(val - __half2float(hh)));   // split-operand lo part
        if (e & 1) { hw[e >> 1] |= h << 16; lw[e >> 1] |= l << 16; } else { hw[e >> 1] = h; lw[e >> 1] = l; }
        cw[e >> 2] |= code << (8 * (e & 3));
    }
    const size_t o = (size_t)k * a.out_pitch + a.out_off + j;
    if (a.w_f16 && a.e4m3) {
        uint32_t e0 = 0u, e1 = 0u;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            e0 |= sfp33_code_to_e4m3((cw[0] >> (8 * e)) & 0xffu) << (8 * e);
            e1 |= sfp33_code_to_e4m3((cw[1] >> (8 * e)) & 0xffu) << (8 * e);
        }
        *reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(a.w_f16) + o) = make_uint2(e0, e1);
    } else
    if (a.w_f16) *reinterpret_cast<uint4*>(a.w_f16 + o) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    if (a.w_f16 && a.lo_off) *reinterpret_cast<uint4*>(a.w_f16 + o + a.lo_off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    if (a.w_codes) *reinterpret_cast<uint2*>(a.w_codes + o) = make_uint2(cw[0], cw[1]);
}

template <int FMT>
__global__ void __launch_bounds__(256) wprep_kernel(WPrepArgs a) {
    __shared__ uint32_t s_tab[16];
    __shared__ WPrepTables tb;
    if (threadIdx.x < 16) s_tab[threadIdx.x] = c_pow2frac[threadIdx.x];
    __syncthreads();
    wprep_tables_init<FMT>(tb, s_tab);
    __syncthreads();
    const size_t total = (size_t)a.K * a.pitch;
    if (a.vec8_ok) {
        for (size_t i = ((size_t)blockIdx.x * 256 + threadIdx.x) * 8; i < total; i += (size_t)gridDim.x * 256 * 8)
            wprep_vec8<FMT>(a, i, tb);
        return;
    }
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256)
        wprep_element<FMT>(a, i, tb);
}

// Every layer of a network in ONE launch (the reference re-quantizes all weights on every forward,
// conv2d_func.py:22; 54 separate launches cost more than the 25 M weights themselves).  The table travels as
// a kernel parameter; block b works on kWBatchChunk consecutive elements of the tensor whose block range holds b.
constexpr int kWBatchMax = 160;
constexpr int kWBatchChunk = 2048;
struct WPrepBatch {
    int n;
    unsigned blk_end[kWBatchMax];      // exclusive prefix sums of blocks per tensor
    WPrepArgs a[kWBatchMax];
};

static_assert(sizeof(WPrepBatch) <= 32000, "kernel parameter space");

template <int FMT>
__global__ void __launch_bounds__(256) wprep_batch_kernel(const __grid_constant__ WPrepBatch b) {
    __shared__ uint32_t s_tab[16];
    __shared__ WPrepTables tb;
    if (threadIdx.x < 16) s_tab[threadIdx.x] = c_pow2frac[threadIdx.x];
    __syncthreads();
    wprep_tables_init<FMT>(tb, s_tab);
    __syncthreads();
    int lo = 0, hi = b.n - 1;                                 // first tensor with blk_end > blockIdx.x
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (b.blk_end[mid] > blockIdx.x) hi = mid; else lo = mid + 1;
    }
    const WPrepArgs& a = b.a[lo];
    const unsigned first = lo ? b.blk_end[lo - 1] : 0u;
    const size_t total = (size_t)a.K * a.pitch;
    const size_t base = (size_t)(blockIdx.x - first) * kWBatchChunk;
    // 16-byte / 8-byte vector stores need 8-element alignment of every output row and pointer (checked on the host
    // side of the jobs call: vec8_ok), and eight consecutive elements inside one tap need c_phys % 8 == 0
    if (a.vec8_ok) {
        static_assert(kWBatchChunk % (256 * 8) == 0, "whole 8-element vectors per thread");
#pragma unroll 1
        for (int j = 0; j < kWBatchChunk / (256 * 8); ++j) {
            const size_t i = base + ((size_t)j * 256 + threadIdx.x) * 8;
            if (i < total) wprep_vec8<FMT>(a, i, tb);
        }
        return;
    }
#pragma unroll 2
    for (int j = 0; j < kWBatchChunk / 256; ++j) {
        const size_t i = base + (size_t)j * 256 + threadIdx.x;
        if (i < total) wprep_element<FMT>(a, i, tb);
    }
}

// ---- weight preparation, row-staged form ------------------------------------------------------------------------
// The gather form above spends ~105 SASS instructions per weight (ncu, profiles/r04_final.md: 845 per 8-element thread
// of which 214 IMAD / 119 ISETP / 74 LDC are index arithmetic, strided 64-bit source addresses and kernel-parameter
// reads) and was issue-bound at 85 % of the issue slots, 0.16 of the HBM peak.  For the common case - a contiguous OIHW
// tensor, whole float4 groups per filter row - a CTA instead takes a CONTIGUOUS run of source floats (several whole
// filter rows, or a channel range of one long row), reads it with coalesced 16-byte loads in source order, encodes each
// element once, scatters value / code into a shared-memory image of the KRSC destination rows ([tap][channel]) and
// writes that image out with 16-byte stores.  The (channel, tap) split of the source index is one magic division per
// float4 plus carries; the parameters are read once per CTA.
constexpr int kWFastCap = 4608;                   // source elements per CTA (512 channels x 9 taps = one ResNet-50 / VGG row)
constexpr int kWFastMaxTaps = 32;                 // staging rows
struct WFastJob {
    const float* w;                               // [K][C][R*S] contiguous, 16-byte aligned
    __half* w_f16;                                // float16 operand or (e4m3 != 0) e4m3 bytes; may be null
    uint8_t* w_codes;                             // may be null
    const float* row_scale;
    size_t out_pitch, out_off;
    uint32_t K, C, Cp, RS, pitch;                 // pitch: elements per destination row that must be written (>= RS * Cp)
    uint32_t nrows, cc, n_cchunks;                // rows per CTA (C == Cp, whole rows) | channels per CTA of a split row
    uint32_t mg_rs, sh_rs, mg_cch, sh_cch, mg_c, sh_c;
    DivK dk;
    int e4m3;
};
constexpr int kWFastMax = 200;
struct WFastBatch {
    int n;
    unsigned blk_end[kWFastMax];
    WFastJob j[kWFastMax];
};
static_assert(sizeof(WFastBatch) <= 32000, "kernel parameter space");

// OUT: 0 float16 operand, 1 code bytes, 2 e4m3 bytes (through w_f16), 3 float16 operand and code bytes
template <int FMT, int OUT>
__global__ void __launch_bounds__(256) wprep_rows_kernel(const __grid_constant__ WFastBatch b) {
    constexpr bool kH = OUT == 0 || OUT == 3, kB = OUT != 0;          // float16 image / byte image
    __shared__ uint32_t s_tab[16];
    __shared__ WPrepTables tb;
    // staging images of the destination rows: [tap][channel slot], row pitch = slots + 8 | 16 (16-byte rows, banks skewed per tap)
    __shared__ __align__(16) __half s_h[kH ? kWFastCap + kWFastMaxTaps * 8 : 8];
    __shared__ __align__(16) uint8_t s_c[kB ? kWFastCap + kWFastMaxTaps * 16 : 16];
    constexpr bool kLut = FMT == SLFP_FMT_SLFP34_WGT;
    __shared__ WgtLutEntry s_lut[kLut ? kWgtLutEntries : 1];
    // persistent CTAs: the tables are built once, then the CTA walks over pieces blockIdx.x, + gridDim.x, ...
    if (threadIdx.x < 16) s_tab[threadIdx.x] = c_pow2frac[threadIdx.x];
    __syncthreads();
    wprep_tables_init<FMT>(tb, s_tab);
    if (kLut) for (uint32_t i = threadIdx.x; i < kWgtLutEntries; i += 256) s_lut[i] = wgt_lut_entry(i, s_tab);
    const uint32_t n_pieces = b.blk_end[b.n - 1];
#pragma unroll 1
    for (uint32_t piece = blockIdx.x; piece < n_pieces; piece += gridDim.x) {
    int lo = 0, hi = b.n - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (b.blk_end[mid] > piece) hi = mid; else lo = mid + 1;
    }
    const WFastJob& a = b.j[lo];
    const uint32_t q = piece - (lo ? b.blk_end[lo - 1] : 0u);
    const uint32_t C = a.C, Cp = a.Cp, RS = a.RS;
    // this CTA's piece: rows [k0, k0 + nr) x channels [c_lo, c_lo + cn); slots = channel positions it owns per tap
    uint32_t k0, nr, c_lo, cn, slots;
    if (a.n_cchunks > 1) {
        k0 = div_magic(q, a.n_cchunks, a.mg_cch, a.sh_cch);
        const uint32_t ci = q - k0 * a.n_cchunks;
        nr = 1; c_lo = ci * a.cc;
        cn = c_lo < C ? min(a.cc, C - c_lo) : 0u;
        slots = (ci + 1 == a.n_cchunks) ? Cp - c_lo : a.cc;               // the last piece also owns the padding channels
    } else {
        k0 = q * a.nrows; nr = min(a.nrows, a.K - k0); c_lo = 0; cn = C;
        slots = nr > 1 ? nr * C : Cp;                                     // nrows > 1 only when C == Cp
    }
    // the first 16-byte load of the piece is issued before the staging set-up; the loop below keeps one load ahead
    const float* src = a.w + ((size_t)k0 * C + c_lo) * RS;
    const uint32_t n_src = nr * cn * RS;                                  // multiple of 4, <= kWFastCap (host)
    uint32_t u0 = threadIdx.x * 4;
    float4 cur = u0 < n_src ? ldg_stream(reinterpret_cast<const float4*>(src + u0)) : make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t ph = slots + 8, pc = slots + 16;
    if (nr == 1 && slots != cn) {                                         // padding channels read as 0 / code 0
        if (kH) for (uint32_t i = threadIdx.x; i < RS * ph / 2; i += 256) reinterpret_cast<uint32_t*>(s_h)[i] = 0u;
        if (kB) for (uint32_t i = threadIdx.x; i < RS * pc / 4; i += 256) reinterpret_cast<uint32_t*>(s_c)[i] = 0u;
    }
    __syncthreads();
    const DivK dk = a.dk;
    const float* rsc = a.row_scale;
    const uint32_t mg_rs = a.mg_rs, sh_rs = a.sh_rs;
#pragma unroll 1
    for (; u0 < n_src; u0 += 256 * 4) {
        const uint32_t un = u0 + 256 * 4;
        const float4 nxt = un < n_src ? ldg_stream(reinterpret_cast<const float4*>(src + un)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float x[4] = {cur.x, cur.y, cur.z, cur.w};
        cur = nxt;
        uint32_t ch = div_magic(u0, RS, mg_rs, sh_rs);                    // channel slot (row_local * C + c when nr > 1)
        uint32_t rs = u0 - ch * RS;
        uint32_t at_h = rs * ph + ch, at_c = rs * pc + ch;                // staging positions, advanced with (rs, ch)
        // magnitudes as integers: NaN and Inf order above every finite value
        const uint32_t a0 = f2u(x[0]) & 0x7fffffffu, a1 = f2u(x[1]) & 0x7fffffffu, a2 = f2u(x[2]) & 0x7fffffffu, a3 = f2u(x[3]) & 0x7fffffffu;
        const uint32_t amax = max(max(a0, a1), max(a2, a3)), amin = min(min(a0, a1), min(a2, a3));
        const bool fast = dk.fast && amin >= 0x21800000u && amax < 0x5d800000u;       // 2^-60 <= |x| < 2^60: reciprocal sequence
        if (kLut && fast && rsc == nullptr) {
            // the common group: finite non-zero quotients, one table look-up per weight
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float q0 = x[e] * dk.rk;
                const float r = fmaf(-q0, dk.k, x[e]);
                const uint32_t qb = f2u(fmaf(r, dk.rk, q0)), qa = qb & 0x7fffffffu;   // == IEEE x / kw
                uint32_t h16;
                const uint32_t code = encode_wgt_lut(qb, s_lut[wgt_lut_index(qa)], h16);
                if (kH) s_h[at_h] = __ushort_as_half((unsigned short)h16);
                if (kB) s_c[at_c] = (uint8_t)code;
                at_h += ph; at_c += pc;
                if (++rs == RS) { rs = 0; ++ch; at_h = ch; at_c = ch; }
            }
        } else {
            float rsk = 1.0f;
            if (kH && rsc) rsk = __ldg(rsc + k0 + (nr > 1 ? div_magic(ch, C, a.mg_c, a.sh_c) : 0u));
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float qv;
                if (fast) {
                    const float q0 = x[e] * dk.rk;
                    const float r = fmaf(-q0, dk.k, x[e]);
                    qv = fmaf(r, dk.rk, q0);
                } else {
                    qv = div_rn(x[e], dk.k);
                }
                const uint32_t code = FMT == SLFP_FMT_SLFP34_WGT ? encode_wgt_bucket(qv, tb.bucket) : encode<FMT>(qv);
                if (kH) {
                    float val = tb.dec[code];
                    if (rsc) val *= rsk;
                    s_h[at_h] = __float2half_rn(val);
                }
                if (kB) s_c[at_c] = (uint8_t)(OUT == 2 ? sfp33_code_to_e4m3(code) : code);
                at_h += ph; at_c += pc;
                if (++rs == RS) {
                    rs = 0; ++ch; at_h = ch; at_c = ch;
                    if (kH && rsc && nr > 1) rsk = __ldg(rsc + k0 + div_magic(ch, C, a.mg_c, a.sh_c));
                }
            }
        }
    }
    __syncthreads();
    // copy-out: 8-element vectors; vector v -> (tap, slot8) -> (row_local, channel) -> destination element
    const uint32_t v_per_tap = slots >> 3;                                // slots % 8 == 0 (host)
    const uint32_t n_vec = RS * v_per_tap;
    const uint32_t c8 = C >> 3;                                           // vectors per row inside a tap (nr > 1)
    uint8_t* const out_b = OUT == 2 ? reinterpret_cast<uint8_t*>(a.w_f16) : a.w_codes;
    // (v < 1 024 and divisors < 1 024: floor((v + 0.5) * (1 / d)) in float32 is the exact quotient)
    const float inv_vpt = 1.0f / (float)v_per_tap, inv_c8 = 1.0f / (float)(c8 ? c8 : 1u);
    for (uint32_t v = threadIdx.x; v < n_vec; v += 256) {
        const uint32_t tap = __float2uint_rz(((float)v + 0.5f) * inv_vpt), sv = v - tap * v_per_tap;
        uint32_t rl = 0, cv = sv;
        if (nr > 1) { rl = __float2uint_rz(((float)sv + 0.5f) * inv_c8); cv = sv - rl * c8; }
        const size_t o = (size_t)(k0 + rl) * a.out_pitch + a.out_off + (size_t)tap * Cp + c_lo + cv * 8;
        if (kH) *reinterpret_cast<uint4*>(a.w_f16 + o) = *reinterpret_cast<const uint4*>(s_h + tap * ph + sv * 8);
        if (kB) *reinterpret_cast<uint2*>(out_b + o) = *reinterpret_cast<const uint2*>(s_c + tap * pc + sv * 8);
    }
    // K-padding taps [RS * Cp, pitch) of the rows this CTA starts
    const uint32_t tail0 = RS * Cp;
    if (c_lo == 0 && a.pitch > tail0) {
        const uint32_t tv = (a.pitch - tail0) >> 3;
        for (uint32_t v = threadIdx.x; v < nr * tv; v += 256) {
            const uint32_t rl = v / tv, t8 = v - rl * tv;
            const size_t o = (size_t)(k0 + rl) * a.out_pitch + a.out_off + tail0 + t8 * 8;
            if (kH) *reinterpret_cast<uint4*>(a.w_f16 + o) = make_uint4(0u, 0u, 0u, 0u);
            if (kB) *reinterpret_cast<uint2*>(out_b + o) = make_uint2(0u, 0u);
        }
    }
    __syncthreads();                                                      // the staging images are reused by the next piece
    }
}

}  // namespace slfp

namespace slfp {
__global__ void __launch_bounds__(256) dequantize_any_kernel(const uint8_t* __restrict__ codes, size_t n, int fmt,
                                                             float* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
        out[i] = decode_act_any(codes[i], fmt, c_pow2frac);
}
// post-ReLU code formats, stand-alone (tests / feeding a fused layer by hand): x / K -> encode_relu_fast
__global__ void __launch_bounds__(256) quantize_relu_kernel(const float* __restrict__ x, size_t n, DivK k, int sfp33,
                                                            uint8_t* __restrict__ codes, float* __restrict__ fakeq) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const float q = div_k(x[i], k);
        const uint32_t c = sfp33 ? encode_relu_fast<true>(q) : encode_relu_fast<false>(q);
        if (codes) codes[i] = (uint8_t)c;
        if (fakeq) fakeq[i] = sfp33 ? decode_relu<true>(c, c_pow2frac) : decode_relu<false>(c, c_pow2frac);
    }
}
}  // namespace slfp

using namespace slfp;

extern "C" int slfp_version(void) { return SLFP_B200_VERSION; }
#ifndef SLFP_SOURCE_HASH
#define SLFP_SOURCE_HASH "unknown"
#endif
extern "C" const char* slfp_build_id(void) { return SLFP_SOURCE_HASH; }
extern "C" const char* slfp_last_error(void) { return g_err; }

extern "C" int slfp_quantize_f32(const float* x, size_t n, float k_div, int fmt, unsigned flags, uint8_t* codes,
                                 float* fakeq, void* f16, slfp_stream_t stream) {
    if (n == 0) return 0;
    if (!x || (!codes && !fakeq && !f16)) return set_error(SLFP_ERR_BAD_ARG, "slfp_quantize_f32: null pointer");
    if (fmt == SLFP_FMT_SFP44_OUT && codes)
        return set_error(SLFP_ERR_BAD_ARG, "slfp_quantize_f32: SFP<4,4> layer-out has no 8-bit code");
    if (fmt == SLFP_FMT_SLFP34_RELU || fmt == SLFP_FMT_SFP33_RELU) {
        if (f16) return set_error(SLFP_ERR_UNSUPPORTED, "slfp_quantize_f32: post-ReLU formats have no float16 output");
        const int grid = (int)min((size_t)num_sms() * 8, ceil_div_sz(n, 256));
        quantize_relu_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, make_divk(k_div), fmt == SLFP_FMT_SFP33_RELU, codes, fakeq);
        return check_launch("quantize_relu_kernel");
    }
    QuantArgs a{x, n, make_divk(k_div), codes, fakeq, (__half*)f16, (flags & SLFP_Q_LAYEROUT_ZERO_IS_ZERO) ? 1 : 0, nullptr, 1.0, nullptr};
    cudaStream_t st = (cudaStream_t)stream;
    switch (fmt) {
        case SLFP_FMT_SFP33: return launch_quantize<SLFP_FMT_SFP33>(a, st);
        case SLFP_FMT_SLFP34_ACT: return launch_quantize<SLFP_FMT_SLFP34_ACT>(a, st);
        case SLFP_FMT_SLFP34_WGT: return launch_quantize<SLFP_FMT_SLFP34_WGT>(a, st);
        case SLFP_FMT_SFP44_OUT: return launch_quantize<SLFP_FMT_SFP44_OUT>(a, st);
    }
    return set_error(SLFP_ERR_BAD_ARG, "slfp_quantize_f32: unknown format %d", fmt);
}

extern "C" int slfp_quantize_dyn_f32(const float* x, size_t n, const float* absmax, double divisor, int fmt, unsigned flags,
                                     uint8_t* codes, float* fakeq, void* f16, float* k_out, slfp_stream_t stream) {
    if (n == 0) return 0;
    if (!x || !absmax || (!codes && !fakeq && !f16)) return set_error(SLFP_ERR_BAD_ARG, "slfp_quantize_dyn_f32: null pointer");
    if (!(divisor > 0.0)) return set_error(SLFP_ERR_BAD_ARG, "slfp_quantize_dyn_f32: divisor must be positive");
    if (fmt == SLFP_FMT_SFP44_OUT && codes)
        return set_error(SLFP_ERR_BAD_ARG, "slfp_quantize_dyn_f32: SFP<4,4> layer-out has no 8-bit code");
    // k_div.fast = 1: the kernel re-derives the whole DivK from device memory; this value only selects the launch shape
    QuantArgs a{x, n, make_divk(1.0f), codes, fakeq, (__half*)f16, (flags & SLFP_Q_LAYEROUT_ZERO_IS_ZERO) ? 1 : 0, absmax, divisor, k_out};
    cudaStream_t st = (cudaStream_t)stream;
    switch (fmt) {
        case SLFP_FMT_SFP33: return launch_quantize<SLFP_FMT_SFP33>(a, st);
        case SLFP_FMT_SLFP34_ACT: return launch_quantize<SLFP_FMT_SLFP34_ACT>(a, st);
        case SLFP_FMT_SLFP34_WGT: return launch_quantize<SLFP_FMT_SLFP34_WGT>(a, st);
        case SLFP_FMT_SFP44_OUT: return launch_quantize<SLFP_FMT_SFP44_OUT>(a, st);
    }
    return set_error(SLFP_ERR_BAD_ARG, "slfp_quantize_dyn_f32: format %d (signed code formats only)", fmt);
}

extern "C" int slfp_quantize_nhwc_f32(const float* x, size_t npix, int c, int c_phys, float k_div, int fmt,
                                      uint8_t* codes, slfp_stream_t stream) {
    if (npix == 0) return 0;
    if (!x || !codes || c <= 0 || c_phys < c) return set_error(SLFP_ERR_BAD_ARG, "slfp_quantize_nhwc_f32: bad arguments");
    if (c == c_phys) return slfp_quantize_f32(x, npix * (size_t)c, k_div, fmt, 0, codes, nullptr, nullptr, stream);
    const int vec4 = ((c_phys & 3) == 0 && (((uintptr_t)codes) & 3u) == 0) ? 1 : 0;
    const int grid = (int)min((size_t)num_sms() * 8, ceil_div_sz(npix * (size_t)(vec4 ? c_phys / 4 : c_phys), 256));
    cudaStream_t st = (cudaStream_t)stream;
    const DivK dk = make_divk(k_div);
    switch (fmt) {
        case SLFP_FMT_SFP33: quantize_pad_kernel<SLFP_FMT_SFP33><<<grid, 256, 0, st>>>(x, npix, c, c_phys, dk, codes, vec4); break;
        case SLFP_FMT_SLFP34_ACT: quantize_pad_kernel<SLFP_FMT_SLFP34_ACT><<<grid, 256, 0, st>>>(x, npix, c, c_phys, dk, codes, vec4); break;
        case SLFP_FMT_SLFP34_WGT: quantize_pad_kernel<SLFP_FMT_SLFP34_WGT><<<grid, 256, 0, st>>>(x, npix, c, c_phys, dk, codes, vec4); break;
        default: return set_error(SLFP_ERR_BAD_ARG, "slfp_quantize_nhwc_f32: format %d has no codes", fmt);
    }
    return check_launch("quantize_pad_kernel");
}

extern "C" int slfp_quantize_nchw_f32(const float* x, int n, int c, size_t hw, int c_phys, float k_div, int fmt,
                                      uint8_t* codes, slfp_stream_t stream) {
    if (n <= 0 || hw == 0) return 0;
    if (!x || !codes || c <= 0 || c_phys < c || (c_phys & 3) || ((uintptr_t)codes & 3u))
        return set_error(SLFP_ERR_BAD_ARG, "slfp_quantize_nchw_f32: bad arguments");
    const size_t total = (size_t)n * hw * (c_phys / 4);
    const int grid = (int)min((size_t)num_sms() * 16, ceil_div_sz(total, 256));
    cudaStream_t st = (cudaStream_t)stream;
    const DivK dk = make_divk(k_div);
    const size_t quads = (size_t)n * (hw / 4);
    if (c_phys == 4 && (hw & 3) == 0 && quads < (1ull << 31) && ((((uintptr_t)x) | ((uintptr_t)codes)) & 15u) == 0 &&
        (fmt == SLFP_FMT_SFP33 || fmt == SLFP_FMT_SLFP34_ACT)) {
        const int g4 = (int)min((size_t)num_sms() * 8, ceil_div_sz(quads, 256));
        if (fmt == SLFP_FMT_SFP33)
            quantize_nchw_c4_kernel<SLFP_FMT_SFP33><<<g4, 256, 0, st>>>(x, (uint32_t)quads, c, (uint32_t)(hw / 4), dk, codes);
        else
            quantize_nchw_c4_kernel<SLFP_FMT_SLFP34_ACT><<<g4, 256, 0, st>>>(x, (uint32_t)quads, c, (uint32_t)(hw / 4), dk, codes);
        return check_launch("quantize_nchw_c4_kernel");
    }
    switch (fmt) {
        case SLFP_FMT_SFP33: quantize_nchw_kernel<SLFP_FMT_SFP33><<<grid, 256, 0, st>>>(x, n, c, hw, c_phys, dk, codes); break;
        case SLFP_FMT_SLFP34_ACT: quantize_nchw_kernel<SLFP_FMT_SLFP34_ACT><<<grid, 256, 0, st>>>(x, n, c, hw, c_phys, dk, codes); break;
        default: return set_error(SLFP_ERR_BAD_ARG, "slfp_quantize_nchw_f32: format %d", fmt);
    }
    return check_launch("quantize_nchw_kernel");
}

extern "C" int slfp_quantize_nchw_s2d_f32(const float* x, int n, int c, int h, int w, int c_phys, float k_div, int fmt,
                                          uint8_t* codes, slfp_stream_t stream) {
    if (n <= 0 || h <= 0 || w <= 0) return 0;
    if (!x || !codes || c <= 0 || (h & 1) || (w & 1) || c_phys < 4 * c || (c_phys & 15) || ((uintptr_t)codes & 15u))
        return set_error(SLFP_ERR_BAD_ARG, "slfp_quantize_nchw_s2d_f32: bad arguments (h, w even; c_phys >= 4c, multiple of 16)");
    const size_t total = (size_t)n * (h / 2) * (w / 2);
    const int grid = (int)min((size_t)num_sms() * 16, ceil_div_sz(total, 256));
    cudaStream_t st = (cudaStream_t)stream;
    const DivK dk = make_divk(k_div);
    if (c == 3 && c_phys == 16 && (w & 3) == 0 && (((uintptr_t)x) & 15u) == 0 && total / 2 < (1ull << 31) &&
        (fmt == SLFP_FMT_SFP33 || fmt == SLFP_FMT_SLFP34_ACT) && getenv("SLFP_S2D_GENERIC") == nullptr) {
        const uint32_t pairs = (uint32_t)(total / 2), Wp = (uint32_t)(w / 4), H2 = (uint32_t)(h / 2);
        uint32_t mg_wp, sh_wp, mg_h2, sh_h2;
        magic_u32(Wp, mg_wp, sh_wp);
        magic_u32(H2, mg_h2, sh_h2);
        static int per_sm[2] = {0, 0};
        const int fi = fmt == SLFP_FMT_SFP33 ? 0 : 1;
        if (!per_sm[fi]) {
            int b = 0;
            cudaError_t oe = fi == 0
                ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, quantize_nchw_s2d_c3_kernel<SLFP_FMT_SFP33>, 256, 0)
                : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, quantize_nchw_s2d_c3_kernel<SLFP_FMT_SLFP34_ACT>, 256, 0);
            if (oe != cudaSuccess) { cudaGetLastError(); b = 0; }
            per_sm[fi] = b > 0 ? b : 4;
        }
        const int g3 = (int)min((size_t)num_sms() * per_sm[fi], ceil_div_sz(pairs, 256));      // one resident wave
        if (fi == 0)
            quantize_nchw_s2d_c3_kernel<SLFP_FMT_SFP33><<<g3, 256, 0, st>>>(x, pairs, h, w, Wp, mg_wp, sh_wp, H2, mg_h2, sh_h2, dk, codes);
        else
            quantize_nchw_s2d_c3_kernel<SLFP_FMT_SLFP34_ACT><<<g3, 256, 0, st>>>(x, pairs, h, w, Wp, mg_wp, sh_wp, H2, mg_h2, sh_h2, dk, codes);
        return check_launch("quantize_nchw_s2d_c3_kernel");
    }
    switch (fmt) {
        case SLFP_FMT_SFP33: quantize_nchw_s2d_kernel<SLFP_FMT_SFP33><<<grid, 256, 0, st>>>(x, n, c, h, w, c_phys, dk, codes); break;
        case SLFP_FMT_SLFP34_ACT: quantize_nchw_s2d_kernel<SLFP_FMT_SLFP34_ACT><<<grid, 256, 0, st>>>(x, n, c, h, w, c_phys, dk, codes); break;
        default: return set_error(SLFP_ERR_BAD_ARG, "slfp_quantize_nchw_s2d_f32: format %d", fmt);
    }
    return check_launch("quantize_nchw_s2d_kernel");
}

// Network input of a 3x3 / stride 1 / padding 1 RGB stem as an explicit im2col matrix of float16 images (SLFP_FMT_F16Q):
// row = output pixel, 64 halves = tap (r * 3 + s) * 4 + c for the 9 taps x 3 channels, zeros elsewhere - exactly the KRSC
// weight row of the c_phys = 4 stem (36 of 64 entries).  The stem then runs as a plain 1x1 layer of the no-decode dense kernel
// (ONE K block per tile) instead of the 4-channel gather kernel.  A CTA takes a band of kIm2colRows image rows: it quantizes
// the band + halo ONCE into shared memory ([rows + 2][W + 2] pixels of 4 halves, zero border = the padding) with the exact
// encoder, then writes the im2col rows - eight lanes per pixel, 16 bytes = two taps each, 512 contiguous bytes per warp.
constexpr int kIm2colRows = 8;
template <int FL>
__global__ void __launch_bounds__(256) quantize_nchw_im2col3x3_kernel(const float* __restrict__ x, int N, int H, int W, int bands,
                                                                      DivK k_div, uint4* __restrict__ out) {
    extern __shared__ uint2 s_q[];
    const int Wp = W + 2;
    const size_t HW = (size_t)H * W;
    for (int blk = blockIdx.x; blk < N * bands; blk += gridDim.x) {
        const int n = blk / bands, y0 = (blk - n * bands) * kIm2colRows;
        const int rows = min(kIm2colRows, H - y0);
        for (int i = threadIdx.x; i < (rows + 2) * Wp; i += 256) {
            const int ry = i / Wp, rx = i - ry * Wp;
            const int yy = y0 + ry - 1, xs = rx - 1;
            uint2 v = make_uint2(0u, 0u);
            if (yy >= 0 && yy < H && xs >= 0 && xs < W) {
                const float* src = x + (size_t)n * 3 * HW + (size_t)yy * W + xs;
                unsigned short h3[3];
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    h3[c] = __half_as_ushort(__float2half_rn(decode<FL == SLFP_FMT_SFP33>(encode<FL>(div_k(__ldg(src + (size_t)c * HW), k_div)), c_pow2frac)));
                v = make_uint2((uint32_t)h3[0] | ((uint32_t)h3[1] << 16), (uint32_t)h3[2]);
            }
            s_q[i] = v;
        }
        __syncthreads();
        uint4* dst = out + ((size_t)n * H + y0) * W * 8;
        for (int i = threadIdx.x; i < rows * W * 8; i += 256) {
            const int pix = i >> 3, j = i & 7;                     // chunk j: halves 8j .. 8j+7 = taps 2j, 2j+1
            const int ly = pix / W, lx = pix - ly * W;
            uint2 e0 = make_uint2(0u, 0u), e1 = e0;
            if (2 * j < 9) e0 = s_q[(ly + (2 * j) / 3) * Wp + lx + (2 * j) % 3];
            if (2 * j + 1 < 9) e1 = s_q[(ly + (2 * j + 1) / 3) * Wp + lx + (2 * j + 1) % 3];
            dst[i] = make_uint4(e0.x, e0.y, e1.x, e1.y);
        }
        __syncthreads();
    }
}

extern "C" int slfp_quantize_nchw_im2col3x3_f16q(const float* x, int n, int h, int w, float k_div, int fmt, void* out_f16,
                                                 slfp_stream_t stream) {
    if (n <= 0 || h <= 0 || w <= 0) return 0;
    if (!x || !out_f16 || ((uintptr_t)out_f16 & 15u) || (fmt != SLFP_FMT_SFP33 && fmt != SLFP_FMT_SLFP34_ACT))
        return set_error(SLFP_ERR_BAD_ARG, "slfp_quantize_nchw_im2col3x3_f16q: bad arguments");
    const size_t smem = (size_t)(kIm2colRows + 2) * (w + 2) * sizeof(uint2);
    if (smem > 48 * 1024) return set_error(SLFP_ERR_UNSUPPORTED, "slfp_quantize_nchw_im2col3x3_f16q: image wider than 612 pixels");
    const int bands = (h + kIm2colRows - 1) / kIm2colRows;
    if ((long long)n * bands >= (1ll << 31)) return set_error(SLFP_ERR_UNSUPPORTED, "slfp_quantize_nchw_im2col3x3_f16q: too many row bands");
    const DivK dk = make_divk(k_div);
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = (int)min((long long)num_sms() * 8, (long long)n * bands);
    if (fmt == SLFP_FMT_SFP33)
        quantize_nchw_im2col3x3_kernel<SLFP_FMT_SFP33><<<grid, 256, smem, st>>>(x, n, h, w, bands, dk, (uint4*)out_f16);
    else
        quantize_nchw_im2col3x3_kernel<SLFP_FMT_SLFP34_ACT><<<grid, 256, smem, st>>>(x, n, h, w, bands, dk, (uint4*)out_f16);
    return check_launch("quantize_nchw_im2col3x3_kernel");
}

extern "C" int slfp_quantize_nchw_s2d_f16q(const float* x, int n, int h, int w, float k_div, int fmt, int pad_top, int pad_left,
                                           int hp, int wp, void* out_f16, slfp_stream_t stream) {
    if (n <= 0 || h <= 0 || w <= 0) return 0;
    if (!x || !out_f16 || (h & 1) || (w & 3) || (((uintptr_t)x | (uintptr_t)out_f16) & 15u) || pad_top < 0 || pad_left < 0 ||
        hp < h / 2 + pad_top || wp < w / 2 + pad_left || (fmt != SLFP_FMT_SFP33 && fmt != SLFP_FMT_SLFP34_ACT))
        return set_error(SLFP_ERR_BAD_ARG, "slfp_quantize_nchw_s2d_f16q: RGB input, h even, w %% 4 == 0, 16-byte aligned pointers, padded extents >= pad + h/2, w/2");
    const size_t total = (size_t)n * (h / 2) * (w / 2);
    if (total / 2 >= (1ull << 31)) return set_error(SLFP_ERR_UNSUPPORTED, "slfp_quantize_nchw_s2d_f16q: more than 2^32 pixels");
    const uint32_t pairs = (uint32_t)(total / 2), Wp = (uint32_t)(w / 4), H2 = (uint32_t)(h / 2);
    uint32_t mg_wp, sh_wp, mg_h2, sh_h2;
    magic_u32(Wp, mg_wp, sh_wp);
    magic_u32(H2, mg_h2, sh_h2);
    const DivK dk = make_divk(k_div);
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = (int)min((size_t)num_sms() * 6, ceil_div_sz(pairs, 256));
    if (fmt == SLFP_FMT_SFP33)
        quantize_nchw_s2d_c3_kernel<SLFP_FMT_SFP33, true><<<grid, 256, 0, st>>>(x, pairs, h, w, Wp, mg_wp, sh_wp, H2, mg_h2, sh_h2, dk,
                                                                                 (uint8_t*)out_f16, pad_top, pad_left, hp, wp);
    else
        quantize_nchw_s2d_c3_kernel<SLFP_FMT_SLFP34_ACT, true><<<grid, 256, 0, st>>>(x, pairs, h, w, Wp, mg_wp, sh_wp, H2, mg_h2, sh_h2, dk,
                                                                                      (uint8_t*)out_f16, pad_top, pad_left, hp, wp);
    return check_launch("quantize_nchw_s2d_c3_kernel<f16q>");
}

extern "C" int slfp_gather_quantize_f16(const SlfpGatherChan* table, size_t npix, int c, int c_phys, float k_div, int fmt,
                                        uint8_t* codes, slfp_stream_t stream) {
    if (npix == 0) return 0;
    if (!table || !codes || c <= 0 || c_phys < c || (c_phys & 3) || ((uintptr_t)codes & 3u) || c_phys > 1024)
        return set_error(SLFP_ERR_BAD_ARG, "slfp_gather_quantize_f16: bad arguments (c <= c_phys <= 1024, c_phys %% 4 == 0; codes 4-byte aligned)");
    const size_t per_cta = (size_t)(256 / (c_phys / 4)) * kGatherPix;
    const int grid = (int)min((size_t)num_sms() * 4, ceil_div_sz(npix, per_cta));      // persistent: each CTA builds the value table once
    cudaStream_t st = (cudaStream_t)stream;
    const DivK dk = make_divk(k_div);
    switch (fmt) {
        case SLFP_FMT_SFP33: gather_quantize_kernel<SLFP_FMT_SFP33><<<grid, 256, 0, st>>>(table, npix, c, c_phys, dk, codes); break;
        case SLFP_FMT_SLFP34_ACT: gather_quantize_kernel<SLFP_FMT_SLFP34_ACT><<<grid, 256, 0, st>>>(table, npix, c, c_phys, dk, codes); break;
        case SLFP_FMT_E4M3: gather_quantize_kernel<SLFP_FMT_SFP33, true><<<grid, 256, 0, st>>>(table, npix, c, c_phys, dk, codes); break;
        default: return set_error(SLFP_ERR_BAD_ARG, "slfp_gather_quantize_f16: format %d", fmt);
    }
    return check_launch("gather_quantize_kernel");
}

extern "C" void slfp_magic_u32(unsigned d, unsigned* magic, unsigned* shift) {
    uint32_t mg, sh;
    magic_u32(d, mg, sh);
    if (magic) *magic = mg;
    if (shift) *shift = sh;
}

extern "C" int slfp_gather_quantize_runs_f16(const SlfpGatherRun* runs, int n_runs, size_t npix, int c, int c_phys,
                                             int stage_bytes_per_pixel, float k_div, int fmt, uint8_t* codes, slfp_stream_t stream) {
    if (npix == 0) return 0;
    if (!runs || n_runs <= 0 || !codes || c <= 0 || c_phys < c || (c_phys & 3) || c_phys > kRunMaxC || ((uintptr_t)codes & 3u) ||
        stage_bytes_per_pixel <= 0 || (stage_bytes_per_pixel & 15) || stage_bytes_per_pixel > kRunStageBytes)
        return set_error(SLFP_ERR_BAD_ARG, "slfp_gather_quantize_runs_f16: bad arguments (c <= c_phys <= 1024, c_phys %% 4 == 0, stage bytes per pixel: multiple of 16, <= 32768)");
    if (npix >= (1ull << 31) / 256) return set_error(SLFP_ERR_UNSUPPORTED, "slfp_gather_quantize_runs_f16: too many pixels");
    int tile_pix = (fmt == SLFP_FMT_E4M3 ? kRunStageBytes : kRunStageBytes / 2) / stage_bytes_per_pixel;
    if (tile_pix < 1) return set_error(SLFP_ERR_UNSUPPORTED, "slfp_gather_quantize_runs_f16: %d staging bytes per pixel do not fit", stage_bytes_per_pixel);
    tile_pix = tile_pix > 1024 ? 1024 : tile_pix;
    const int grid = (int)min((size_t)num_sms() * 4, ceil_div_sz(npix, (size_t)tile_pix));
    cudaStream_t st = (cudaStream_t)stream;
    const DivK dk = make_divk(k_div);
    switch (fmt) {
        case SLFP_FMT_SFP33: gather_runs_kernel<SLFP_FMT_SFP33, false><<<grid, 256, 0, st>>>(runs, n_runs, npix, c, c_phys, tile_pix, dk, codes); break;
        case SLFP_FMT_SLFP34_ACT: gather_runs_kernel<SLFP_FMT_SLFP34_ACT, false><<<grid, 256, 0, st>>>(runs, n_runs, npix, c, c_phys, tile_pix, dk, codes); break;
        case SLFP_FMT_E4M3: gather_runs_kernel<SLFP_FMT_SFP33, true><<<grid, 256, 0, st>>>(runs, n_runs, npix, c, c_phys, tile_pix, dk, codes); break;
        default: return set_error(SLFP_ERR_BAD_ARG, "slfp_gather_quantize_runs_f16: format %d", fmt);
    }
    return check_launch("gather_runs_kernel");
}

extern "C" int slfp_dequantize(const uint8_t* codes, size_t n, int fmt, float* out, slfp_stream_t stream) {
    if (n == 0) return 0;
    if (!codes || !out) return set_error(SLFP_ERR_BAD_ARG, "slfp_dequantize: null pointer");
    int grid = (int)min((size_t)num_sms() * 8, ceil_div_sz(n, 1024));
    if (fmt == SLFP_FMT_SLFP34_RELU || fmt == SLFP_FMT_SFP33_RELU || fmt == SLFP_FMT_SFP33_SFAST || fmt == SLFP_FMT_E4M3) {
        dequantize_any_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(codes, n, fmt, out);
        return check_launch("dequantize_any_kernel");
    }
    if (fmt == SLFP_FMT_SFP33)
        dequantize_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(codes, n, out);
    else if (fmt == SLFP_FMT_SLFP34_ACT || fmt == SLFP_FMT_SLFP34_WGT)
        dequantize_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(codes, n, out);
    else
        return set_error(SLFP_ERR_BAD_ARG, "slfp_dequantize: format %d has no codes", fmt);
    return check_launch("dequantize_kernel");
}

extern "C" int slfp_absmax_f32(const float* x, size_t n, float* max_out, int init_zero, slfp_stream_t stream) {
    if (!max_out) return set_error(SLFP_ERR_BAD_ARG, "slfp_absmax_f32: null output");
    cudaStream_t st = (cudaStream_t)stream;
    if (init_zero) {
        cudaError_t e = cudaMemsetAsync(max_out, 0, sizeof(float), st);
        if (e != cudaSuccess) return set_error((int)e, "slfp_absmax_f32: memset: %s", cudaGetErrorString(e));
    }
    if (n == 0) return 0;
    if (!x) return set_error(SLFP_ERR_BAD_ARG, "slfp_absmax_f32: null input");
    int grid = (int)min((size_t)num_sms() * 8, ceil_div_sz(n, 256 * 16));
    absmax_kernel<<<grid, 256, 0, st>>>(x, n, reinterpret_cast<unsigned int*>(max_out));
    return check_launch("absmax_kernel");
}

extern "C" size_t slfp_conv_wpitch(const SlfpConvDesc* d) {
    if (!d) return 0;
    const int cg = d->groups > 1 ? (d->c / d->groups) : d->c_phys;
    size_t k = (size_t)d->r * d->s * cg;
    return d->groups > 1 ? k : (k + 63) / 64 * 64;
}

static int fill_wprep(const SlfpConvDesc* d, const float* w, long long so, long long sc, long long sr, long long ss, float kw,
                      void* w_f16, uint8_t* w_codes, float* w_fakeq, WPrepArgs& a) {
    if (!d || !w) return set_error(SLFP_ERR_BAD_ARG, "slfp_prepare_weights: null pointer");
    a.w = w; a.so = so; a.sc = sc; a.sr = sr; a.ss = ss;
    a.K = d->k; a.R = d->r; a.S = d->s;
    if (d->groups > 1) { a.C = d->c / d->groups; a.Cp = a.C; }
    else { a.C = d->c; a.Cp = d->c_phys; }
    a.pitch = slfp_conv_wpitch(d);
    a.kw = kw; a.dk = make_divk(kw); a.w_f16 = (__half*)w_f16; a.w_codes = w_codes; a.w_fakeq = w_fakeq;
    a.out_pitch = a.pitch; a.out_off = 0; a.row_scale = nullptr; a.lo_off = 0;
    a.e4m3 = (d->flags & SLFP_CONV_E4M3_OPERANDS) ? 1 : 0;
    a.vec8_ok = ((a.Cp & 7) == 0 && (a.pitch & 7) == 0 && (((uintptr_t)a.w_f16) & 15u) == 0 && (((uintptr_t)a.w_codes) & 7u) == 0 &&
                 getenv("SLFP_WPREP_SCALAR") == nullptr) ? 1 : 0;
    if ((size_t)a.K * a.pitch >= (1ull << 31)) return set_error(SLFP_ERR_UNSUPPORTED, "slfp_prepare_weights: tensor with 2^31 or more elements");
    magic_u32((uint32_t)a.pitch, a.mg_pitch, a.sh_pitch);
    magic_u32((uint32_t)a.Cp, a.mg_cp, a.sh_cp);
    magic_u32((uint32_t)a.S, a.mg_s, a.sh_s);
    return 0;
}

// Row-staged fast path (wprep_rows_kernel): fills f and returns the number of CTAs, or 0 when the job needs the gather form.
static unsigned fill_wfast(const WPrepArgs& a, int wfmt, WFastJob& f) {
    if (getenv("SLFP_WPREP_GATHER") != nullptr || wfmt < 0 || a.w_fakeq || a.lo_off || (a.e4m3 && a.w_codes) || (!a.w_f16 && !a.w_codes)) return 0;
    const long long RS = (long long)a.R * a.S, C = a.C;
    if (RS < 1 || RS > kWFastMaxTaps || C < 1 || a.K < 1) return 0;
    const bool contiguous = (RS == 1 || ((a.S == 1 || a.ss == 1) && (a.R == 1 || a.sr == a.S))) && (C == 1 || a.sc == RS) && a.so == C * RS;
    if (!contiguous || (C * RS) % 4 || (a.Cp & 7) || (a.pitch & 7) || (a.out_pitch & 7) || (a.out_off & 7) || a.pitch < (size_t)RS * a.Cp) return 0;
    if ((((uintptr_t)a.w) & 15u) || (((uintptr_t)a.w_f16) & 15u) || (((uintptr_t)a.w_codes) & 7u)) return 0;
    f.w = a.w; f.w_f16 = a.w_f16; f.w_codes = a.w_codes; f.row_scale = a.row_scale;
    f.out_pitch = a.out_pitch; f.out_off = a.out_off;
    f.K = (uint32_t)a.K; f.C = (uint32_t)C; f.Cp = (uint32_t)a.Cp; f.RS = (uint32_t)RS; f.pitch = (uint32_t)a.pitch;
    f.dk = a.dk; f.e4m3 = a.e4m3;
    unsigned blocks;
    if (RS * a.Cp <= kWFastCap) {
        f.n_cchunks = 1; f.cc = f.Cp;
        f.nrows = (a.C == a.Cp) ? (uint32_t)(kWFastCap / (C * RS)) : 1u;
        if (f.nrows < 1) f.nrows = 1;
        blocks = (f.K + f.nrows - 1) / f.nrows;
    } else {
        f.cc = (uint32_t)(kWFastCap / RS) & ~7u;
        if (f.cc < 8) return 0;
        f.n_cchunks = (f.Cp + f.cc - 1) / f.cc;
        f.nrows = 1;
        blocks = f.K * f.n_cchunks;
    }
    magic_u32(f.RS, f.mg_rs, f.sh_rs);
    magic_u32(f.n_cchunks, f.mg_cch, f.sh_cch);
    magic_u32(f.C, f.mg_c, f.sh_c);
    return blocks;
}

template <int FMT>
static void launch_wfast_fmt(const WFastBatch& b, unsigned blocks, int out, cudaStream_t st) {
    const unsigned grid = min(blocks, (unsigned)num_sms() * 6u);          // persistent CTAs (6 resident per SM at 40 registers)
    switch (out) {
        case 0: wprep_rows_kernel<FMT, 0><<<grid, 256, 0, st>>>(b); break;
        case 1: wprep_rows_kernel<FMT, 1><<<grid, 256, 0, st>>>(b); break;
        case 2: wprep_rows_kernel<FMT, 2><<<grid, 256, 0, st>>>(b); break;
        default: wprep_rows_kernel<FMT, 3><<<grid, 256, 0, st>>>(b); break;
    }
}
static int launch_wfast(const WFastBatch& b, unsigned blocks, int wfmt, int out, cudaStream_t st) {
    switch (wfmt) {
        case SLFP_FMT_SFP33: launch_wfast_fmt<SLFP_FMT_SFP33>(b, blocks, out, st); break;
        case SLFP_FMT_SLFP34_WGT: launch_wfast_fmt<SLFP_FMT_SLFP34_WGT>(b, blocks, out, st); break;
        case SLFP_FMT_SLFP34_ACT: launch_wfast_fmt<SLFP_FMT_SLFP34_ACT>(b, blocks, out, st); break;
        default: return set_error(SLFP_ERR_BAD_ARG, "slfp_prepare_weights: bad weight format %d", wfmt);
    }
    return check_launch("wprep_rows_kernel");
}
// output kind of a row-staged job (template argument OUT of wprep_rows_kernel)
static int wfast_out_kind(const WPrepArgs& a) {
    if (a.w_f16 && a.e4m3) return 2;
    if (a.w_f16) return a.w_codes ? 3 : 0;
    return 1;
}

extern "C" int slfp_prepare_weights(const SlfpConvDesc* d, const float* w, long long so, long long sc,
                                    long long sr, long long ss, float kw, int wfmt, void* w_f16,
                                    uint8_t* w_codes, float* w_fakeq, slfp_stream_t stream) {
    WPrepArgs a;
    if (int rc0 = fill_wprep(d, w, so, sc, sr, ss, kw, w_f16, w_codes, w_fakeq, a)) return rc0;
    const size_t total = (size_t)a.K * a.pitch;
    if (total == 0) return 0;
    if (total >= (1ull << 32)) return set_error(SLFP_ERR_UNSUPPORTED, "slfp_prepare_weights: tensor with 2^32 or more elements");
    int grid = (int)min((size_t)num_sms() * 8, ceil_div_sz(total, 256));
    cudaStream_t st = (cudaStream_t)stream;
    {
        static thread_local WFastBatch fb;
        if (unsigned blocks = fill_wfast(a, wfmt, fb.j[0])) {
            fb.n = 1; fb.blk_end[0] = blocks;
            return launch_wfast(fb, blocks, wfmt, wfast_out_kind(a), st);
        }
    }
    switch (wfmt) {
        case SLFP_FMT_SFP33: wprep_kernel<SLFP_FMT_SFP33><<<grid, 256, 0, st>>>(a); break;
        case SLFP_FMT_SLFP34_WGT: wprep_kernel<SLFP_FMT_SLFP34_WGT><<<grid, 256, 0, st>>>(a); break;
        case SLFP_FMT_SLFP34_ACT: wprep_kernel<SLFP_FMT_SLFP34_ACT><<<grid, 256, 0, st>>>(a); break;
        case -1: wprep_kernel<-1><<<grid, 256, 0, st>>>(a); break;
        default: return set_error(SLFP_ERR_BAD_ARG, "slfp_prepare_weights: bad weight format %d", wfmt);
    }
    return check_launch("wprep_kernel");
}

extern "C" int slfp_prepare_weights_jobs(int n, const SlfpWeightJob* host_jobs, int wfmt, slfp_stream_t stream) {
    if (n <= 0) return 0;
    if (!host_jobs) return set_error(SLFP_ERR_BAD_ARG, "slfp_prepare_weights_jobs: null table");
    cudaStream_t st = (cudaStream_t)stream;
    for (int t0 = 0; t0 < n; t0 += kWBatchMax) {
        static thread_local WPrepBatch b;
        static thread_local WFastBatch fb[4];                                 // one batch per output kind
        b.n = 0;
        unsigned blocks = 0, fblocks[4] = {0u, 0u, 0u, 0u};
        for (int o = 0; o < 4; ++o) fb[o].n = 0;
        for (int t = t0; t < n && t < t0 + kWBatchMax; ++t) {
            const SlfpWeightJob& jb = host_jobs[t];
            WPrepArgs& a = b.a[b.n];
            int rc = fill_wprep(jb.desc, jb.w, jb.w_stride[0], jb.w_stride[1], jb.w_stride[2], jb.w_stride[3], jb.kw, jb.w_f16,
                                jb.w_codes, nullptr, a);
            if (rc) return rc;
            if (jb.out_pitch) {
                if (jb.out_pitch < jb.out_offset + a.pitch)
                    return set_error(SLFP_ERR_BAD_ARG, "slfp_prepare_weights_jobs: out_pitch %zu < out_offset %zu + row %zu", jb.out_pitch,
                                     jb.out_offset, a.pitch);
                a.out_pitch = jb.out_pitch; a.out_off = jb.out_offset;
                if (jb.lo_offset && jb.out_pitch < jb.out_offset + jb.lo_offset + a.pitch)
                    return set_error(SLFP_ERR_BAD_ARG, "slfp_prepare_weights_jobs: out_pitch %zu too small for the lo part", jb.out_pitch);
            }
            a.row_scale = jb.row_scale;
            a.lo_off = jb.out_pitch ? jb.lo_offset : 0;
            if (a.e4m3 && (wfmt != SLFP_FMT_SFP33 || a.row_scale || a.lo_off))
                return set_error(SLFP_ERR_BAD_ARG, "slfp_prepare_weights_jobs: e4m3 operands are SFP<3,3> weights without row scale / lo part");
            a.vec8_ok = ((a.Cp & 7) == 0 && (a.pitch & 7) == 0 && (a.out_pitch & 7) == 0 && (a.out_off & 7) == 0 && (a.lo_off & 7) == 0 &&
                         (((uintptr_t)a.w_f16) & 15u) == 0 && (((uintptr_t)a.w_codes) & 7u) == 0 &&
                         getenv("SLFP_WPREP_SCALAR") == nullptr) ? 1 : 0;
            const size_t total = (size_t)a.K * a.pitch;
            if (total == 0) continue;
            if (total >= (1ull << 32)) return set_error(SLFP_ERR_UNSUPPORTED, "slfp_prepare_weights_jobs: tensor with 2^32 or more elements");
            WFastBatch& f = fb[wfast_out_kind(a)];
            if (unsigned fbk = fill_wfast(a, wfmt, f.j[f.n])) {                // row-staged form; the slot in b is reused
                unsigned& fbl = fblocks[wfast_out_kind(a)];
                fbl += fbk;
                f.blk_end[f.n++] = fbl;
                continue;
            }
            blocks += (unsigned)ceil_div_sz(total, kWBatchChunk);
            b.blk_end[b.n++] = blocks;
        }
        for (int o = 0; o < 4; ++o) {
            if (fb[o].n == 0) continue;
            int rc = launch_wfast(fb[o], fblocks[o], wfmt, o, st);
            if (rc) return rc;
        }
        if (b.n == 0) continue;
        switch (wfmt) {
            case SLFP_FMT_SFP33: wprep_batch_kernel<SLFP_FMT_SFP33><<<blocks, 256, 0, st>>>(b); break;
            case SLFP_FMT_SLFP34_WGT: wprep_batch_kernel<SLFP_FMT_SLFP34_WGT><<<blocks, 256, 0, st>>>(b); break;
            case SLFP_FMT_SLFP34_ACT: wprep_batch_kernel<SLFP_FMT_SLFP34_ACT><<<blocks, 256, 0, st>>>(b); break;
            case -1: wprep_batch_kernel<-1><<<blocks, 256, 0, st>>>(b); break;
            default: return set_error(SLFP_ERR_BAD_ARG, "slfp_prepare_weights_jobs: bad weight format %d", wfmt);
        }
        int rc = check_launch("wprep_batch_kernel");
        if (rc) return rc;
    }
    return 0;
}

extern "C" int slfp_quantize_host_f32(const float* host_x, size_t n, float k_div, int fmt, uint8_t* host_codes,
                                      float* host_fakeq) {
    if (n == 0) return 0;
    if (!host_x || (!host_codes && !host_fakeq)) return set_error(SLFP_ERR_BAD_ARG, "slfp_quantize_host_f32: null");
    float* dx = nullptr; uint8_t* dc = nullptr; float* dq = nullptr;
    cudaStream_t st = nullptr;
    cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    int rc = 0;
    if (e == cudaSuccess) e = cudaMalloc(&dx, n * 4);
    if (e == cudaSuccess && host_codes) e = cudaMalloc(&dc, n);
    if (e == cudaSuccess && host_fakeq) e = cudaMalloc(&dq, n * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dx, host_x, n * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) rc = slfp_quantize_f32(dx, n, k_div, fmt, 0, dc, dq, nullptr, st);
    if (e == cudaSuccess && rc == 0 && host_codes) e = cudaMemcpyAsync(host_codes, dc, n, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && rc == 0 && host_fakeq) e = cudaMemcpyAsync(host_fakeq, dq, n * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(dx); cudaFree(dc); cudaFree(dq);
    if (st) cudaStreamDestroy(st);
    if (e != cudaSuccess) return set_error((int)e, "slfp_quantize_host_f32: %s", cudaGetErrorString(e));
    return rc;
}
