// conv_direct.cu -- CUDA-core kernels around the tensor-core path:
//   * depthwise / grouped convolution forward on 8-bit codes (MobileNet, ShuffleNet): an HBM-bound
//     stencil, 4 channels per thread, fp32 accumulate (Conv2d_Q.forward with groups > 1,
//     utils/conv2d_func.py:20-25; callers nets_imgnet/mobilenetv1.py:27, nets_cifar/shufflenet_v2.py:68)
//   * convolution backward with the identity straight-through estimator (utils/sfp_quant.py:50-53):
//       dx = dgrad(gy*Ka*Kw, w_q)/Ka   dw = wgrad(gy*Ka*Kw, x_q)/Kw   db = sum(gy)
//   * max-pool on codes and global average pool (glue of the fused eval pipeline).
#include "slfp_common.cuh"

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
    if (e.relu) t = fmaxf(t, 0.0f);
    if (e.y_f32) e.y_f32[off] = t;
    if (e.y_f16) reinterpret_cast<__half*>(e.y_f16)[off] = __float2half_rn(t);
    if (e.y_codes) {
        const float q = div_rn(t, e.next_k_div);
        e.y_codes[pix * e.k_phys_out + k] =
            (uint8_t)((e.next_fmt == SLFP_FMT_SFP33) ? encode<SLFP_FMT_SFP33>(q) : encode<SLFP_FMT_SLFP34_ACT>(q));
    }
    if (e.y_codes2) {
        const float q = div_rn(t, e.next_k_div2);
        e.y_codes2[pix * e.k_phys_out + k] =
            (uint8_t)((e.next_fmt == SLFP_FMT_SFP33) ? encode<SLFP_FMT_SFP33>(q) : encode<SLFP_FMT_SLFP34_ACT>(q));
    }
}

// Depthwise (C == K == groups): thread = (pixel, 4 consecutive channels).  Loads are one 32-bit word
// of codes per tap; a warp covers 128 contiguous channels (or several pixels when C < 128).
template <bool SFP33>
__global__ void __launch_bounds__(256) dwconv_fwd_kernel(DirectParams p) {
    __shared__ uint32_t s_tab[16];
    if (threadIdx.x < 16) s_tab[threadIdx.x] = c_pow2frac[threadIdx.x];
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
                    if (c0 + j < p.C) {
                        const float xv = decode<SFP33>((wd >> (8 * j)) & 0xffu, s_tab);
                        acc[j] = fmaf(xv, decode<SFP33>(__ldg(p.w + ((size_t)(c0 + j) * p.R + r) * p.S + s), s_tab), acc[j]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (c0 + j < p.K) epilogue_store(p.epi, acc[j], pix, c0 + j, p.K, s_tab);
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
    if ((d->fmt != SLFP_FMT_SFP33 && d->fmt != SLFP_FMT_SLFP34_ACT) ||
        (epi->y_codes && epi->next_fmt != SLFP_FMT_SFP33 && epi->next_fmt != SLFP_FMT_SLFP34_ACT))
        return set_error(SLFP_ERR_UNSUPPORTED, "conv2d_fwd(grouped): the stencil kernels read and write the signed code formats only");
    const bool sfp = d->fmt == SLFP_FMT_SFP33;
    const bool depthwise = d->groups == d->c && d->k == d->c && (d->c_phys % 4) == 0 && (((uintptr_t)x_codes) & 3u) == 0;
    const size_t total = depthwise ? (size_t)d->n * p.Ho * p.Wo * (d->c_phys / 4) : (size_t)d->n * p.Ho * p.Wo * d->k;
    const int grid = (int)min((size_t)num_sms() * 16, ceil_div_sz(total, 256));
    if (depthwise) {
        if (sfp) dwconv_fwd_kernel<true><<<grid, 256, 0, st>>>(p); else dwconv_fwd_kernel<false><<<grid, 256, 0, st>>>(p);
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
        uint4 best = make_uint4(0u, 0u, 0u, 0u);
        for (int r = 0; r < kh; ++r) {
            const int hi = ho * stride - pad + r;
            if (hi < 0 || hi >= H) continue;
            for (int s = 0; s < kw; ++s) {
                const int wi = wo * stride - pad + s;
                if (wi < 0 || wi >= W) continue;
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)n * H + hi) * W + wi) * Cp + c0));
                best.x = __vmaxu4(best.x, v.x); best.y = __vmaxu4(best.y, v.y);
                best.z = __vmaxu4(best.z, v.z); best.w = __vmaxu4(best.w, v.w);
            }
        }
        *reinterpret_cast<uint4*>(y + pix * Cp + c0) = best;
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
