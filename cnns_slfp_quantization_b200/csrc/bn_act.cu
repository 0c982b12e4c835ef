// Training-mode BatchNorm2d + residual add + ReLU of the QAT step (BASELINE config 4), fused, on float32 NHWC
// tensors (the layout Conv2d_Q returns).  Replaces the caller net's stock sequence
//     out = relu(bn(conv(x)))            nets_imgnet/resnet50.py:71-77
//     out = bn3(conv3(out)); out += identity; out = relu(out)      :79-88
// which PyTorch runs as BatchNorm (2 reads + 1 write), add (2 reads + 1 write) and ReLU (1 read + 1 write) forward and
// again backward - half of the 34.6 ms QAT step (profiles/r01_final.md section 2).  Here:
//   forward   bn_reduce_kernel<0>  1 read   per-channel sum / sum of squares; the last CTA turns the sums into
//                                        mean, 1/sqrt(var + eps), the folded scale / shift and the running statistics
//             bn_apply_kernel   1-2 reads + 1 write   y = relu(x * scale + shift + residual)
//   backward  bn_reduce_kernel<1>  3 reads   g = gy * (y > 0); sum g, sum g (x - mean); last CTA -> dgamma, dbeta and
//                                             the three per-channel coefficients of dx
//             bn_bwd_apply_kernel   3 reads + 1-2 writes   dx = a (g - b - (x - mean) c);  d_residual = g
// Reductions: float32 per thread and CTA, then one double-precision atomicAdd per CTA and channel (600 float32 partials in
// double: the float32 results do not depend on the arrival order except for ties at the 1e-13 level); the last CTA to
// arrive turns the sums into the per-channel coefficients and clears the accumulators for the next launch.
// HBM-bound; algorithmic bytes per element: forward 12 (16 with a residual), backward 28 (32).
#include "slfp_common.cuh"

namespace slfp {

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

constexpr int kBnThreads = 256;
constexpr int kBnMaxBlocks = 592;                       // 148 SMs x 4

// Thread geometry: a row of C floats is covered by CT = min(C / 4, 256) threads with one float4 each (grid.y walks
// over further 1 024-channel chunks), RT = 256 / CT rows are in flight per CTA iteration.
struct BnGeom {
    int C, C4, CT, RT;
    size_t M;
};

// workspace layout: [0] arrival counter (uint, 16 bytes reserved), double acc[2][C] - everything zero between launches
// whatever C the previous launch had.  The per-channel coefficients (forward: scale, shift | backward: a, b, c) go
// to a separate caller-provided scratch vector `coef` [4][C] that needs no initialisation.
__host__ __device__ __forceinline__ double* ws_acc(float* ws) { return reinterpret_cast<double*>(ws + 4); }

template <int MODE>   // 0: forward statistics of x; 1: backward sums of g and g (x - mean)
__global__ void __launch_bounds__(kBnThreads) bn_reduce_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                               const float* __restrict__ y, BnGeom g, int relu,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               float eps, float momentum, float* __restrict__ running_mean,
                                                               float* __restrict__ running_var, float* __restrict__ save_mean,
                                                               float* __restrict__ save_invstd, float* __restrict__ dgamma,
                                                               float* __restrict__ dbeta, float* __restrict__ ws,
                                                               float* __restrict__ coef, float* __restrict__ dx_absmax) {
    __shared__ float4 s_a[kBnThreads], s_b[kBnThreads];
    __shared__ bool s_last;
    const int tx = threadIdx.x % g.CT, ty = threadIdx.x / g.CT;
    const int c4 = blockIdx.y * 256 + tx;                                  // float4 column
    const bool active = ty < g.RT && c4 < g.C4;
    float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sb = sa;
    float4 mean4 = sa, sc4 = sa, sh4 = sa;
    const bool remask = MODE == 1 && relu && y == nullptr;                // ReLU mask recomputed from x (no residual): y is not read
    if (MODE == 1 && active) {
        mean4 = *reinterpret_cast<const float4*>(save_mean + 4 * c4);
        if (remask) {                                                     // the forward's own scale / shift, bit for bit
            const float4 gm = *reinterpret_cast<const float4*>(gamma + 4 * c4), bt = *reinterpret_cast<const float4*>(beta + 4 * c4);
            const float4 is = *reinterpret_cast<const float4*>(save_invstd + 4 * c4);
            sc4 = make_float4(gm.x * is.x, gm.y * is.y, gm.z * is.z, gm.w * is.w);
            sh4 = make_float4(bt.x - mean4.x * sc4.x, bt.y - mean4.y * sc4.y, bt.z - mean4.z * sc4.z, bt.w - mean4.w * sc4.w);
        }
    }
    if (active) {
        const size_t stride = (size_t)gridDim.x * g.RT;
        // U rows per round, every load of a round issued before the arithmetic (a plain unrolled loop keeps its exit test
        // between the loads: ONE 16-byte load in flight per thread, 3.6 TB/s on the statistics pass - ncu, profiles/r04_final.md);
        // rows past the end read as zeros, which add nothing to either sum
        constexpr int U = MODE == 0 ? 8 : 4;
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (size_t r0 = (size_t)blockIdx.x * g.RT + ty; r0 < g.M; r0 += stride * U) {
            float4 v[U], gg[U], yy[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const size_t r = r0 + (size_t)u * stride;
                const bool in = r < g.M;
                const size_t o = (in ? r : r0) * g.C4 + c4;
                v[u] = in ? __ldg(reinterpret_cast<const float4*>(x) + o) : zero4;
                if (MODE == 1) {
                    gg[u] = in ? __ldg(reinterpret_cast<const float4*>(gy) + o) : zero4;
                    if (relu && !remask) yy[u] = in ? __ldg(reinterpret_cast<const float4*>(y) + o) : zero4;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (MODE == 0) {
                    sa.x += v[u].x; sa.y += v[u].y; sa.z += v[u].z; sa.w += v[u].w;
                    sb.x += v[u].x * v[u].x; sb.y += v[u].y * v[u].y; sb.z += v[u].z * v[u].z; sb.w += v[u].w * v[u].w;
                } else {
                    float4 q = gg[u];
                    if (remask) {
                        q.x = v[u].x * sc4.x + sh4.x > 0.f ? q.x : 0.f; q.y = v[u].y * sc4.y + sh4.y > 0.f ? q.y : 0.f;
                        q.z = v[u].z * sc4.z + sh4.z > 0.f ? q.z : 0.f; q.w = v[u].w * sc4.w + sh4.w > 0.f ? q.w : 0.f;
                    } else if (relu) {
                        q.x = yy[u].x > 0.f ? q.x : 0.f; q.y = yy[u].y > 0.f ? q.y : 0.f;
                        q.z = yy[u].z > 0.f ? q.z : 0.f; q.w = yy[u].w > 0.f ? q.w : 0.f;
                    }
                    sa.x += q.x; sa.y += q.y; sa.z += q.z; sa.w += q.w;
                    sb.x += q.x * (v[u].x - mean4.x); sb.y += q.y * (v[u].y - mean4.y);
                    sb.z += q.z * (v[u].z - mean4.z); sb.w += q.w * (v[u].w - mean4.w);
                }
            }
        }
    }
    s_a[threadIdx.x] = sa; s_b[threadIdx.x] = sb;
    __syncthreads();
    double* acc = ws_acc(ws);
    if (ty == 0 && c4 < g.C4) {                                           // fixed order over the rows of the CTA
        for (int j = 1; j < g.RT; ++j) {
            const float4 a = s_a[j * g.CT + tx], b = s_b[j * g.CT + tx];
            sa.x += a.x; sa.y += a.y; sa.z += a.z; sa.w += a.w;
            sb.x += b.x; sb.y += b.y; sb.z += b.z; sb.w += b.w;
        }
        double* pa = acc + 4 * c4;
        double* pb = acc + g.C + 4 * c4;
        atomicAdd(pa + 0, (double)sa.x); atomicAdd(pa + 1, (double)sa.y); atomicAdd(pa + 2, (double)sa.z); atomicAdd(pa + 3, (double)sa.w);
        atomicAdd(pb + 0, (double)sb.x); atomicAdd(pb + 1, (double)sb.y); atomicAdd(pb + 2, (double)sb.z); atomicAdd(pb + 3, (double)sb.w);
    }
    // last CTA to arrive (over the whole grid) finishes the reduction
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned total = gridDim.x * gridDim.y;
        const unsigned prev = atomicAdd(reinterpret_cast<unsigned*>(ws), 1u);
        s_last = prev + 1 == total;
        if (s_last) *reinterpret_cast<unsigned*>(ws) = 0u;                // self-cleaning for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const double inv_m = 1.0 / (double)g.M;
    if (MODE == 1 && dx_absmax && threadIdx.x == 0) *dx_absmax = 0.f;     // bn_bwd_apply_kernel accumulates max |dx| into it
    // four channels per thread and round, every load of a round issued before the arithmetic (the loop used to expose
    // three dependent L2 latencies per channel: 24 us for 2 048 channels, all other SMs idle)
    for (int c0 = threadIdx.x; c0 < g.C; c0 += 4 * kBnThreads) {
        double sa[4], sb[4];
        float gm[4], bt[4], rm[4], rv[4], is[4], sm[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + j * kBnThreads;
            if (c < g.C) {
                sa[j] = __ldcg(acc + c); sb[j] = __ldcg(acc + g.C + c);
                gm[j] = __ldg(gamma + c);
                bt[j] = (MODE == 0 || remask) ? __ldg(beta + c) : 0.f;
                if (MODE == 0 && running_mean) { rm[j] = running_mean[c]; rv[j] = running_var[c]; }
                if (MODE == 1) { is[j] = save_invstd[c]; sm[j] = save_mean[c]; }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + j * kBnThreads;
            if (c >= g.C) continue;
            acc[c] = 0.0; acc[g.C + c] = 0.0;                             // self-cleaning
            if (MODE == 0) {
                const double mean = sa[j] * inv_m;
                double var = sb[j] * inv_m - mean * mean;
                var = var < 0.0 ? 0.0 : var;
                const float invstd = 1.0f / sqrtf((float)(var + (double)eps));
                save_mean[c] = (float)mean;
                save_invstd[c] = invstd;
                const float sc = gm[j] * invstd;
                coef[c] = sc;
                coef[g.C + c] = bt[j] - (float)mean * sc;
                if (running_mean) {                                       // nn.BatchNorm2d: unbiased variance into the running estimate
                    const double unb = g.M > 1 ? var * ((double)g.M / (double)(g.M - 1)) : var;
                    running_mean[c] = (1.f - momentum) * rm[j] + momentum * (float)mean;
                    running_var[c] = (1.f - momentum) * rv[j] + momentum * (float)unb;
                }
            } else {
                const float invstd = is[j];
                dbeta[c] = (float)sa[j];
                dgamma[c] = (float)(sb[j] * (double)invstd);
                coef[c] = gm[j] * invstd;                                                       // a
                coef[g.C + c] = (float)(sa[j] * inv_m);                                         // b = mean(g)
                coef[2 * g.C + c] = (float)(sb[j] * inv_m * (double)invstd * (double)invstd);   // c = mean(g (x - mean)) / var
                if (remask) coef[3 * g.C + c] = bt[j] - sm[j] * (gm[j] * invstd);               // the forward's shift (its scale is a)
            }
        }
    }
}

// up to two consumers of y with their own activation scales (a block output feeds the next block's conv1 and its
// downsample conv): each gets the 8-bit codes of the reference quantizer of y / K (utils/conv2d_func.py:21) - what
// slfp_quantize_nhwc_f32 would produce from y in a pass of its own.  Post-ReLU values only (y >= 0): encode_relu<>.
struct BnQuantOut {
    uint8_t* codes[2];
    DivK dk[2];
    int n;
};

template <int FMT>    // activation code format of the quantized outputs, or -1: none
__global__ void __launch_bounds__(kBnThreads) bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ res, size_t n4,
                                                              int C4, int relu, const float* __restrict__ coef, int C,
                                                              float* __restrict__ y, BnQuantOut q) {
    for (size_t i = (size_t)blockIdx.x * kBnThreads + threadIdx.x; i < n4; i += (size_t)gridDim.x * kBnThreads) {
        const int c4 = (C4 & (C4 - 1)) == 0 ? (int)(i & (size_t)(C4 - 1)) : (int)(i % (size_t)C4);
        const float4 v = ldg_stream(reinterpret_cast<const float4*>(x) + i);
        const float4 sc = __ldg(reinterpret_cast<const float4*>(coef) + c4);
        const float4 sh = __ldg(reinterpret_cast<const float4*>(coef + C) + c4);
        float4 o = make_float4(v.x * sc.x + sh.x, v.y * sc.y + sh.y, v.z * sc.z + sh.z, v.w * sc.w + sh.w);
        if (res) {
            const float4 r = ldg_stream(reinterpret_cast<const float4*>(res) + i);
            o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        reinterpret_cast<float4*>(y)[i] = o;
        if (FMT >= 0) {
            constexpr int F = FMT < 0 ? 0 : FMT;
            const float ov[4] = {o.x, o.y, o.z, o.w};
            const bool has_nan = !(o.x == o.x && o.y == o.y && o.z == o.z && o.w == o.w);     // fmaxf(NaN, 0) = 0, but a NaN residual ...
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (j < q.n) {
                    uint32_t w = 0u;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float qv = div_k_fused(ov[e], q.dk[j]);
                        w |= (has_nan ? encode<F>(div_k(ov[e], q.dk[j])) : encode_relu<F>(qv)) << (8 * e);
                    }
                    reinterpret_cast<uint32_t*>(q.codes[j])[i] = w;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kBnThreads) bn_bwd_apply_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                                                                  const float* __restrict__ y, size_t n4, int C4, int relu,
                                                                  const float* __restrict__ coef, int C,
                                                                  const float* __restrict__ save_mean, float* __restrict__ dx,
                                                                  float* __restrict__ dres, float* __restrict__ dx_absmax) {
    float amax = 0.f;
    for (size_t i = (size_t)blockIdx.x * kBnThreads + threadIdx.x; i < n4; i += (size_t)gridDim.x * kBnThreads) {
        const int c4 = (C4 & (C4 - 1)) == 0 ? (int)(i & (size_t)(C4 - 1)) : (int)(i % (size_t)C4);
        float4 g = ldg_stream(reinterpret_cast<const float4*>(gy) + i);
        const float4 v = ldg_stream(reinterpret_cast<const float4*>(x) + i);
        const float4 a = __ldg(reinterpret_cast<const float4*>(coef) + c4);
        if (relu && y == nullptr) {
            const float4 sh = __ldg(reinterpret_cast<const float4*>(coef + 3 * C) + c4);
            g.x = v.x * a.x + sh.x > 0.f ? g.x : 0.f; g.y = v.y * a.y + sh.y > 0.f ? g.y : 0.f;
            g.z = v.z * a.z + sh.z > 0.f ? g.z : 0.f; g.w = v.w * a.w + sh.w > 0.f ? g.w : 0.f;
        } else if (relu) {
            const float4 yy = ldg_stream(reinterpret_cast<const float4*>(y) + i);
            g.x = yy.x > 0.f ? g.x : 0.f; g.y = yy.y > 0.f ? g.y : 0.f;
            g.z = yy.z > 0.f ? g.z : 0.f; g.w = yy.w > 0.f ? g.w : 0.f;
        }
        const float4 b = __ldg(reinterpret_cast<const float4*>(coef + C) + c4);
        const float4 c = __ldg(reinterpret_cast<const float4*>(coef + 2 * C) + c4);
        const float4 m = __ldg(reinterpret_cast<const float4*>(save_mean) + c4);
        float4 o;
        o.x = a.x * ((g.x - b.x) - (v.x - m.x) * c.x);
        o.y = a.y * ((g.y - b.y) - (v.y - m.y) * c.y);
        o.z = a.z * ((g.z - b.z) - (v.z - m.z) * c.z);
        o.w = a.w * ((g.w - b.w) - (v.w - m.w) * c.w);
        reinterpret_cast<float4*>(dx)[i] = o;
        if (dres) reinterpret_cast<float4*>(dres)[i] = g;
        // max |dx| on the bit patterns (non-negative floats order like unsigned integers; NaN / Inf order above everything,
        // like slfp_absmax_f32): the convolution that receives dx as its gy skips its own abs-max pass
        const uint32_t m01 = max(__float_as_uint(o.x) & 0x7fffffffu, __float_as_uint(o.y) & 0x7fffffffu);
        const uint32_t m23 = max(__float_as_uint(o.z) & 0x7fffffffu, __float_as_uint(o.w) & 0x7fffffffu);
        amax = __uint_as_float(max(__float_as_uint(amax), max(m01, m23)));
    }
    if (dx_absmax) {
        uint32_t m = __float_as_uint(amax);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((threadIdx.x & 31) == 0 && m != 0u) atomicMax(reinterpret_cast<unsigned int*>(dx_absmax), m);
    }
}

// max_blocks: one resident wave of the kernel that runs on the grid (the statistics pass holds 2 CTAs per SM at 89 registers)
static bool bn_geom(size_t m, int c, BnGeom& g, dim3& grid, int max_blocks = kBnMaxBlocks) {
    if (c <= 0 || (c & 3) || m == 0) return false;
    g.C = c; g.C4 = c / 4; g.M = m;
    g.CT = g.C4 < 256 ? g.C4 : 256;
    g.RT = kBnThreads / g.CT;
    const size_t want = (m + (size_t)g.RT * 4 - 1) / ((size_t)g.RT * 4);  // >= 4 rows per thread before another CTA pays off
    const int gy_ = (g.C4 + 255) / 256;
    int gx = max_blocks / gy_;
    if ((size_t)gx > want) gx = (int)want;
    if (gx < 1) gx = 1;
    grid = dim3((unsigned)gx, (unsigned)gy_);
    return true;
}

}  // namespace slfp

using namespace slfp;

extern "C" size_t slfp_bn_act_workspace_floats(int c) {
    return 4 + (size_t)4 * (c > 0 ? c : 0);
}

static int bn_act_fwd_impl(const float* x, size_t m, int c, const float* gamma, const float* beta, const float* residual,
                           int relu, float eps, float momentum, float* running_mean, float* running_var, float* y,
                           float* save_mean, float* save_invstd, float* workspace, float* coef, int fmt, const BnQuantOut& q,
                           slfp_stream_t stream) {
    if (m == 0) return 0;
    if (!x || !gamma || !beta || !y || !save_mean || !save_invstd || !workspace || !coef)
        return set_error(SLFP_ERR_BAD_ARG, "slfp_bn_act_fwd_train: null pointer");
    BnGeom g; dim3 grid;
    if (!bn_geom(m, c, g, grid, kBnMaxBlocks / 2) || ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)residual | (uintptr_t)workspace | (uintptr_t)save_mean | (uintptr_t)coef) & 15u)))
        return set_error(SLFP_ERR_BAD_ARG, "slfp_bn_act_fwd_train: needs c %% 4 == 0 and 16-byte aligned tensors");
    if ((running_mean == nullptr) != (running_var == nullptr)) return set_error(SLFP_ERR_BAD_ARG, "slfp_bn_act_fwd_train: running statistics come in pairs");
    cudaStream_t st = (cudaStream_t)stream;
    bn_reduce_kernel<0><<<grid, kBnThreads, 0, st>>>(x, nullptr, nullptr, g, 0, gamma, beta, eps, momentum, running_mean, running_var,
                                                     save_mean, save_invstd, nullptr, nullptr, workspace, coef, nullptr);
    if (int rc = check_launch("bn_reduce_kernel<0>")) return rc;
    const size_t n4 = m * (size_t)g.C4;
    const int ag = (int)min((size_t)num_sms() * 8, (n4 + kBnThreads - 1) / kBnThreads);
    if (q.n == 0) bn_apply_kernel<-1><<<ag, kBnThreads, 0, st>>>(x, residual, n4, g.C4, relu, coef, c, y, q);
    else if (fmt == SLFP_FMT_SFP33) bn_apply_kernel<SLFP_FMT_SFP33><<<ag, kBnThreads, 0, st>>>(x, residual, n4, g.C4, relu, coef, c, y, q);
    else bn_apply_kernel<SLFP_FMT_SLFP34_ACT><<<ag, kBnThreads, 0, st>>>(x, residual, n4, g.C4, relu, coef, c, y, q);
    return check_launch("bn_apply_kernel");
}

extern "C" int slfp_bn_act_fwd_train(const float* x, size_t m, int c, const float* gamma, const float* beta, const float* residual,
                                     int relu, float eps, float momentum, float* running_mean, float* running_var, float* y,
                                     float* save_mean, float* save_invstd, float* workspace, float* coef, slfp_stream_t stream) {
    BnQuantOut q;
    q.n = 0; q.codes[0] = q.codes[1] = nullptr; q.dk[0] = q.dk[1] = make_divk(1.0f);
    return bn_act_fwd_impl(x, m, c, gamma, beta, residual, relu, eps, momentum, running_mean, running_var, y, save_mean, save_invstd,
                           workspace, coef, -1, q, stream);
}

extern "C" int slfp_bn_act_fwd_train_quant(const float* x, size_t m, int c, const float* gamma, const float* beta, const float* residual,
                                           float eps, float momentum, float* running_mean, float* running_var, float* y,
                                           float* save_mean, float* save_invstd, float* workspace, float* coef, int fmt, int n_codes,
                                           const float* k_div, uint8_t* const* codes, slfp_stream_t stream) {
    if (n_codes < 1 || n_codes > 2 || !k_div || !codes || (fmt != SLFP_FMT_SFP33 && fmt != SLFP_FMT_SLFP34_ACT))
        return set_error(SLFP_ERR_BAD_ARG, "slfp_bn_act_fwd_train_quant: 1 or 2 code outputs in SFP<3,3> / SLFP<3,4> activation format");
    BnQuantOut q;
    q.n = n_codes; q.codes[1] = nullptr; q.dk[1] = make_divk(1.0f);
    for (int j = 0; j < n_codes; ++j) {
        if (!codes[j] || (((uintptr_t)codes[j]) & 3u) || !(k_div[j] > 0.f))
            return set_error(SLFP_ERR_BAD_ARG, "slfp_bn_act_fwd_train_quant: code tensors are 4-byte aligned, scales positive");
        q.codes[j] = codes[j]; q.dk[j] = make_divk(k_div[j]);
    }
    return bn_act_fwd_impl(x, m, c, gamma, beta, residual, /*relu=*/1, eps, momentum, running_mean, running_var, y, save_mean, save_invstd,
                           workspace, coef, fmt, q, stream);
}

extern "C" int slfp_bn_act_bwd(const float* gy, const float* x, const float* y, size_t m, int c, const float* gamma, const float* beta,
                               const float* save_mean, const float* save_invstd, int relu, float* dx, float* d_residual,
                               float* dgamma, float* dbeta, float* workspace, float* coef, float* dx_absmax, slfp_stream_t stream) {
    if (m == 0) return 0;
    if (!gy || !x || (relu && !y && !beta) || !gamma || !save_mean || !save_invstd || !dx || !dgamma || !dbeta || !workspace || !coef)
        return set_error(SLFP_ERR_BAD_ARG, "slfp_bn_act_bwd: null pointer");
    BnGeom g; dim3 grid;
    if (!bn_geom(m, c, g, grid) ||
        ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)gy | (uintptr_t)dx | (uintptr_t)d_residual | (uintptr_t)workspace | (uintptr_t)save_mean | (uintptr_t)coef) & 15u)))
        return set_error(SLFP_ERR_BAD_ARG, "slfp_bn_act_bwd: needs c %% 4 == 0 and 16-byte aligned tensors");
    cudaStream_t st = (cudaStream_t)stream;
    bn_reduce_kernel<1><<<grid, kBnThreads, 0, st>>>(x, gy, y, g, relu, gamma, beta, 0.f, 0.f, nullptr, nullptr,
                                                     const_cast<float*>(save_mean), const_cast<float*>(save_invstd), dgamma, dbeta, workspace, coef, dx_absmax);
    if (int rc = check_launch("bn_reduce_kernel<1>")) return rc;
    const size_t n4 = m * (size_t)g.C4;
    const int ag = (int)min((size_t)num_sms() * 8, (n4 + kBnThreads - 1) / kBnThreads);
    bn_bwd_apply_kernel<<<ag, kBnThreads, 0, st>>>(gy, x, y, n4, g.C4, relu, coef, c, save_mean, dx, d_residual, dx_absmax);
    return check_launch("bn_bwd_apply_kernel");
}

// ---- 3x3 / stride 2 / padding 1 max-pool of the QAT step (the ResNet stem pool, nets_imgnet/resnet50.py:237) on float32 NHWC
// The stock NHWC kernels write 8-byte indices and their backward ran at 0.8 TB/s (0.44 + 0.88 ms of the batch-128 step).
// Forward: one thread = 4 channels of one output pixel, nine predicated 16-byte loads, the window position of the maximum
// (PyTorch's rule: scan rows then columns, strict '>', NaN wins; ties keep the first) as ONE byte per element.  Backward:
// a gather - an input pixel lies in at most 2 x 2 windows; it takes gy of those whose recorded position is its own -
// deterministic, no atomics, no zero-fill.
namespace slfp {

__global__ void __launch_bounds__(256) maxpool3x3s2_fwd_f32_kernel(const float* __restrict__ x, int N, int H, int W, int C4, int Ho, int Wo,
                                                                   float* __restrict__ y, uint32_t* __restrict__ idx) {
    const size_t total = (size_t)N * Ho * Wo * C4;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const int c4 = (int)(i % (size_t)C4);
        size_t t = i / (size_t)C4;
        const int wo = (int)(t % (size_t)Wo); t /= (size_t)Wo;
        const int ho = (int)(t % (size_t)Ho);
        const int n = (int)(t / (size_t)Ho);
        float4 v[9];
        bool ok[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int h = 2 * ho - 1 + r, w = 2 * wo - 1 + s;
                ok[r * 3 + s] = h >= 0 && h < H && w >= 0 && w < W;
                v[r * 3 + s] = ok[r * 3 + s] ? __ldg(reinterpret_cast<const float4*>(x) + (((size_t)n * H + h) * W + w) * C4 + c4)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        uint32_t k[4] = {9u, 9u, 9u, 9u};
#pragma unroll
        for (int p = 0; p < 9; ++p) {
            if (!ok[p]) continue;
            const float e[4] = {v[p].x, v[p].y, v[p].z, v[p].w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (k[j] == 9u || e[j] > m[j] || e[j] != e[j]) { m[j] = e[j]; k[j] = (uint32_t)p; }
        }
        reinterpret_cast<float4*>(y)[i] = make_float4(m[0], m[1], m[2], m[3]);
        idx[i] = k[0] | (k[1] << 8) | (k[2] << 16) | (k[3] << 24);
    }
}

__global__ void __launch_bounds__(256) maxpool3x3s2_bwd_f32_kernel(const float* __restrict__ gy, const uint32_t* __restrict__ idx, int N,
                                                                   int H, int W, int C4, int Ho, int Wo, float* __restrict__ dx) {
    const size_t total = (size_t)N * H * W * C4;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const int c4 = (int)(i % (size_t)C4);
        size_t t = i / (size_t)C4;
        const int w = (int)(t % (size_t)W); t /= (size_t)W;
        const int h = (int)(t % (size_t)H);
        const int n = (int)(t / (size_t)H);
        // windows whose rows 2 ho - 1 .. 2 ho + 1 contain h: even h -> ho = h / 2 (r = 1); odd h -> (h - 1) / 2 (r = 2), (h + 1) / 2 (r = 0)
        int hos[2], rs[2], nh = 0, wos[2], ss[2], nw = 0;
        if (h & 1) { hos[nh] = (h - 1) >> 1; rs[nh++] = 2; if (((h + 1) >> 1) < Ho) { hos[nh] = (h + 1) >> 1; rs[nh++] = 0; } }
        else if ((h >> 1) < Ho) { hos[nh] = h >> 1; rs[nh++] = 1; }
        if (w & 1) { wos[nw] = (w - 1) >> 1; ss[nw++] = 2; if (((w + 1) >> 1) < Wo) { wos[nw] = (w + 1) >> 1; ss[nw++] = 0; } }
        else if ((w >> 1) < Wo) { wos[nw] = w >> 1; ss[nw++] = 1; }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int a = 0; a < nh; ++a)
            for (int b = 0; b < nw; ++b) {
                if (hos[a] >= Ho || wos[b] >= Wo) continue;
                const size_t o = (((size_t)n * Ho + hos[a]) * Wo + wos[b]) * C4 + c4;
                const uint32_t kk = __ldg(idx + o);
                const float4 g = __ldg(reinterpret_cast<const float4*>(gy) + o);
                const uint32_t me = (uint32_t)(rs[a] * 3 + ss[b]);
                if ((kk & 0xffu) == me) acc.x += g.x;
                if (((kk >> 8) & 0xffu) == me) acc.y += g.y;
                if (((kk >> 16) & 0xffu) == me) acc.z += g.z;
                if ((kk >> 24) == me) acc.w += g.w;
            }
        reinterpret_cast<float4*>(dx)[i] = acc;
    }
}

}  // namespace slfp

extern "C" int slfp_maxpool3x3s2_fwd_f32(const float* x, int n, int h, int w, int c, float* y, uint8_t* idx, slfp_stream_t stream) {
    if (!x || !y || !idx || n <= 0 || h <= 0 || w <= 0 || c <= 0 || (c & 3) || ((((uintptr_t)x | (uintptr_t)y) & 15u)) || (((uintptr_t)idx) & 3u))
        return set_error(SLFP_ERR_BAD_ARG, "slfp_maxpool3x3s2_fwd_f32: needs c %% 4 == 0 and aligned tensors");
    const int Ho = (h + 2 - 3) / 2 + 1, Wo = (w + 2 - 3) / 2 + 1;
    const size_t total = (size_t)n * Ho * Wo * (c / 4);
    const int grid = (int)min((size_t)num_sms() * 8, (total + 255) / 256);
    maxpool3x3s2_fwd_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, h, w, c / 4, Ho, Wo, y, reinterpret_cast<uint32_t*>(idx));
    return check_launch("maxpool3x3s2_fwd_f32_kernel");
}

extern "C" int slfp_maxpool3x3s2_bwd_f32(const float* gy, const uint8_t* idx, int n, int h, int w, int c, float* dx, slfp_stream_t stream) {
    if (!gy || !dx || !idx || n <= 0 || h <= 0 || w <= 0 || c <= 0 || (c & 3) || ((((uintptr_t)gy | (uintptr_t)dx) & 15u)) || (((uintptr_t)idx) & 3u))
        return set_error(SLFP_ERR_BAD_ARG, "slfp_maxpool3x3s2_bwd_f32: needs c %% 4 == 0 and aligned tensors");
    const int Ho = (h + 2 - 3) / 2 + 1, Wo = (w + 2 - 3) / 2 + 1;
    const size_t total = (size_t)n * h * w * (c / 4);
    const int grid = (int)min((size_t)num_sms() * 8, (total + 255) / 256);
    maxpool3x3s2_bwd_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gy, reinterpret_cast<const uint32_t*>(idx), n, h, w, c / 4, Ho, Wo, dx);
    return check_launch("maxpool3x3s2_bwd_f32_kernel");
}
