// elementwise.cu -- activations (utils/activation_func.py:6-36) and the fused multi-tensor
// revised-SGD step (utils/optimizer.py:30-73 DSGD, :98-132 SSGD, :154-190 NormalSGD).
#include <vector>

#include "slfp_common.cuh"

namespace slfp {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int KIND>
__global__ void __launch_bounds__(256) act_fwd_kernel(const float* __restrict__ x, size_t n, float* __restrict__ y) {
    const bool vec = ((((uintptr_t)x) | ((uintptr_t)y)) & 15u) == 0;
    const size_t n4 = vec ? n / 4 : 0;
    auto f = [](float v) -> float {
        if (KIND == SLFP_ACT_STL) {
            // where(|x| <= 1, x, sign(x) * (ln|x| + 1))            activation_func.py:10
            const float a = fabsf(v);
            const float sg = (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f);
            return (a <= 1.f) ? v : sg * (logf(a) + 1.f);
        } else if (KIND == SLFP_ACT_SWISH) {
            return v * sigmoidf_(v);                                 // :30-32
        } else {
            return sigmoidf_(v);                                     // :34-36
        }
    };
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
        float4 v = reinterpret_cast<const float4*>(x)[i];
        v.x = f(v.x); v.y = f(v.y); v.z = f(v.z); v.w = f(v.w);
        reinterpret_cast<float4*>(y)[i] = v;
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) y[i] = f(x[i]);
}

template <int KIND>
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy, size_t n,
                                                      float* __restrict__ gx) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const float g = gy[i];
        float r;
        if (KIND == SLFP_ACT_STL) {
            // where(|g| <= 1, 1, 1/|g|) * g : clips the gradient by its own magnitude      :16
            const float a = fabsf(g);
            r = ((a <= 1.f) ? 1.f : 1.f / a) * g;
        } else {
            const float v = x[i], sg = sigmoidf_(v);
            r = (KIND == SLFP_ACT_SWISH) ? g * (sg * (1.f + v * (1.f - sg))) : g * sg * (1.f - sg);
        }
        gx[i] = r;
    }
}

// ---- multi-tensor SGD ----------------------------------------------------------------------------------
struct SgdTensor { float* p; float* g; float* buf; size_t n; size_t first_chunk; };
constexpr int kSgdChunk = 256 * 8;    // elements per CTA-chunk

struct SgdArgs {
    const SgdTensor* tensors;  // device array
    int n_tensors;
    size_t total_chunks;
    int mode, qfmt;
    float neg_lr, momentum, one_minus_damp, wd;
    int nesterov, first_step;
};

__device__ __forceinline__ float sgd_quant(float w, int qfmt, const uint32_t* tab) {
    if (qfmt == SLFP_FMT_SFP33) return decode<true>(encode<SLFP_FMT_SFP33>(w), tab);
    if (qfmt == SLFP_FMT_SLFP34_WGT) return decode<false>(encode<SLFP_FMT_SLFP34_WGT>(w), tab);
    return w;
}

__global__ void __launch_bounds__(256) sgd_step_kernel(SgdArgs a) {
    __shared__ uint32_t s_tab[16];
    if (threadIdx.x < 16) s_tab[threadIdx.x] = c_pow2frac[threadIdx.x];
    __syncthreads();
    for (size_t chunk = blockIdx.x; chunk < a.total_chunks; chunk += gridDim.x) {
        // binary search: last tensor whose first_chunk <= chunk
        int lo = 0, hi = a.n_tensors - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (a.tensors[mid].first_chunk <= chunk) lo = mid; else hi = mid - 1;
        }
        const SgdTensor t = a.tensors[lo];
        const size_t base = (chunk - t.first_chunk) * kSgdChunk;
#pragma unroll
        for (int j = 0; j < kSgdChunk / 256; ++j) {
            const size_t i = base + (size_t)j * 256 + threadIdx.x;
            if (i >= t.n) break;
            float g = t.g[i];
            const float w0 = t.p[i];
            if (a.wd != 0.f) { g = fmaf(a.wd, w0, g); t.g[i] = g; }            // d_p.add_(wd, p)   optimizer.py:46
            float d = g;
            if (a.momentum != 0.f) {
                float b;
                if (a.first_step) b = g;                                   // buf = clone(d_p)   :50
                else { b = t.buf[i] * a.momentum; b = fmaf(a.one_minus_damp, g, b); }   // :53
                t.buf[i] = b;
                d = a.nesterov ? fmaf(a.momentum, b, g) : b;                 // :54-57
            }
            const float step = a.neg_lr * d;
            float w1 = w0 + step;                                          // :59
            if (a.mode != SLFP_SGD_NORMAL) {
                float scale;
                if (a.mode == SLFP_SGD_DSGD) {
                    const float q0 = sgd_quant(w0, a.qfmt, s_tab);         // quantizes the RAW weight :58,:60
                    const float q1 = sgd_quant(w1, a.qfmt, s_tab);
                    scale = (fabsf(q0 - q1) < 0.0001f) ? 2.f : 0.f;        // :61-63
                } else {
                    scale = fabsf(w1) + 1.f;                               // SSGD :130
                }
                w1 = w1 + step * scale;                                    // :64 / :131
            }
            t.p[i] = w1;
        }
    }
}

}  // namespace slfp

using namespace slfp;

extern "C" int slfp_act_fwd(const float* x, size_t n, int kind, float* y, slfp_stream_t stream) {
    if (n == 0) return 0;
    if (!x || !y) return set_error(SLFP_ERR_BAD_ARG, "slfp_act_fwd: null pointer");
    const int grid = (int)min((size_t)num_sms() * 8, ceil_div_sz(n, 1024));
    cudaStream_t st = (cudaStream_t)stream;
    switch (kind) {
        case SLFP_ACT_STL: act_fwd_kernel<SLFP_ACT_STL><<<grid, 256, 0, st>>>(x, n, y); break;
        case SLFP_ACT_SWISH: act_fwd_kernel<SLFP_ACT_SWISH><<<grid, 256, 0, st>>>(x, n, y); break;
        case SLFP_ACT_SIGMOID: act_fwd_kernel<SLFP_ACT_SIGMOID><<<grid, 256, 0, st>>>(x, n, y); break;
        default: return set_error(SLFP_ERR_BAD_ARG, "slfp_act_fwd: kind %d", kind);
    }
    return check_launch("act_fwd_kernel");
}

extern "C" int slfp_act_bwd(const float* x, const float* gy, size_t n, int kind, float* gx, slfp_stream_t stream) {
    if (n == 0) return 0;
    if (!gy || !gx || (kind != SLFP_ACT_STL && !x)) return set_error(SLFP_ERR_BAD_ARG, "slfp_act_bwd: null pointer");
    const int grid = (int)min((size_t)num_sms() * 8, ceil_div_sz(n, 256));
    cudaStream_t st = (cudaStream_t)stream;
    switch (kind) {
        case SLFP_ACT_STL: act_bwd_kernel<SLFP_ACT_STL><<<grid, 256, 0, st>>>(x, gy, n, gx); break;
        case SLFP_ACT_SWISH: act_bwd_kernel<SLFP_ACT_SWISH><<<grid, 256, 0, st>>>(x, gy, n, gx); break;
        case SLFP_ACT_SIGMOID: act_bwd_kernel<SLFP_ACT_SIGMOID><<<grid, 256, 0, st>>>(x, gy, n, gx); break;
        default: return set_error(SLFP_ERR_BAD_ARG, "slfp_act_bwd: kind %d", kind);
    }
    return check_launch("act_bwd_kernel");
}

// The tensor table lives in a small device buffer owned by the library (grown on demand, reused);
// it is filled with a stream-ordered copy from a pinned staging buffer.
extern "C" int slfp_sgd_step(int n_tensors, float* const* host_params, float* const* host_grads, float* const* host_bufs,
                             const size_t* host_sizes, int mode, int q_fmt, double lr, double momentum, double dampening,
                             double weight_decay, int nesterov, int first_step, slfp_stream_t stream) {
    if (n_tensors <= 0) return 0;
    if (!host_params || !host_grads || !host_sizes || (momentum != 0.0 && !host_bufs))
        return set_error(SLFP_ERR_BAD_ARG, "slfp_sgd_step: null table");
    static thread_local SgdTensor* d_tab = nullptr;
    static thread_local SgdTensor* h_tab = nullptr;
    static thread_local int cap = 0;
    static thread_local cudaEvent_t ev = nullptr;
    static thread_local int tab_dev = -1;                  // the device the table (and the event) belong to
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
    if (dev != tab_dev && d_tab) {                         // the caller switched devices: the table must live on the current one
        if (ev) { cudaEventSynchronize(ev); cudaEventDestroy(ev); ev = nullptr; }
        cudaFree(d_tab); cudaFreeHost(h_tab);
        d_tab = nullptr; h_tab = nullptr; cap = 0;
    }
    tab_dev = dev;
    if (n_tensors > cap) {
        if (d_tab) cudaFree(d_tab);
        if (h_tab) cudaFreeHost(h_tab);
        cap = n_tensors * 2;
        if ((e = cudaMalloc(&d_tab, sizeof(SgdTensor) * cap)) != cudaSuccess ||
            (e = cudaMallocHost(&h_tab, sizeof(SgdTensor) * cap)) != cudaSuccess) {
            cap = 0; d_tab = nullptr; h_tab = nullptr;
            return set_error((int)e, "slfp_sgd_step: table allocation: %s", cudaGetErrorString(e));
        }
        if (!ev) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    } else if (ev) {
        cudaEventSynchronize(ev);      // previous copy out of the pinned table has completed
    }
    size_t chunks = 0;
    for (int i = 0; i < n_tensors; ++i) {
        h_tab[i] = SgdTensor{host_params[i], host_grads[i], host_bufs ? host_bufs[i] : nullptr, host_sizes[i], chunks};
        chunks += (host_sizes[i] + kSgdChunk - 1) / kSgdChunk;
    }
    if (chunks == 0) return 0;
    if ((e = cudaMemcpyAsync(d_tab, h_tab, sizeof(SgdTensor) * n_tensors, cudaMemcpyHostToDevice, st)) != cudaSuccess)
        return set_error((int)e, "slfp_sgd_step: table copy: %s", cudaGetErrorString(e));
    if (ev) cudaEventRecord(ev, st);
    SgdArgs a{d_tab, n_tensors, chunks, mode, q_fmt, (float)(-lr), (float)momentum, (float)(1.0 - dampening),
              (float)weight_decay, nesterov, first_step};
    const int grid = (int)min((size_t)num_sms() * 8, chunks);
    sgd_step_kernel<<<grid, 256, 0, st>>>(a);
    return check_launch("sgd_step_kernel");
}
