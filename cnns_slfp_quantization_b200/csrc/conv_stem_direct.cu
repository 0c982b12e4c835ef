// conv_stem_direct.cu -- 3x3 network stems on RGB input (c_phys = 4 codes per pixel) as a CUDA-core direct convolution.
//
// Replaces Conv2d_Q.forward (utils/conv2d_func.py:20-25) for the first layer of MobileNetV1 (3 -> 32, stride 2 / 1;
// nets_imgnet/mobilenetv1.py:20-27, nets_cifar/mobilenetv1.py), ShuffleNetV2 (3 -> 24, stride 1;
// nets_cifar/shufflenet_v2.py:153) and VGG-16 (3 -> 64; nets_cifar/vgg16.py:31).  K = 27 taps is far too thin for
// the tensor-core path: through the 4-channel implicit-GEMM kernel these layers ran at 7-14 TFLOP/s and were the
// largest single launch of their nets (profiles/r01_mbv1_depthwise.md, profiles/r02_config5.md).  Here a thread owns
// ONE output pixel and all K <= 64 output channels in registers: the 27 inputs are decoded once (256-entry float32
// table in shared memory), the float32 filter sits in shared memory as [tap][k] and is read with warp-uniform
// (broadcast) 16-byte loads, 27 x K FFMAs per pixel - FFMA-issue-bound, float32 operands (exact, unlike the float16
// tensor-core operands).  Epilogue: folded per-channel affine (+ ReLU) and quantize-on-store for one or two consumers
// in the fused pipeline's code formats (post-ReLU codes, the signed fast SFP<3,3> codes, or the exact signed encoder).
#include <stdlib.h>

#include "slfp_common.cuh"
#include "sm100_ptx.cuh"

namespace slfp {

struct StemParams {
    const uint8_t* x;          // NHWC codes, 4 per pixel
    const __half* w;           // KRSC float16 image of weight_q, row pitch `wpitch` halves, element (r*3+s)*4 + c
    int wpitch;
    int N, H, W, Ho, Wo, K, stride, pad;
    int act_fmt;
    const float* ch_mul;
    const float* ch_add;
    int relu;
    int out_mode;              // 0: post-ReLU fast codes, 1: signed fast SFP<3,3>, 2: exact signed encoder, 3: e4m3 bytes
    int out_sfp33;
    uint8_t* y1;
    uint8_t* y2;
    int k_phys_out;
    float sc1, sc2;            // 1 / (16 Ka_next)
    DivK kd1, kd2;
};

template <int KP>
__global__ void __launch_bounds__(256) stem3x3_direct_kernel(const StemParams p) {
    __shared__ float s_dec[256];
    __shared__ __align__(16) float s_w[27 * KP];
    __shared__ __align__(16) float s_mul[KP];
    __shared__ __align__(16) float s_add[KP];
    s_dec[threadIdx.x] = decode_act_any((uint32_t)threadIdx.x, p.act_fmt, c_pow2frac);
    for (int i = threadIdx.x; i < 27 * KP; i += 256) {
        const int j = i / KP, k = i - j * KP;                  // j = tap * 3 + c
        const int t = j / 3, c = j - t * 3;
        s_w[i] = k < p.K ? __half2float(p.w[(size_t)k * p.wpitch + t * 4 + c]) : 0.0f;
    }
    for (int k = threadIdx.x; k < KP; k += 256) {
        s_mul[k] = k < p.K ? __ldg(p.ch_mul + k) : 0.0f;
        s_add[k] = k < p.K ? __ldg(p.ch_add + k) : 0.0f;
    }
    __syncthreads();
    const uint32_t w_base = ptx::smem_u32(s_w);
    const uint32_t total = (uint32_t)p.N * p.Ho * p.Wo;
    const int enc_shift = p.out_sfp33 ? 19 : 18, enc_bias = p.out_sfp33 ? 0x76F : 0xEDF;      // encode_relu_fast_raw16<>
    for (uint32_t pix = blockIdx.x * 256u + threadIdx.x; pix < total; pix += gridDim.x * 256u) {
        const int wo = (int)(pix % (uint32_t)p.Wo);
        const uint32_t rest = pix / (uint32_t)p.Wo;
        const int ho = (int)(rest % (uint32_t)p.Ho), n = (int)(rest / (uint32_t)p.Ho);
        uint32_t cw[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int hi = ho * p.stride - p.pad + t / 3, wi = wo * p.stride - p.pad + t % 3;
            cw[t] = 0u;                                        // zero padding: code 0 = 0.0
            if (hi >= 0 && hi < p.H && wi >= 0 && wi < p.W)
                cw[t] = __ldg(reinterpret_cast<const uint32_t*>(p.x) + ((size_t)n * p.H + hi) * p.W + wi);
        }
        float acc[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) acc[k] = 0.0f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float xv = s_dec[(cw[t] >> (8 * c)) & 0xffu];
#pragma unroll
                for (int k4 = 0; k4 < KP / 4; ++k4) {
                    const float4 w4 = ptx::lds128_f4(w_base + (uint32_t)(((t * 3 + c) * KP + 4 * k4) * 4));
                    acc[4 * k4 + 0] = fmaf(xv, w4.x, acc[4 * k4 + 0]);
                    acc[4 * k4 + 1] = fmaf(xv, w4.y, acc[4 * k4 + 1]);
                    acc[4 * k4 + 2] = fmaf(xv, w4.z, acc[4 * k4 + 2]);
                    acc[4 * k4 + 3] = fmaf(xv, w4.w, acc[4 * k4 + 3]);
                }
            }
        }
        // ---- epilogue: affine (+ ReLU) -> codes for one or two consumers, 16 bytes at a time ----------------------
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            uint8_t* y = pass ? p.y2 : p.y1;
            if (y == nullptr) continue;
            const float sc = pass ? p.sc2 : p.sc1;
            const DivK kd = pass ? p.kd2 : p.kd1;
            uint8_t* dst = y + (size_t)pix * p.k_phys_out;
#pragma unroll
            for (int ch = 0; ch < (KP + 15) / 16; ++ch) {
                uint32_t pk[4] = {0u, 0u, 0u, 0u};
                if (p.out_mode == 3) {
                    // e4m3 bytes: two values per cvt, 1 / Ka_next folded into the affine (KP % 8 == 0: whole pairs)
                    const float rk = kd.rk;
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const int k = ch * 16 + i;
                        if (k >= KP) continue;
                        const float v0 = fmaf(acc[k], s_mul[k] * rk, s_add[k] * rk), v1 = fmaf(acc[k + 1], s_mul[k + 1] * rk, s_add[k + 1] * rk);
                        uint32_t two = p.relu ? encode_e4m3x2_relu(v0, v1) : encode_e4m3x2(v0, v1);
                        if (k >= p.K) two = 0u; else if (k + 1 >= p.K) two &= 0xffu;
                        pk[i >> 2] |= two << (8 * (i & 3));
                    }
                    if (ch * 16 < p.k_phys_out) *reinterpret_cast<uint4*>(dst + ch * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    continue;
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int k = ch * 16 + i;
                    if (k >= KP) continue;
                    float v = fmaf(acc[k], s_mul[k], s_add[k]);
                    if (p.relu) v = fmaxf(v, 0.0f);
                    uint32_t code;
                    if (p.out_mode == 0) {
                        const int32_t m = ((int32_t)__float_as_uint(__saturatef(v * sc)) >> enc_shift) - enc_bias;
                        code = (uint32_t)(m < 0 ? 0 : (m > 255 ? 255 : m));
                    } else if (p.out_mode == 3) {
                        code = encode_e4m3(div_k_fused(v, kd));
                    } else if (p.out_mode == 1) {
                        const int32_t m = ((int32_t)__float_as_uint(__saturatef(fabsf(v) * sc)) >> 19) - 0x76F;
                        code = (uint32_t)(m < 0 ? 0 : (m > 127 ? 127 : m)) | ((__float_as_uint(v) >> 24) & 0x80u);
                    } else {
                        code = p.out_sfp33 ? encode_q<SLFP_FMT_SFP33>(div_k_fused(v, kd), v)
                                           : encode_q<SLFP_FMT_SLFP34_ACT>(div_k_fused(v, kd), v);
                    }
                    code = k < p.K ? code : 0u;
                    pk[i >> 2] |= code << (8 * (i & 3));
                }
                if (ch * 16 < p.k_phys_out) *reinterpret_cast<uint4*>(dst + ch * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
    }
}

bool conv2d_fwd_stem_direct_supported(const SlfpConvDesc* d, const SlfpEpilogue* e) {
    static const bool off = getenv("SLFP_NO_STEM_DIRECT") != nullptr;
    if (off || d->groups != 1 || d->c_phys != 4 || d->c > 3 || d->r != 3 || d->s != 3 || d->dil_h != 1 || d->dil_w != 1) return false;
    if (d->stride_h != d->stride_w || d->pad_h != d->pad_w || d->pad_h_extra || d->pad_w_extra) return false;
    if (d->k != 24 && d->k != 32 && !(d->k == 64 && getenv("SLFP_STEM_DIRECT_64"))) return false;   // K = 64 (VGG-16): the 4-channel tcgen05 kernel is faster (measured)
    if (!e->ch_mul || !e->ch_add || e->residual || e->y_f32 || e->y_f16 || !e->y_codes || e->layerout) return false;
    if (e->k_phys_out % 16 != 0 || e->k_phys_out < d->k || e->k_phys_out > 64) return false;
    const int f = e->next_fmt;
    if (f == SLFP_FMT_SLFP34_RELU || f == SLFP_FMT_SFP33_RELU) return e->relu != 0 && e->next_k_div > 0.f;
    if (f == SLFP_FMT_SFP33_SFAST || f == SLFP_FMT_E4M3) return e->next_k_div > 0.f;
    return f == SLFP_FMT_SFP33 || f == SLFP_FMT_SLFP34_ACT;
}

int conv2d_fwd_stem_direct(const SlfpConvDesc* d, const uint8_t* x_codes, const void* w_f16, const SlfpEpilogue* e, cudaStream_t st) {
    StemParams p;
    p.x = x_codes; p.w = reinterpret_cast<const __half*>(w_f16); p.wpitch = (int)slfp_conv_wpitch(d);
    p.N = d->n; p.H = d->h; p.W = d->w; p.K = d->k; p.stride = d->stride_h; p.pad = d->pad_h;
    p.Ho = (d->h + 2 * d->pad_h - 3) / d->stride_h + 1;
    p.Wo = (d->w + 2 * d->pad_w - 3) / d->stride_w + 1;
    if (p.Ho <= 0 || p.Wo <= 0 || d->n <= 0) return set_error(SLFP_ERR_BAD_ARG, "conv2d_fwd(stem): empty output");
    if ((unsigned long long)d->n * p.Ho * p.Wo >= (1ull << 31)) return set_error(SLFP_ERR_UNSUPPORTED, "conv2d_fwd(stem): more than 2^31 output pixels");
    if ((((uintptr_t)x_codes) & 3u) || (((uintptr_t)e->y_codes | (uintptr_t)e->y_codes2) & 15u))
        return set_error(SLFP_ERR_BAD_ARG, "conv2d_fwd(stem): misaligned tensors");
    p.act_fmt = d->fmt;
    p.ch_mul = e->ch_mul; p.ch_add = e->ch_add; p.relu = e->relu;
    const int f = e->next_fmt;
    p.out_mode = (f == SLFP_FMT_SLFP34_RELU || f == SLFP_FMT_SFP33_RELU) ? 0 : (f == SLFP_FMT_SFP33_SFAST ? 1 : (f == SLFP_FMT_E4M3 ? 3 : 2));
    p.out_sfp33 = (f == SLFP_FMT_SFP33_RELU || f == SLFP_FMT_SFP33_SFAST || f == SLFP_FMT_SFP33) ? 1 : 0;
    p.y1 = e->y_codes; p.y2 = e->y_codes2; p.k_phys_out = e->k_phys_out;
    p.sc1 = (float)(1.0 / (16.0 * (double)e->next_k_div));
    p.sc2 = (float)(1.0 / (16.0 * (double)(e->y_codes2 ? e->next_k_div2 : 1.0f)));
    p.kd1 = make_divk(e->next_k_div);
    p.kd2 = make_divk(e->y_codes2 ? e->next_k_div2 : 1.0f);
    const size_t total = (size_t)d->n * p.Ho * p.Wo;
    const int grid = (int)min((size_t)num_sms() * 8, ceil_div_sz(total, 256));
    switch (d->k) {
        case 24: stem3x3_direct_kernel<24><<<grid, 256, 0, st>>>(p); break;
        case 32: stem3x3_direct_kernel<32><<<grid, 256, 0, st>>>(p); break;
        default: stem3x3_direct_kernel<64><<<grid, 256, 0, st>>>(p); break;
    }
    return check_launch("stem3x3_direct_kernel");
}

}  // namespace slfp
