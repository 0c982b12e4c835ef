/*
 * slfp_b200.h -- C ABI of libslfp_b200.so: the SLFP/SFP quantized-convolution hot path of
 * happyxtt/CNNs_SLFP_quantization, as hand-written sm_100a (B200) CUDA kernels.
 *
 * This is the drop-in boundary.  The reference has no native layer: its hot
 * path is Python (utils/sfp_quant.py, utils/conv2d_func.py, utils/activation_func.py,
 * utils/optimizer.py) on top of ATen/cuDNN.  Each entry point below names the reference
 * interface (file:line under the reference root) whose arithmetic it replaces.  The Python
 * mirror of the reference's callables (cnns_slfp_quantization_b200/utils/...) binds these
 * symbols with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*.
 *   - every call is asynchronous on `stream` (a cudaStream_t), never allocates device memory,
 *     never synchronises (the *_host helpers in the "host buffer" section excepted).
 *   - return value: 0 = ok, otherwise a cudaError_t value or one of the SLFP_ERR_* codes;
 *     slfp_last_error() returns a thread-local message for the last failing call.
 *   - activations are NHWC ("channels last"), weights are KRSC (out-ch, filter row, filter col,
 *     in-ch).  8-bit storage codes are defined in the table below; they are ours (the reference
 *     only ever materialises fake-quant float32) and decode bit-exactly to the reference's values.
 *
 * 8-bit codes:  bit 7 = sign, low 7 bits u
 *     SLFP<3,4> (q_bit 8): u = E*16 + M, E = 1..7, M = 0..15   ->  2^(E-4 + M/16)
 *     SFP<3,3>  (q_bit 7): u = E*8  + m, E = 1..7, m = 0..7    ->  (1 + m/8) * 2^(E-4)
 *     escapes: u = 0 -> 0.0 ; u = 1 -> +-1e-10 ; u = 2 -> +-15.32165 (SLFP saturation literal) ;
 *              u = 3 -> NaN
 *
 * post-ReLU codes (SLFP_FMT_SLFP34_RELU / SLFP_FMT_SFP33_RELU), unsigned byte c, value >= 0:
 *     c = (float32_bits(q) >> (22 - mbits)) - base, saturated to [0, 255]   (mbits 4 / 3; base 0xF5F / 0x7AF)
 *     i.e. the truncated bit pattern with ONE more mantissa bit than the grid keeps.  c = 0 -> 0.0 (q < 0.0625);
 *     u = c - 1 = E * 2^(mbits+1) + h:  E = 0 -> 0.125;  E >= 1 -> 2^(E-4) * grid[(h + 1) >> 1], clamped to the top
 *     grid value.  Class boundaries (0.0625, 0.125, saturation) are exact; the mantissa rounding (ties up) and the
 *     SLFP log converter [0,1,3,4,...,15,15] (utils/sfp_quant.py:88-89) are applied by the consumer's decode
 *     table.  +-1e-10 and the 15.32165 literal are represented by 0 and the top grid value (equal after the
 *     float16 rounding of the tensor-core operand).
 */
#ifndef SLFP_B200_H_
#define SLFP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLFP_B200_VERSION 200

typedef void *slfp_stream_t; /* cudaStream_t */

enum {
    SLFP_FMT_SFP33 = 0,      /* utils/sfp_quant.py:14-30, 63-78   q_bit 7, weights and activations */
    SLFP_FMT_SLFP34_ACT = 1, /* utils/sfp_quant.py:80-96          q_bit 8 activations              */
    SLFP_FMT_SLFP34_WGT = 2, /* utils/sfp_quant.py:32-47          q_bit 8 weights                  */
    SLFP_FMT_SFP44_OUT = 3,  /* utils/sfp_quant.py:105-127        layer-out quantizer (fp32 only)  */
    /* Unsigned post-ReLU activation codes exchanged by FUSED layers (quantize-on-store of a conv
     * whose epilogue ends in a ReLU; see "post-ReLU codes" below).  Same grids as SLFP34_ACT / SFP33. */
    SLFP_FMT_SLFP34_RELU = 4,
    SLFP_FMT_SFP33_RELU = 5,
    /* Signed variant of SFP33_RELU for fused layers WITHOUT a ReLU (ShuffleNetV2's depthwise conv -> BN -> next conv):
     * bit 7 = sign, low 7 bits = min(c, 127) of the post-ReLU code of |q| (c = 127 and 128 both decode to the top
     * value 15, so nothing is lost).  SFP<3,3> only - the SLFP<3,4> half-step code needs 9 bits with a sign. */
    SLFP_FMT_SFP33_SFAST = 6,
    /* SFP<3,3> activations stored as OCP e4m3 bytes [s:1][e:4][m:3] (bias 7).  Every SFP<3,3> value (1 + m/8) * 2^e,
     * e in [-3, 3], and the clamps 0.125 / 15 are exactly representable, +-1e-10 is stored as +-0
     * (what it becomes as a tensor-core operand anyway).  The code IS the tensor-core operand: dense layers with
     * c_phys % 64 == 0 run tcgen05.mma kind::f8f6f4 straight on the TMA-loaded code tile (no decode warps, no operand
     * rounding, 2x the f16 rate - SLFP_CONV_E4M3_OPERANDS); the encoder is one cvt.rn.satfinite.e4m3x2.f32 behind the
     * reference's clamps, i.e. round-half-EVEN like utils/sfp_quant.py:69 (the other fused formats round ties up). */
    SLFP_FMT_E4M3 = 7,
    /* Activations stored as the float16 IMAGE of the quantized value - float16_rn(decode(code)), the very number the
     * dense kernel's decode table hands to the tensor core - 2 bytes per element, NHWC, c_phys % 64 == 0.  The tensor IS
     * the tcgen05 kind::f16 A operand: TMA im2col drops it into 128-byte-swizzled shared memory and the MMA reads it
     * there (no decode table, no decode warps).  Written by a fused epilogue with SlfpEpilogue.store_f16; a layer's
     * result is bit-identical to the same layer fed with the codes.  Trades 1 B/element of HBM traffic for the
     * per-(tap x output-tile) table look-ups of the consumer - used where the consumer is decode-bound (3x3 layers). */
    SLFP_FMT_F16Q = 8
};

enum { SLFP_ACT_STL = 0, SLFP_ACT_SWISH = 1, SLFP_ACT_SIGMOID = 2 };
enum { SLFP_SGD_NORMAL = 0, SLFP_SGD_DSGD = 1, SLFP_SGD_SSGD = 2 };

enum {
    SLFP_ERR_BAD_ARG = 10001,
    SLFP_ERR_UNSUPPORTED = 10002,
    SLFP_ERR_NO_DEVICE = 10003,
    SLFP_ERR_DRIVER = 10004
};

/* flags of slfp_quantize_f32 */
#define SLFP_Q_LAYEROUT_ZERO_IS_ZERO 1u /* SFP44_OUT: map exact 0 to 0 instead of the reference's NaN */

int slfp_version(void);
const char *slfp_last_error(void);
/* sha256 prefix of the sources this binary was compiled from (cnns_slfp_quantization_b200/build.py: source_hash);
 * the Python binding refuses a library whose id differs from the sources next to it. */
const char *slfp_build_id(void);

/* ---------------------------------------------------------------------------------------------
 * Quantizers.  Replaces quantize_weight(k)/quantize_act(k)/quantize_layerout(k).forward
 * (utils/sfp_quant.py:7-48, 56-97, 105-127) together with the pre-scale division
 * `input/self.Ka`, `self.weight/self.Kw` of utils/conv2d_func.py:21-22.
 *   y = x / k_div (IEEE float32 division), then round onto the format's grid.
 *   codes (n bytes) and/or fakeq (n floats) and/or f16 (n halves) may be NULL; at least one is not.
 * One fused pass: 4 B read + (1 | 4 | 2) B written per element.
 * ------------------------------------------------------------------------------------------- */
int slfp_quantize_f32(const float *x, size_t n, float k_div, int fmt, unsigned flags,
                      uint8_t *codes, float *fakeq, void *f16, slfp_stream_t stream);

/* Dynamic max-scaling form of the same quantizer (north_star: "tensor abs-max reduction, max-scaling, round to
 * nearest"; the reference derives K = max|x| / 15.5 offline, cifar100_train_eval.py:213-277, nets_cifar/mobilenetv1.py:15).
 * K is NOT a host value here: the kernel reads *absmax (DEVICE float, e.g. written by slfp_absmax_f32 and optionally
 * max-allreduced over ranks) when it starts and uses K = float32(double(*absmax) / divisor) - the reference's float64
 * Python division rounded to the float32 its tensor arithmetic uses.  No host round trip: abs-max -> [allreduce(MAX)]
 * -> quantize is three asynchronous launches on one stream and can be captured in a CUDA graph.  k_out (DEVICE float,
 * may be NULL) receives the K that was used. */
int slfp_quantize_dyn_f32(const float *x, size_t n, const float *absmax, double divisor, int fmt, unsigned flags,
                          uint8_t *codes, float *fakeq, void *f16, float *k_out, slfp_stream_t stream);

/* Same quantizer for an NHWC activation tensor whose code tensor has a padded channel count
 * (c_phys >= c; pad channels receive code 0 = exact zero): npix pixels of c floats -> npix * c_phys
 * codes.  This is the layout slfp_conv2d_fwd consumes. */
int slfp_quantize_nhwc_f32(const float *x, size_t npix, int c, int c_phys, float k_div, int fmt,
                           uint8_t *codes, slfp_stream_t stream);

/* Same quantizer reading an NCHW float32 tensor (the network input as the reference's DataLoader
 * delivers it: n x c x hw) and writing NHWC codes with c_phys (multiple of 4) channels. */
int slfp_quantize_nchw_f32(const float *x, int n, int c, size_t hw, int c_phys, float k_div, int fmt,
                           uint8_t *codes, slfp_stream_t stream);

/* Space-to-depth variant for a stride-2 stem (ResNet's 7x7/2 on 3 channels): writes NHWC codes of the
 * 2x2-folded image, [n, h/2, w/2, c_phys] with channel (dy*2 + dx)*c + ch for pixel (2y+dy, 2x+dx) and
 * c_phys = round_up(4c, 16).  The stride-2 RxR convolution on c channels is then a stride-1
 * ceil((R+1)/2)^2 convolution on 4c channels whose 16-byte channel vectors the TMA im2col path can fetch
 * (see DESIGN.md "stem").  h and w must be even. */
int slfp_quantize_nchw_s2d_f32(const float *x, int n, int c, int h, int w, int c_phys, float k_div, int fmt,
                               uint8_t *codes, slfp_stream_t stream);

/* The same space-to-depth input quantizer for an RGB image (c = 3, c_phys = 16) writing SLFP_FMT_F16Q - the float16 image
 * of each code's value, 32 bytes per folded pixel - into a physically ZERO-PADDED tensor [n, hp, wp, 16] at offset
 * (pad_top, pad_left); the caller zeroes the buffer once, the kernel only writes the interior.  This is the input of the
 * SLFP_CONV_FOLD_W stem (below): padding is materialised because the folded rows overlap.  w % 4 == 0, h even. */
int slfp_quantize_nchw_s2d_f16q(const float *x, int n, int h, int w, float k_div, int fmt, int pad_top, int pad_left,
                                int hp, int wp, void *out_f16, slfp_stream_t stream);

/* Input quantizer of a 3x3 / stride 1 / padding 1 RGB stem (the CIFAR nets) writing the layer's im2col matrix directly:
 * out[n, h, w, 64] SLFP_FMT_F16Q halves, entry (r * 3 + s) * 4 + c = quantize(x[n, c, y + r - 1, x + s - 1] / k_div) (0 outside
 * the image and in the unused entries) - the order of the KRSC weight row of the c_phys = 4 stem, so the stem runs as a plain
 * 1x1 SLFP_FMT_F16Q layer with c_phys = 64 (one K block, no decode) on the weight operand slfp_prepare_weights writes for it. */
int slfp_quantize_nchw_im2col3x3_f16q(const float *x, int n, int h, int w, float k_div, int fmt, void *out_f16,
                                      slfp_stream_t stream);

/* Gather + quantize: the activation quantizer fed from SEVERAL float16 NHWC tensors through a per-channel table -
 * codes[p, j] = encode(float(src_j[p * stride_j + ch_j]) / k_div) for j < c, code 0 for c <= j < c_phys.  This is how
 * the fused pipeline evaluates torch.split / torch.cat / channel_shuffle (nets_cifar/shufflenet_v2.py:20-45, 100-115)
 * without moving data: the logical channel order is an index map resolved by the consumer's quantizer.  `table` is a
 * DEVICE array of c entries. */
typedef struct {
    const void *src;     /* float16 NHWC tensor                       */
    int stride;          /* its channels per pixel (elements)         */
    int ch;              /* the channel to read                       */
} SlfpGatherChan;
int slfp_gather_quantize_f16(const SlfpGatherChan *table, size_t npix, int c, int c_phys, float k_div, int fmt,
                             uint8_t *codes, slfp_stream_t stream);

/* The same gather-quantizer driven by RUNS instead of single channels - the form the fused ShuffleNetV2 plan uses: after
 * split / cat / shuffle every source tensor contributes ONE run of consecutive channels whose logical positions are an
 * arithmetic sequence (dst_start + i * dst_step, the step doubling with every unit the values were passed through).
 * A CTA stages the runs' covering 16-byte chunks for a tile of pixels in shared memory (sector-aligned 128-bit loads), then
 * every thread assembles one 32-bit word of codes from the staged float16 values and writes it straight to global memory.
 * Sources hold post-layerout, post-ReLU values (non-negative float16 <= 248).  `runs` is a DEVICE array; positions not
 * covered by a run (the pad channels) get code 0. */
typedef struct {
    const void *src;     /* float16 NHWC tensor                                          */
    int stride;          /* its channels per pixel (elements)                            */
    int ch0;             /* first channel of the run                                     */
    int len;             /* channels in the run                                          */
    int dst_start;       /* logical (output) channel of the run's first element          */
    int dst_step;        /* distance between consecutive elements in the output          */
    unsigned magic, shift; /* magic-number division by nck = the number of 16-byte (8-channel) aligned chunks that cover
                              [ch0, ch0 + len): idx / nck == umulhi(idx, magic) >> shift; the caller fills them with
                              slfp_magic_u32(nck, ...).  stride must be a multiple of 8, src 16-byte aligned. */
} SlfpGatherRun;
void slfp_magic_u32(unsigned d, unsigned *magic, unsigned *shift);
/* c: logical channels (the union of the runs' positions); stage_bytes_per_pixel: 16 * the sum of the runs' chunk counts. */
int slfp_gather_quantize_runs_f16(const SlfpGatherRun *runs, int n_runs, size_t npix, int c, int c_phys,
                                  int stage_bytes_per_pixel, float k_div, int fmt, uint8_t *codes, slfp_stream_t stream);

/* codes -> float32 (exactly the value the reference's fake-quant tensor would hold) */
int slfp_dequantize(const uint8_t *codes, size_t n, int fmt, float *out, slfp_stream_t stream);

/* max(|x|) into *max_out (device float).  Replaces torch.max(torch.abs(torch.cat(list))) of the
 * calibration pass, cifar100_train_eval.py:261-271.  The kernel folds its result into *max_out
 * with an atomic max, so a caller accumulating over batches zeroes it once; init_zero != 0 makes
 * the call zero it first. */
int slfp_absmax_f32(const float *x, size_t n, float *max_out, int init_zero, slfp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Convolution.  Replaces Conv2d_Q.forward, utils/conv2d_func.py:20-25 (no bias) and :41-47
 * (bias), and Linear_Q.forward :60-65 (a linear layer is the 1x1 case with h = w = 1).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int n, h, w, c;       /* input  NHWC; c = logical input channels                          */
    int c_phys;           /* physical (padded) channel count of the code tensor, multiple of 16
                             (4 allowed when c <= 4: the 3-channel stem)                        */
    int k;                /* output channels                                                   */
    int r, s;             /* filter height / width                                             */
    int stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
    int groups;           /* 1 = dense implicit GEMM (tcgen05); c == k == groups = depthwise     */
    int fmt;              /* code layout of the activation codes: SLFP_FMT_SLFP34_ACT (q_bit 8) or
                             SLFP_FMT_SFP33 (q_bit 7); dense convs also take the post-ReLU formats  */
    int pad_h_extra;      /* bottom / right padding minus top / left padding (0 = symmetric, as in      */
    int pad_w_extra;      /* PyTorch).  Non-zero only for the space-to-depth form of a strided stem:     */
                          /* dense, c_phys % 16 == 0.                                                   */
    int flags;            /* SLFP_CONV_* bits, 0 = default                                               */
} SlfpConvDesc;

/* High-fidelity (split-operand) mode of the dense forward: both tensor-core operands are float16 PAIRS hi + lo
 * (hi = float16(v), lo = float16(v - hi): 2^-21 .. 2^-22 relative instead of 2^-12 for the irrational SLFP grid values)
 * and the GEMM runs as three passes over K, x_hi*w_hi + x_hi*w_lo + x_lo*w_hi, into the same float32 accumulator -
 * conv outputs then agree with the reference's float32 arithmetic to accumulation order.  About 3x the tensor and
 * decode work; dense layers with c_phys % 16 == 0 only.  The weight operand has rows of 2 * slfp_conv_wpitch() halves,
 * [hi | lo] (SlfpWeightJob.lo_offset = slfp_conv_wpitch()). */
#define SLFP_CONV_SPLIT_OPERANDS 1
/* The weight operand is e4m3 BYTES (KRSC, row pitch slfp_conv_wpitch() bytes) instead of float16, the activation codes
 * are SLFP_FMT_E4M3: kind::f8f6f4 MMA on the codes themselves.  q_bit 7 (SFP<3,3>) only, dense, c_phys % 64 == 0.
 * slfp_prepare_weights* writes that operand through the w_f16 pointer when the descriptor carries this flag. */
#define SLFP_CONV_E4M3_OPERANDS 2
/* Width-folded input (the ResNet stem after space-to-depth): the descriptor describes a VIRTUAL tensor [n, h, w, c_phys = 64]
 * whose pixel (y, x) is the concatenation of the four physical pixels (y, x .. x + 3) of a zero-padded SLFP_FMT_F16Q tensor
 * [n, h, w + 3, 16] - pixel pitch 16 elements, so consecutive virtual pixels OVERLAP.  An R x 4 stride-1 convolution on 16
 * channels becomes an R x 1 convolution on 64: one 128-byte TMA row carries the four horizontal taps of a pixel, K blocks
 * are whole filter rows, and the A tile needs no decode.  The weight operand is the KRSC tensor of the R x 4 filter on 16
 * channels (same memory).  Requires fmt == SLFP_FMT_F16Q, c_phys == 64, s == 1, stride 1, no padding, no dilation. */
#define SLFP_CONV_FOLD_W 4

typedef struct {
    const float *bias_q;   /* [k] added to the accumulator BEFORE the post-scale (bias/Ka/Kw,
                              conv2d_func.py:44), or NULL                                        */
    float post_a, post_b;  /* y = ((acc + bias_q) * post_a) * post_b   (Ka then Kw, :24 / :46)   */
    const float *ch_scale; /* optional per-channel affine applied after the post-scale (an eval */
    const float *ch_shift; /* BatchNorm folded by the caller): y = y*ch_scale[k] + ch_shift[k]    */
    const void *residual;  /* optional NHWC tensor added after the affine, or NULL               */
    int residual_f16;      /* 0: residual is float32, 1: float16                                 */
    int relu;              /* apply max(y, 0) last                                               */
    /* outputs: any subset, all NHWC with k channels (codes: k_phys_out channels, zero padded)  */
    float *y_f32;
    void *y_f16;
    uint8_t *y_codes;      /* quantize-on-store for the next layer: encode(y / next_k_div)       */
    float next_k_div;
    int next_fmt;
    int k_phys_out;        /* physical channel count of y_codes (>= k, multiple of 16)           */
    uint8_t *y_codes2;     /* second consumer with a different Ka (e.g. a downsample branch)     */
    float next_k_div2;
    /* Folded form used by the fused eval pipeline: when ch_mul != NULL the per-channel affine
     *     y = acc * ch_mul[k] + ch_add[k]
     * REPLACES bias_q / post_a / post_b / ch_scale / ch_shift (the caller folds them in float64:
     * ch_mul = Ka*Kw*bn_scale, ch_add = bias_q*Ka*Kw*bn_scale + bn_shift).  One FMA per element;
     * differs from the unfolded order by float32 rounding only. */
    const float *ch_mul;
    const float *ch_add;
    /* quantize_layerout (SFP<4,4>, utils/sfp_quant.py:105-127) applied to y after the affine / residual and BEFORE the
     * ReLU - the reference's conv -> BatchNorm -> layerout_quantize_func -> ReLU chains (nets_cifar/shufflenet_v2.py:
     * 64-72, MobileNetV1_swish, VGG16_gelu).  0 = off; 1 = the reference's arithmetic including NaN at exact 0 (its
     * `2^(-8)` is XOR: no low clamp, 0 * NaN); 2 = exact 0 stays 0 (the evidently intended value).  Generic epilogue. */
    int layerout;
    /* bit 0: y_codes receives SLFP_FMT_F16Q halves (k_phys_out per pixel) instead of the code bytes of next_fmt - the
     * float16 image of exactly the code the byte path would have stored (post-ReLU formats, codes-only fast epilogue). */
    int store_f16;
} SlfpEpilogue;

/* Weight preparation: replaces `self.quantize_weight(self.weight/self.Kw)` (conv2d_func.py:22):
 * quantizes the OIHW float32 parameter, and emits any of
 *   w_f16   : KRSC float16 operand for the tensor-core path, row pitch slfp_conv_wpitch(desc) halves
 *   w_codes : KRSC 8-bit codes (same pitch, in bytes)
 *   w_fakeq : OIHW float32 fake-quant tensor (the reference's `weight_q` calibration tap)
 * w_stride_{o,c,r,s}: element strides of the input parameter (so channels_last params work). */
size_t slfp_conv_wpitch(const SlfpConvDesc *desc);
int slfp_prepare_weights(const SlfpConvDesc *desc, const float *w, long long w_stride_o,
                         long long w_stride_c, long long w_stride_r, long long w_stride_s,
                         float kw, int wfmt, void *w_f16, uint8_t *w_codes, float *w_fakeq,
                         slfp_stream_t stream);

/* The same for n layers in ONE launch (the reference re-quantizes every layer's weights on every forward,
 * conv2d_func.py:22).  The job table is host memory and is read before the call returns.  out_pitch / out_offset
 * place a layer's rows inside a wider operand (0 = the layer's own pitch): the fused block tail concatenates the
 * last conv of a residual block and its downsample conv along K; row_scale (device, [k], or NULL) multiplies the
 * float16 operand of output channel k (the ratio of the two branches' folded BatchNorm scales). */
typedef struct {
    const SlfpConvDesc *desc;
    const float *w;            /* device: the OIHW float32 parameter                                   */
    long long w_stride[4];     /* element strides (o, c, r, s)                                          */
    float kw;
    void *w_f16;               /* device outputs, either may be NULL                                   */
    uint8_t *w_codes;
    size_t out_pitch, out_offset;
    const float *row_scale;
    size_t lo_offset;          /* != 0: also write lo = float16(w_q - float(w_f16)) at out_offset + lo_offset inside the
                                  row (the split-operand mode, SLFP_CONV_SPLIT_OPERANDS); 0 = no lo part             */
} SlfpWeightJob;
int slfp_prepare_weights_jobs(int n, const SlfpWeightJob *host_jobs, int wfmt, slfp_stream_t stream);

/* Forward on codes.  x_codes NHWC [n,h,w,c_phys]; w_f16 from slfp_prepare_weights (dense) or
 * w_f32 KRSC float32 for the depthwise / grouped stencil path. */
int slfp_conv2d_fwd(const SlfpConvDesc *desc, const uint8_t *x_codes, const void *w_prepared,
                    const SlfpEpilogue *epi, slfp_stream_t stream);

/* Fused residual-block tail: the last convolution of a block and the block's downsample convolution as ONE
 * GEMM over the concatenated K dimension,
 *     y = epilogue( sum_k A1[m,k] W1[n,k] + sum_k A2[m,k] W2'[n,k] ),
 * desc1 / x1_codes: the main branch (e.g. conv3, 1x1); desc2 / x2_codes: the downsample branch (1x1, any stride,
 * no padding) - same n, same output size, same k, same code format.  w_cat: [k, pitch1 + pitch2] float16 from
 * slfp_prepare_weights_jobs (W2' carries the ratio of the two branches' scales; the epilogue's folded affine
 * ch_mul / ch_add belongs to branch 1 with the two shifts added).  Removes the downsample launch, its float16
 * output and the residual read of it.  Replaces nets_imgnet/resnet50.py:80-88 (`out = bn3(conv3(out))`,
 * `identity = downsample(x)`, `out += identity`, `relu`). */
int slfp_conv2d_fwd_dual(const SlfpConvDesc *desc1, const uint8_t *x1_codes, const SlfpConvDesc *desc2,
                         const uint8_t *x2_codes, const void *w_cat, const SlfpEpilogue *epi, slfp_stream_t stream);

/* Backward of Conv2d_Q.forward with the identity STE (utils/sfp_quant.py:50-53, 99-102):
 *   dx = dgrad(gy*Ka*Kw, w_q)/Ka      dw = wgrad(gy*Ka*Kw, x_q)/Kw      db = sum(gy)
 * gy NHWC float32 [n,ho,wo,k]; x_codes NHWC; w_codes KRSC (pitch slfp_conv_wpitch);
 * dx NHWC float32 [n,h,w,c]; dw float32 with the given element strides; any output may be NULL. */
int slfp_conv2d_bwd(const SlfpConvDesc *desc, const float *gy, const uint8_t *x_codes,
                    const uint8_t *w_codes, int wfmt, float ka, float kw, float *dx, float *dw,
                    long long dw_stride_o, long long dw_stride_c, long long dw_stride_r,
                    long long dw_stride_s, float *db, slfp_stream_t stream);

/* The same backward as two implicit GEMMs on tcgen05 tensor cores (csrc/conv_bwd_sm100.cu): float16 operands
 * (gy scaled by a power of two taken from max|gy|, the float16 images of the saved codes), float32 accumulation.
 * dgrad splits a strided convolution into its output-parity classes (no zero-dilated gradient); wgrad reduces
 * over output pixels with both NHWC operands MN-major and split-K partial sums added by red.global.v4.f32, so dw
 * is not bit-reproducible run to run.  The library never allocates: the caller passes a 256-byte-aligned scratch
 * buffer of slfp_conv2d_bwd_workspace_size() bytes (0 = shape not covered: grouped convolutions, more than 32
 * filter taps; slfp_conv2d_bwd_ws then runs the direct kernels of slfp_conv2d_bwd, as it does for a dw whose
 * c_phys is not a multiple of 64). */
size_t slfp_conv2d_bwd_workspace_size(const SlfpConvDesc *desc, int need_dx, int need_dw);
int slfp_conv2d_bwd_ws(const SlfpConvDesc *desc, const float *gy, const uint8_t *x_codes,
                       const uint8_t *w_codes, int wfmt, float ka, float kw, float *dx, float *dw,
                       long long dw_stride_o, long long dw_stride_c, long long dw_stride_r,
                       long long dw_stride_s, float *db, void *workspace, size_t workspace_bytes,
                       slfp_stream_t stream);
/* slfp_conv2d_bwd_ws with max |gy| supplied by the caller (device scalar written by the kernel that produced gy, e.g.
 * slfp_bn_act_bwd): the tensor-core path scales gy by it for its float16 operand and skips its own abs-max pass.
 * gy_absmax == NULL: same as slfp_conv2d_bwd_ws. */
int slfp_conv2d_bwd_ws_absmax(const SlfpConvDesc *desc, const float *gy, const float *gy_absmax, const uint8_t *x_codes,
                              const uint8_t *w_codes, int wfmt, float ka, float kw, float *dx, float *dw, long long so,
                              long long sc, long long sr, long long ss, float *db, void *workspace, size_t workspace_bytes,
                              slfp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Activations: STLFunction / STL, Swish, Sigmoid (utils/activation_func.py:6-36).
 * bwd: STL clips the incoming gradient to [-1,1] by its own magnitude (:16); Swish / Sigmoid are
 * the autograd derivatives and need x.
 * ------------------------------------------------------------------------------------------- */
int slfp_act_fwd(const float *x, size_t n, int kind, float *y, slfp_stream_t stream);
int slfp_act_bwd(const float *x, const float *gy, size_t n, int kind, float *gx, slfp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Revised SGD: DSGD.step / SSGD.step / NormalSGD.step (utils/optimizer.py:30-73, 98-132,
 * 154-190), one fused multi-tensor launch for a whole parameter group.
 * host_* arrays have n_tensors entries and are read before the call returns.  grad is updated in
 * place by the weight-decay term like the reference (:46).  Hyper-parameters are doubles (Python
 * floats) and are rounded to float32 the way the reference's tensor-scalar ops round them:
 * float32(-lr), float32(momentum), float32(1 - dampening), float32(weight_decay).  q_fmt: SLFP_FMT_SLFP34_WGT,
 * SLFP_FMT_SFP33, or -1 for q_bit 32 (identity quantizer).
 * ------------------------------------------------------------------------------------------- */
int slfp_sgd_step(int n_tensors, float *const *host_params, float *const *host_grads,
                  float *const *host_bufs, const size_t *host_sizes, int mode, int q_fmt,
                  double lr, double momentum, double dampening, double weight_decay, int nesterov,
                  int first_step, slfp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Small NHWC helpers used between quantized layers by the fused eval pipeline.
 * ------------------------------------------------------------------------------------------- */
/* 2-D max pooling directly on activation codes (the quantizer is monotone, so
 * quantize(maxpool(x)) == maxpool(quantize(x)) bit for bit).  pad positions never win. */
int slfp_maxpool_codes(const uint8_t *x, int n, int h, int w, int c_phys, int fmt, int kh, int kw_,
                       int stride, int pad, uint8_t *y, slfp_stream_t stream);
/* global average pool NHWC float16/float32 -> [n, c] float32 */
int slfp_avgpool_nhwc(const void *x, int is_f16, int n, int hw, int c, float *y, slfp_stream_t stream);
/* The same pool on a float16 NHWC tensor with the classifier's activation quantizer fused (replaces
 * `x = avgpool(x); x = flatten(x, 1)` followed by the act quantizer of `fc`, nets_imgnet/resnet50.py:242-244 and
 * utils/conv2d_func.py:61): y (optional) receives the float32 means [n, c], codes (optional) encode(mean / k_div) in
 * fmt = SLFP_FMT_SFP33 | SLFP_FMT_SLFP34_ACT, [n, c] bytes - bit-identical to slfp_avgpool_nhwc + slfp_quantize_nhwc_f32.
 * c % 8 == 0, 16-byte aligned tensors. */
int slfp_avgpool_quantize_nhwc_f16(const void *x, int n, int hw, int c, float *y, float k_div, int fmt, uint8_t *codes,
                                   slfp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * QAT step (BASELINE config 4): training-mode BatchNorm2d + residual add + ReLU of the caller net, fused, float32 NHWC
 * [m = n*h*w rows, c channels], c % 4 == 0.  Replaces the stock sequences of nets_imgnet/resnet50.py:71-88
 * (`relu(bn(conv(x)))`, `bn3(conv3(out)); out += identity; relu(out)`) around the quantized convolutions
 * (cifar100_train_eval.py:170-179 runs them forward and backward every step).
 *   y = relu?( (x - mean_c) / sqrt(var_c + eps) * gamma_c + beta_c + residual? )     batch statistics over the m rows
 * running_mean / running_var (optional pair) are updated like nn.BatchNorm2d (momentum, unbiased variance).
 * save_mean / save_invstd [c] are outputs for the backward.  workspace: slfp_bn_act_workspace_floats(c) floats, 16-byte
 * aligned, ZERO-FILLED once by the caller (the kernels leave it zeroed); one workspace per stream.  coef: scratch
 * vector of 4 * c floats, 16-byte aligned, no initialisation (per call: it is read by the second kernel of the call).
 * Backward: dx [m, c], d_residual [m, c] (optional: the gradient of the residual input), dgamma / dbeta [c].  The ReLU
 * mask comes from y; for a layer without residual y may be NULL and beta given instead: the mask is then recomputed
 * from x with the forward's own scale / shift arithmetic (one read pass less, y need not be kept).  dx_absmax
 * (optional device scalar) receives max |dx| - what slfp_absmax_f32(dx) would compute in a pass of its own - for
 * slfp_conv2d_bwd_ws_absmax of the convolution that produced x. */
size_t slfp_bn_act_workspace_floats(int c);
int slfp_bn_act_fwd_train(const float *x, size_t m, int c, const float *gamma, const float *beta, const float *residual,
                          int relu, float eps, float momentum, float *running_mean, float *running_var, float *y,
                          float *save_mean, float *save_invstd, float *workspace, float *coef, slfp_stream_t stream);
/* The same forward (ReLU always applied) that ALSO emits the activation codes of y for its consumers - the quantizer
 * `quantize_act(y / Ka)` of the next Conv2d_Q (utils/conv2d_func.py:21), bit-identical to slfp_quantize_nhwc_f32(y) -
 * for n_codes = 1 or 2 scales k_div[j] (a block output feeds conv1 and the downsample conv of the next block):
 * codes[j] is [m, c] bytes (c_phys == c), fmt = SLFP_FMT_SFP33 | SLFP_FMT_SLFP34_ACT.  k_div / codes are HOST arrays. */
int slfp_bn_act_fwd_train_quant(const float *x, size_t m, int c, const float *gamma, const float *beta, const float *residual,
                                float eps, float momentum, float *running_mean, float *running_var, float *y,
                                float *save_mean, float *save_invstd, float *workspace, float *coef, int fmt, int n_codes,
                                const float *k_div, uint8_t *const *codes, slfp_stream_t stream);
int slfp_bn_act_bwd(const float *gy, const float *x, const float *y, size_t m, int c, const float *gamma, const float *beta,
                    const float *save_mean, const float *save_invstd, int relu, float *dx, float *d_residual,
                    float *dgamma, float *dbeta, float *workspace, float *coef, float *dx_absmax, slfp_stream_t stream);

/* The 3x3 / stride 2 / padding 1 max-pool of a training step on float32 NHWC [n, h, w, c] (c % 4 == 0): replaces
 * `self.maxpool` (nn.MaxPool2d(3, 2, 1), nets_imgnet/resnet50.py:237) forward and backward.  idx [n, ho, wo, c] bytes:
 * the window position (row * 3 + column) of each maximum, PyTorch's tie / NaN rule.  The backward is a deterministic
 * gather (no atomics, dx needs no zero-fill). */
int slfp_maxpool3x3s2_fwd_f32(const float *x, int n, int h, int w, int c, float *y, uint8_t *idx, slfp_stream_t stream);
int slfp_maxpool3x3s2_bwd_f32(const float *gy, const uint8_t *idx, int n, int h, int w, int c, float *dx, slfp_stream_t stream);

/* Debug aid: a host-mapped buffer (>= 16 bytes) into which a timed-out barrier wait of the dense conv kernel
 * records which wait it was before it traps (the kernels never hang: every wait is bounded).  NULL removes it. */
int slfp_debug_set_buffer(void *host_mapped_device_ptr);

/* ---------------------------------------------------------------------------------------------
 * Host-buffer convenience entry (the end-to-end path a non-torch caller binds): quantizes a HOST
 * float32 buffer into HOST codes through a device round trip on an internal stream and
 * synchronises.  Used by the C smoke program and by bench.py's e2e leg for the quantizer metric.
 * ------------------------------------------------------------------------------------------- */
int slfp_quantize_host_f32(const float *host_x, size_t n, float k_div, int fmt,
                           uint8_t *host_codes, float *host_fakeq);

#ifdef __cplusplus
}
#endif
#endif /* SLFP_B200_H_ */
