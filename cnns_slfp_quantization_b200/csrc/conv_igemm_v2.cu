// conv_igemm_v2.cu -- dense SLFP/SFP convolution forward, warp-specialised implicit GEMM for sm_100a.
//
// Replaces Conv2d_Q.forward / Linear_Q.forward of the reference (utils/conv2d_func.py:20-25, 41-47,
// 60-65).  Same math as conv_igemm_sm100.cu (which stays as the path for the 4-channel stem layout):
//      D[m, n] = sum_k A[m, k] * B[n, k]     m = output pixel, n = output channel, k = (r, s, c)
// with A = the float16 image of the 8-bit activation codes (NHWC, c_phys % 16 == 0) and B = the float16
// image of weight_q (KRSC).  What changed is who does what - the first kernel was bound by CUDA-core
// instruction issue (profiles/r01_launches_baseline.md), so every role below is as thin as it can be:
//
//   warp 0      code producer: ONE thread issues TMA im2col loads (cp.async.bulk.tensor.4d...im2col) of the
//               128 pixels x 64 channels code tile of a K block - stride, padding, dilation, image borders
//               (zero fill = code 0 = 0.0) and the M tail are the TMA unit's business - into a 4-deep ring
//   warp 1      weight producer: TMA tile loads of the BLOCK_N x 64 float16 weight tile, 128B swizzle
//   warp 2      MMA issuer: one thread, 4 x tcgen05.mma (M128 x BLOCK_N x K16, kind::f16) per K block,
//               accumulators double-buffered in TMEM; tcgen05.commit frees the smem stage
//   warps 4-11  decode: codes (smem) -> float16 through a bank-conflict-free lookup table (one 32-bit
//               LDS per code; the table also applies the SLFP log converter for the post-ReLU code
//               formats) -> the 128B-swizzled K-major A tile (smem) the UMMA descriptor reads
//   warps 12-19 epilogue: tcgen05.ld -> folded per-channel affine (+ residual, ReLU) -> float32 / float16
//               and quantize-on-store.  For the post-ReLU code formats the encoder is 2 integer
//               instructions per element plus a saturating pack (slfp_common.cuh encode_relu_fast).
// Barriers: code_full/empty[4], ab_full/empty[stages], tmem_full/empty[2].
#include <cudaTypedefs.h>

#include <stdlib.h>
#include <string.h>

#include "slfp_common.cuh"
#include "sm100_ptx.cuh"

// Role profiler (build with -DSLFP_ROLE_PROFILE, tools/role_profile.py): clock64() around the pipeline waits of
// each role's lead thread, summed over CTAs into the slfp_debug_set_buffer() array from word 8 on.
#ifdef SLFP_ROLE_PROFILE
#define PROF_VARS unsigned long long prof_a = 0, prof_b = 0, prof_c = 0; const long long prof_t0 = clock64()
#define PROF(n, ...) { const long long prof_t = clock64(); __VA_ARGS__; prof_##n += (unsigned long long)(clock64() - prof_t); }
#define PROF_FLUSH(slot) if (ptx::g_slfp_dbg) { atomicAdd(ptx::g_slfp_dbg + (slot), prof_a); atomicAdd(ptx::g_slfp_dbg + (slot) + 1, prof_b); \
        atomicAdd(ptx::g_slfp_dbg + (slot) + 2, prof_c); atomicAdd(ptx::g_slfp_dbg + (slot) + 3, (unsigned long long)(clock64() - prof_t0)); }
#else
#define PROF_VARS
#define PROF(n, ...) { __VA_ARGS__; }
#define PROF_FLUSH(slot)
#endif

#include <type_traits>

namespace slfp {
namespace v2 {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kCodeBytes = kBM * kBK;                  // 8 KB of codes per K block
constexpr int kABytes = kBM * kBK * 2;                 // 16 KB float16 A tile
constexpr int kLutBytes = 256 * 32 * 4;                // code -> f16, one copy per bank
constexpr int kOutLutEntries = 258;                    // post-ReLU codes 0..257 (before the byte saturation) -> f16
constexpr int kWarpCode = 0, kWarpWgt = 1, kWarpMma = 2, kWarpCode2 = 3;
// 28 warps: 4 control (2 producers, MMA issuer, spare) + DW decode + (24 - DW) epilogue.  Two role splits, chosen
// per layer on the host: DW = 16 for decode-heavy layers (3x3, large K), DW = 8 for epilogue-heavy ones (the 1x1
// block tails: residual + float16 + codes).  Launched with 72 registers per thread, re-balanced with setmaxnreg:
//   DW 16:  4 x 40 + 16 x 56 +  8 x 120   (x 32 lanes) = 64 512 registers
//   DW  8:  4 x 40 +  8 x 56 + 16 x  88                = 64 512
constexpr int kCtrlWarps = 4, kDecWarp0 = 4, kWorkWarps = 24;
constexpr int kThreads = (kCtrlWarps + kWorkWarps) * 32;   // 896
constexpr int kRegsCtrl = 40;
template <int DW> struct Roles {
    static constexpr int kDecWarps = DW, kEpiWarps = kWorkWarps - DW, kEpiWarp0 = kDecWarp0 + DW;
    static constexpr int kGroups = kEpiWarps / 4;                      // column groups per TMEM lane quadrant
    static constexpr int kRegsDec = 56, kRegsEpi = DW == 16 ? 120 : 88;
    static constexpr int kChunksPerThread = 512 / (32 * DW);           // 16-byte code chunks per decode thread per K block
};

// n / d for n < 2^31 as umulhi(n, mg) >> sh (d > 1): the kernel divides tile indices by n_tiles / Ho*Wo / Wo several times per tile
// in every role; a hardware-less 32-bit division is ~30 dependent instructions each (the single-thread producer roles and
// the epilogue warps of the short-K layers were paying 100-150 of them per tile).
struct Magic {
    uint32_t d, mg, sh;
};
static Magic make_magic(uint32_t d) {
    Magic k{d, 0u, 0u};
    if (d > 1) {
        uint32_t lg = 31 - (uint32_t)__builtin_clz(d);
        if (d & (d - 1)) ++lg;
        const uint32_t pw = 31 + lg;
        k.mg = (uint32_t)(((1ull << pw) + d - 1) / d);
        k.sh = pw - 32;
    }
    return k;
}
__device__ __forceinline__ uint32_t mdiv(uint32_t n, const Magic& k) { return k.d == 1u ? n : (__umulhi(n, k.mg) >> k.sh); }

struct Params {
    Magic k_ntiles, k_howo, k_wo;
    uint32_t M;
    int Kout;
    int HoWo, Wo;
    int S, sh, sw, ph, pw, dh, dw;
    int num_kb, taps;
    int cblocks;                // c_phys / 64 when c_phys % 64 == 0, else 0 (16-channel pieces)
    int c16s;                   // c_phys / 16
    int m_tiles, n_tiles, num_tiles;
    int act_fmt;
    int plain_1x1;              // 1x1 / stride 1 / no padding: the code tile is a plain 2-D tile of the [M, C] matrix
    int nkb1;                   // K blocks taken from input 1 (== num_kb unless a second input is concatenated along K)
    int e4m3_out;               // next_fmt == SLFP_FMT_E4M3: quantize-on-store writes e4m3 bytes (round-half-even)
    int kb_base;                // K blocks of ONE pass over K (== num_kb; num_kb / 3 in the split-operand mode)
    int w_lo_col;               // split-operand mode: first column of the lo weights inside a weight row (= the row pitch)
    int sh2, sw2;               // second input (fused downsample branch): 1x1, stride (sh2, sw2), no padding, Cp % 64 == 0
    SlfpEpilogue epi;
    DivK next_div, next_div2;   // exact quantize-on-store (signed code formats)
    float rk1, rk2;             // 1 / next_k_div{,2} for the post-ReLU formats
    int stg_groups;             // staged epilogue: 2 = two groups of 8 epilogue warps work on alternate tiles (see run2), else 1
    int epi_mode;               // 0 generic; 1 codes-only fast path; 2 fast path with float16 residual / float16 output / two consumers
    float sc1, sc2;             // rk / 16 (fast paths quantize clamp(q/16, 0, 1))
};

// NODEC (SLFP_CONV_E4M3_OPERANDS): activation codes and weights are e4m3 bytes - the TMA-loaded code tile IS the A
// operand of tcgen05.mma kind::f8f6f4 (64-byte swizzle, two K = 32 instructions per 64-channel K block): no decode
// table, no decode warps, no A staging; the code ring and the weight ring share one barrier pair per stage.
// A16 (SLFP_FMT_F16Q input): the activation tensor already holds the float16 image of the quantized values, so the TMA-loaded
// [128 pixels x 64 channels] tile (128-byte rows, 128-byte swizzle) IS the A operand of tcgen05.mma kind::f16 - the NODEC
// structure with 2-byte elements.  Trades 1 B/element of HBM traffic for the decode work of the consumer.
// ACC1 (256-column decode tiles): ONE accumulator buffer of 256 TMEM columns + the decoded A ring in the remaining columns.
// The double-buffered 256-column form needs all 512 columns for accumulators and stages A in shared memory (measured ~1 180
// cycles per K block against ~640 with A in tensor memory); with a single buffer the epilogue of a tile no longer overlaps
// the next tile's MMAs, but a tile of a K-heavy layer (>= 16 K blocks) spends ~10x longer in its main loop than in its
// epilogue, and every decoded A tile now feeds 256 instead of 128 output columns: half the table look-ups per output.
template <int BLOCK_N, bool STG = false, bool NODEC = false, bool A16 = false, bool ACC1 = false>
struct Cfg {
    static_assert(!A16 || NODEC, "A16 is a no-decode variant");
    static_assert(!ACC1 || (BLOCK_N == 256 && !NODEC && !STG), "single accumulator buffer: the 256-column decode variant");
    static constexpr int kAccBufs = ACC1 ? 1 : 2;
    // BLOCK_N <= 128: the decoded A tile goes to TENSOR memory (tcgen05.st; the MMA reads A from TMEM), which takes
    // the A tile's write (16 KB) and the MMA's read of it (16 KB) per K block off the shared-memory pipe - the
    // measured bottleneck of the decode-heavy layers (profiles/r01_conv_v2.md).  BLOCK_N = 256 needs all 512 TMEM
    // columns for the double-buffered accumulator and keeps the A tile in shared memory.
    static constexpr bool kATmem = !NODEC && (BLOCK_N <= 128 || ACC1);
    static constexpr int kBBytes = (NODEC && !A16) ? BLOCK_N * kBK : BLOCK_N * kBK * 2;
    static constexpr int kCodeTile = A16 ? kABytes : kCodeBytes;      // bytes of one [128 x 64] activation tile in the code ring
    // Ring depths.  The code tiles and weight tiles arrive through TMA with ~1.5-2 us of latency under load; the
    // first version's 4 x 8 KB of codes in flight per SM left the decode warps waiting on the code barrier most
    // of the time (profiles/r01_conv_v2.md).  With A in TMEM the freed shared memory deepens both rings.
    static constexpr int kA16Stages = STG ? 4 : (BLOCK_N >= 256 ? 4 : (BLOCK_N >= 128 ? 6 : 8));      // 16 KB + BLOCK_N x 128 B per stage
    static constexpr int kStages = A16 ? kA16Stages
                                 : NODEC ? ((BLOCK_N >= 256 || STG) ? 6 : 8)
                                         : ((BLOCK_N >= 256 || STG) ? 3 : 4);          // weight (and A) stages (ACC1: 3 x 32 KB of weights)
    // The code ring depth must be a MULTIPLE of the number of decode groups (2 or 4), so that a code stage is always
    // consumed by the same group: TMA loads of different stages may land out of order, and a group that moved on to
    // K block j + CS while the load of K block j (same stage, other group) was still in flight would find the stage's
    // `full` barrier one phase behind - a parity wait then passes immediately (it cannot tell "phase n - 1 done" from
    // "phase n + 1 done"), the group decodes stale codes and its arrivals desynchronise the ring (seen as a rare
    // `unspecified launch failure` once the weights were re-quantized between steps, ~1 step in 1000).
    static constexpr int kCodeStages = A16 ? kA16Stages : ((BLOCK_N >= 256 || STG) ? 6 : 8);
    static_assert(!NODEC || kCodeStages == kStages, "no-decode variants: one ring, one barrier pair per stage");
    // STG: the epilogue's global traffic goes through shared-memory staging + TMA (see the staged epilogue below):
    // two float16 [128 x BLOCK_N] buffers (residual in / float16 out, in place) and two code tiles.
    static constexpr int kIoBytes = kBM * BLOCK_N * 2, kCoBytes = kBM * BLOCK_N;
    static constexpr int kStageBytes = STG ? 2 * kIoBytes + 2 * kCoBytes : 0;
    static constexpr int kATmemCol = kAccBufs * BLOCK_N;    // first TMEM column of the A ring (32 columns per stage)
    static constexpr int kTmemCols = kATmem ? (kAccBufs * BLOCK_N + kStages * 32 <= 256 ? 256 : 512) : 2 * BLOCK_N;
    static constexpr int kParBytes = (STG ? 2 : 1) * 2 * BLOCK_N * 4;   // this tile's per-channel mul / add (fast epilogues; staged: double buffered)
    static constexpr int kLutB = NODEC ? 0 : kLutBytes;
    // SlfpEpilogue.store_f16: code -> float16 table of the NEXT layer's format for the epilogue.  Decode variants: one copy per
    // bank like the decode table (their decode warps already load the shared-memory pipe); no-decode variants: 16 copies
    // (lanes l and l + 16 share a bank, at worst a 2-way conflict) so that it also fits next to the 256-column rings.
    // The staged and the 256-column decode variants have no room for it.
    static constexpr int kOutLutCopies = (NODEC || ACC1) ? 16 : 32;
    static constexpr int kOutLutB = (STG || (!NODEC && BLOCK_N >= 256 && !ACC1)) ? 0 : kOutLutEntries * kOutLutCopies * 4;
    static constexpr int kSmemBytes = kStages * (((kATmem || NODEC) ? 0 : kABytes) + kBBytes) + kCodeStages * kCodeTile + kStageBytes + kLutB + kOutLutB + kParBytes + 1024;
    static_assert(kSmemBytes <= 232448, "shared memory budget");
};

// sixteen relu'd values -> sixteen codes.  fast: the post-ReLU formats (saturating scale, re-based bit pattern, pack).
// exact: the signed quantizer formats through the exact encoder (IEEE quotient, round-half-even) - used after
// quantize_layerout, whose 5-bit values divided by a scale that is itself max / 15.5 of such values land EXACTLY on
// rounding ties of the next grid (e.g. v = vmax / 2 -> 7.75), where the fast formats' ties-up would differ from the
// reference's round-half-even on a visible share of the elements instead of a sliver.
template <bool SFP33>
__device__ __forceinline__ uint4 encode16_fast(const float (&v)[16], float sc) {
    int32_t t[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = encode_relu_fast_raw16<SFP33>(__saturatef(v[i] * sc));
    return make_uint4(ptx::pack_sat_u8x4(t[0], t[1], t[2], t[3]), ptx::pack_sat_u8x4(t[4], t[5], t[6], t[7]),
                      ptx::pack_sat_u8x4(t[8], t[9], t[10], t[11]), ptx::pack_sat_u8x4(t[12], t[13], t[14], t[15]));
}
// sixteen values -> sixteen e4m3 bytes (SLFP_FMT_E4M3): exact quotient, the reference's clamps, round-half-even in the cvt
__device__ __forceinline__ uint4 encode16_e4m3(const float (&v)[16], const DivK& kd, bool relu) {
    uint32_t h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float a = relu ? fmaxf(v[2 * i], 0.0f) : v[2 * i], b = relu ? fmaxf(v[2 * i + 1], 0.0f) : v[2 * i + 1];
        h[i] = encode_e4m3x2(div_k_fused(a, kd), div_k_fused(b, kd));
    }
    return make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
}
// fast forms for the fused epilogues that end in a ReLU: q = v * rk (reciprocal multiply like the other fast formats;
// rk = 1 when the affine vectors were pre-scaled), ReLU folded into the low clamp
__device__ __forceinline__ uint4 encode16_e4m3_relu(const float (&v)[16], float rk) {
    uint32_t h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = encode_e4m3x2_relu(v[2 * i] * rk, v[2 * i + 1] * rk);
    return make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
}
__device__ __forceinline__ uint4 encode16_e4m3_relu_prescaled(const float (&v)[16]) {
    uint32_t h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = encode_e4m3x2_relu(v[2 * i], v[2 * i + 1]);
    return make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
}
template <bool SFP33>
__device__ __forceinline__ uint4 encode16_exact(const float (&v)[16], const DivK& kd) {
    uint32_t c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i)
        c[i] = SFP33 ? encode_relu<SLFP_FMT_SFP33>(div_k_fused(v[i], kd)) : encode_relu<SLFP_FMT_SLFP34_ACT>(div_k_fused(v[i], kd));
    uint32_t pk[4];
#pragma unroll
    for (int g = 0; g < 4; ++g)
        pk[g] = __byte_perm(__byte_perm(c[4 * g], c[4 * g + 1], 0x0040), __byte_perm(c[4 * g + 2], c[4 * g + 3], 0x0040), 0x5410);
    return make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

// sixteen pre-scaled relu inputs u = v / (16 Ka_next) -> the float16 images of their post-ReLU codes (SlfpEpilogue.store_f16):
// the code is formed exactly like encode16_fast (saturating clamp, truncated bit pattern) and looked up in the kernel's
// output table instead of being packed to a byte.  olut_rel = table base + (lane % copies) * 4 - base_code * copies * 4; bit 0
// of it (the address is 4-byte aligned) flags the 16-copy table.
template <bool SFP33, bool PRESAT = false>
__device__ __forceinline__ void encode16_q16(const float (&u)[16], uint32_t olut_rel_f, uint32_t (&h)[8]) {
    const uint32_t half = olut_rel_f & 1u, olut_rel = olut_rel_f & ~1u;                 // warp-uniform
    const uint32_t sh = (SFP33 ? 12u : 11u) + half, mask = half ? 0xffffffc0u : 0xffffff80u;
    const float lo = __uint_as_float(SFP33 ? (0x76Fu << 19) : (0xEDFu << 18));      // raw code 0
    uint32_t e[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint32_t b = __float_as_uint(fmaxf(PRESAT ? u[i] : __saturatef(u[i]), lo));
        e[i] = ptx::lds32_off((b >> sh) & mask, olut_rel);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = ptx::pack16_fma(e[2 * i], e[2 * i + 1]);
}

// ---- epilogue of one 32-row x (BLOCK_N/2)-column slab -----------------------------------------------------
template <int BLOCK_N, int G>
__device__ __forceinline__ void epilogue_slab(const Params& p, int tile, uint32_t tmem_acc, int quad, int half, int lane) {
    const SlfpEpilogue& e = p.epi;
    const int Kout = p.Kout;
    const bool vec4 = (Kout & 3) == 0, vec8 = (Kout & 7) == 0;
    constexpr int kCols = BLOCK_N / G;
    const uint32_t tile_m = mdiv((uint32_t)tile, p.k_ntiles);
    const uint32_t m = tile_m * kBM + (uint32_t)(quad * 32 + lane);
    const int n_base = (int)((uint32_t)tile - tile_m * (uint32_t)p.n_tiles) * BLOCK_N + half * kCols;
    const bool row_ok = m < p.M;
    const bool folded = e.ch_mul != nullptr;
    const bool fast_codes = e.next_fmt == SLFP_FMT_SLFP34_RELU || e.next_fmt == SLFP_FMT_SFP33_RELU;
#pragma unroll 1
    for (int ch = 0; ch < kCols / 16; ++ch) {
        const int n0 = n_base + ch * 16;
        const bool store_f = n0 < Kout;                                      // warp-uniform
        const bool store_c = (e.y_codes != nullptr) && n0 < e.k_phys_out;     // warp-uniform
        if (!store_f && !store_c) continue;
        uint32_t acc[16];
        ptx::tmem_ld16(tmem_acc + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * kCols + ch * 16), acc);
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
        float v[16];
        if (full && folded) {
            // per-channel vectors: warp-uniform 16-byte loads (L1 broadcast), issued before the TMEM wait
            float4 m4[4], a4[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                m4[g] = __ldg(reinterpret_cast<const float4*>(e.ch_mul + n0) + g);
                a4[g] = __ldg(reinterpret_cast<const float4*>(e.ch_add + n0) + g);
            }
            ptx::tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                v[4 * g + 0] = fmaf(__uint_as_float(acc[4 * g + 0]), m4[g].x, a4[g].x);
                v[4 * g + 1] = fmaf(__uint_as_float(acc[4 * g + 1]), m4[g].y, a4[g].y);
                v[4 * g + 2] = fmaf(__uint_as_float(acc[4 * g + 2]), m4[g].z, a4[g].z);
                v[4 * g + 3] = fmaf(__uint_as_float(acc[4 * g + 3]), m4[g].w, a4[g].w);
            }
        } else {
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int n = n0 + i;
                const int nc = n < Kout ? n : Kout - 1;
                float t = __uint_as_float(acc[i]);
                if (folded) {
                    t = fmaf(t, __ldg(e.ch_mul + nc), __ldg(e.ch_add + nc));
                } else {
                    // the reference's order: (acc + bias_q) * Ka * Kw  (conv2d_func.py:24, :46), then the caller's affine
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
        if (e.layerout) {
            // conv -> BN -> layerout_quantize_func (SFP<4,4>) -> ReLU; torch.relu propagates the reference's NaN at exact 0
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                v[i] = layerout_quantize(v[i], e.layerout == 2);
                if (e.relu) v[i] = (v[i] != v[i]) ? v[i] : fmaxf(v[i], 0.0f);
            }
        } else if (e.relu) {
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
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                uint8_t* yc = pass ? e.y_codes2 : e.y_codes;
                if (!yc) continue;
                uint32_t pk[4];
                if (p.e4m3_out) {
                    const uint4 q4 = encode16_e4m3(v, pass ? p.next_div2 : p.next_div, false);     // v is already relu'd when asked
                    pk[0] = q4.x; pk[1] = q4.y; pk[2] = q4.z; pk[3] = q4.w;
                    if (!full) {
#pragma unroll
                        for (int g = 0; g < 4; ++g)
#pragma unroll
                            for (int bte = 0; bte < 4; ++bte)
                                if (n0 + 4 * g + bte >= Kout) pk[g] &= ~(0xffu << (8 * bte));     // zero code in pad channels
                    }
                } else if (fast_codes) {
                    // post-ReLU codes (v >= +0 here: the host requires relu for these formats): re-based
                    // rounded bit pattern, saturating pack.  slfp_common.cuh encode_relu_fast.
                    const float rk = pass ? p.rk2 : p.rk1;
                    int32_t t[16];
                    if (e.next_fmt == SLFP_FMT_SFP33_RELU) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) t[i] = encode_relu_fast_raw<true>(v[i] * rk);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) t[i] = encode_relu_fast_raw<false>(v[i] * rk);
                    }
                    if (!full) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) t[i] = (n0 + i < Kout) ? t[i] : 0;   // zero code in pad channels
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g) pk[g] = ptx::pack_sat_u8x4(t[4 * g], t[4 * g + 1], t[4 * g + 2], t[4 * g + 3]);
                } else {
                    // signed code formats: the exact quantizer (IEEE quotient via the reciprocal sequence)
                    const DivK kd = pass ? p.next_div2 : p.next_div;
                    uint32_t c[16];
                    const bool relu_path = e.relu && kd.k > 0.f && !e.layerout;    // (layerout may hand a NaN through the ReLU)
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
                        for (int i = 0; i < 16; ++i) c[i] = (n0 + i < Kout) ? c[i] : 0u;
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        pk[g] = __byte_perm(__byte_perm(c[4 * g], c[4 * g + 1], 0x0040), __byte_perm(c[4 * g + 2], c[4 * g + 3], 0x0040), 0x5410);
                }
                *reinterpret_cast<uint4*>(yc + (size_t)m * e.k_phys_out + n0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
    }
}

// ---- fast epilogues (the fused pipeline's common cases; chosen on the host, Params::epi_mode) -----------------
// Requirements: folded per-channel affine, ReLU, Kout % 16 == 0, no float32 output, post-ReLU code format.
//   MODE 1: one code tensor, nothing else.  The per-channel vectors staged in shared memory are pre-scaled by
//           1/(16 Ka_next), so ONE saturating FMA yields clamp(relu(y)/Ka_next/16, 0, 1); 2 integer ops + half a
//           pack instruction turn it into the code (slfp_common.cuh encode_relu_fast_raw16).
//   MODE 2: optional float16 residual, optional float16 output of relu(y), one or two code tensors.
//
// Stores.  A thread owns one output pixel (TMEM lane = row), so a plain store instruction would write 16 bytes
// of 32 different rows: half-filled 32-byte sectors, measured at ~40 % of the write bandwidth of full sectors
// (tools/ubench/store_pattern.cu).  Lane pairs therefore swap one 16-byte piece (4 SHFL) so that every store
// instruction writes 32 contiguous bytes per row: lanes 2i / 2i+1 hold pieces (A, B) = 32 contiguous bytes of
// rows 2i / 2i+1; afterwards `first` belongs to row 2i and `second` to row 2i+1, both at piece index lane & 1.
__device__ __forceinline__ void pair_exchange(const uint4& A, const uint4& B, bool odd, uint4& first, uint4& second) {
    uint4 send, recv;
    send.x = odd ? A.x : B.x; send.y = odd ? A.y : B.y; send.z = odd ? A.z : B.z; send.w = odd ? A.w : B.w;
    recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
    recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
    recv.z = __shfl_xor_sync(0xffffffffu, send.z, 1);
    recv.w = __shfl_xor_sync(0xffffffffu, send.w, 1);
    first.x = odd ? recv.x : A.x; first.y = odd ? recv.y : A.y; first.z = odd ? recv.z : A.z; first.w = odd ? recv.w : A.w;
    second.x = odd ? B.x : recv.x; second.y = odd ? B.y : recv.y; second.z = odd ? B.z : recv.z; second.w = odd ? B.w : recv.w;
}

// OUT (MODE 1): what the single output tensor receives - 0 post-ReLU code bytes, 1 e4m3 bytes, 2 float16 images of the codes
// (store_f16); -1 = decided at run time (MODE 2).  A compile-time OUT lets the clamp fold into the affine (FFMA.SAT) and drops
// the per-chunk flag tests: the 1x1 reduce layers are bound by the instruction issue of their epilogue warps.
template <int BLOCK_N, int G, int MODE, bool SFP33, int OUT = -1>
__device__ __forceinline__ void epilogue_fast(const Params& p, int tile, int next_tile, uint32_t tmem_acc, int quad, int half,
                                              int lane, uint32_t s_mul, uint32_t s_add, uint32_t olut_rel) {
    static_assert(BLOCK_N / G >= 32, "32-column steps");
    constexpr int kCols = BLOCK_N / G;
    const int Kout = p.Kout;
    const bool kE4m3 = OUT < 0 ? (p.e4m3_out != 0) : OUT == 1, kQ16 = OUT < 0 ? (p.epi.store_f16 != 0) : OUT == 2;
    const uint32_t tile_m = mdiv((uint32_t)tile, p.k_ntiles);
    const uint32_t m = tile_m * kBM + (uint32_t)(quad * 32 + lane);
    const int n_tile0 = (int)((uint32_t)tile - tile_m * (uint32_t)p.n_tiles) * BLOCK_N;
    const bool odd = (lane & 1) != 0;
    // rows this lane STORES to after the pair exchange (m's parity is the lane's)
    const uint32_t m_first = m & ~1u, m_second = m | 1u;
    const bool ok_first = m_first < p.M, ok_second = m_second < p.M, row_ok = m < p.M;
    const size_t row = (size_t)m * (size_t)Kout;
    const size_t row_first = (size_t)m_first * (size_t)Kout, row_second = (size_t)m_second * (size_t)Kout;
    const __half* resp = MODE == 2 ? reinterpret_cast<const __half*>(p.epi.residual) : nullptr;
    __half* y16 = MODE == 2 ? reinterpret_cast<__half*>(p.epi.y_f16) : nullptr;
    uint8_t* yc1 = p.epi.y_codes;
    uint8_t* yc2 = MODE == 2 ? p.epi.y_codes2 : nullptr;
    const float sc1 = p.sc1, sc2 = p.sc2;
    if (MODE == 2 && resp != nullptr && next_tile < p.num_tiles) {
        // push this lane's row segment of the NEXT tile's residual towards L2 (no register, no scoreboard)
        const uint32_t next_m = mdiv((uint32_t)next_tile, p.k_ntiles);
        const uint32_t mn = next_m * kBM + (uint32_t)(quad * 32 + lane);
        const int nb = (int)((uint32_t)next_tile - next_m * (uint32_t)p.n_tiles) * BLOCK_N + half * kCols;
        if (mn < p.M) {
#pragma unroll
            for (int c = 0; c < kCols; c += 64)
                if (nb + c < Kout) ptx::prefetch_l2(resp + (size_t)mn * Kout + nb + c);
        }
    }
#pragma unroll 1
    for (int ch = 0; ch < kCols / 32; ++ch) {
        const int col = half * kCols + ch * 32;
        const int n0 = n_tile0 + col;
        if (n0 >= Kout) break;                                   // warp-uniform
        const bool two = n0 + 32 <= Kout;                        // else 16 valid columns (Kout % 16 == 0)
        uint32_t acc[32];
        ptx::tmem_ld32(tmem_acc + ((uint32_t)(quad * 32) << 16) + (uint32_t)col, acc);
        uint4 rr[4];
        if (MODE == 2 && resp != nullptr) {
            rr[0] = rr[1] = rr[2] = rr[3] = make_uint4(0u, 0u, 0u, 0u);
            if (row_ok) {
                const uint4* rp = reinterpret_cast<const uint4*>(resp + row + n0);
                rr[0] = __ldg(rp); rr[1] = __ldg(rp + 1);
                if (two) { rr[2] = __ldg(rp + 2); rr[3] = __ldg(rp + 3); }
            }
        }
        ptx::tmem_ld_wait();
        uint4 pk1[2], pk2[2];                                    // code pieces of the two 16-column halves
#pragma unroll
        for (int o = 0; o < 32; o += 16) {
            if (o == 16 && !two) break;
            float v[16];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const float4 m4 = ptx::lds128_f4(s_mul + (uint32_t)(col + o + 4 * g) * 4u);
                const float4 a4 = ptx::lds128_f4(s_add + (uint32_t)(col + o + 4 * g) * 4u);
                v[4 * g + 0] = fmaf(__uint_as_float(acc[o + 4 * g + 0]), m4.x, a4.x);
                v[4 * g + 1] = fmaf(__uint_as_float(acc[o + 4 * g + 1]), m4.y, a4.y);
                v[4 * g + 2] = fmaf(__uint_as_float(acc[o + 4 * g + 2]), m4.z, a4.z);
                v[4 * g + 3] = fmaf(__uint_as_float(acc[o + 4 * g + 3]), m4.w, a4.w);
                if (MODE == 1 && (OUT == 0 || OUT == 2)) {           // the clamp to [0, 1] folds into the FMA
                    v[4 * g + 0] = __saturatef(v[4 * g + 0]); v[4 * g + 1] = __saturatef(v[4 * g + 1]);
                    v[4 * g + 2] = __saturatef(v[4 * g + 2]); v[4 * g + 3] = __saturatef(v[4 * g + 3]);
                }
            }
            if (MODE == 1 && kE4m3) {
                pk1[o / 16] = encode16_e4m3_relu_prescaled(v);       // the staged affine carries 1 / Ka_next
            } else if (MODE == 1 && kQ16) {
                // float16 images of the codes: 32 contiguous bytes (one sector) of this lane's row
                uint32_t hq[8];
                encode16_q16<SFP33, OUT == 2>(v, olut_rel, hq);
                // lane pairs swap 16-byte pieces: every store instruction writes whole 32-byte sectors
                uint4 f1, f2;
                pair_exchange(make_uint4(hq[0], hq[1], hq[2], hq[3]), make_uint4(hq[4], hq[5], hq[6], hq[7]), odd, f1, f2);
                __half* yh = reinterpret_cast<__half*>(yc1);
                if (ok_first) *reinterpret_cast<uint4*>(yh + row_first + n0 + o + (odd ? 8 : 0)) = f1;
                if (ok_second) *reinterpret_cast<uint4*>(yh + row_second + n0 + o + (odd ? 8 : 0)) = f2;
            } else if (MODE == 1) {
                int32_t t[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) t[i] = encode_relu_fast_raw16<SFP33>(OUT == 0 ? v[i] : __saturatef(v[i]));
                pk1[o / 16] = make_uint4(ptx::pack_sat_u8x4(t[0], t[1], t[2], t[3]), ptx::pack_sat_u8x4(t[4], t[5], t[6], t[7]),
                                         ptx::pack_sat_u8x4(t[8], t[9], t[10], t[11]), ptx::pack_sat_u8x4(t[12], t[13], t[14], t[15]));
            } else {
                if (resp != nullptr) {
                    const uint4 ra = rr[o / 8], rb = rr[o / 8 + 1];
                    const uint32_t rw[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&rw[i]));
                        v[2 * i] += f.x; v[2 * i + 1] += f.y;
                    }
                }
                if (p.epi.layerout) {                            // conv -> BN -> quantize_layerout -> ReLU
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = layerout_relu(v[i]);
                }
                if (y16 != nullptr) {
                    uint32_t hw[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const __half2 h = __floats2half2_rn(fmaxf(v[2 * i], 0.0f), fmaxf(v[2 * i + 1], 0.0f));
                        hw[i] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    uint4 f1, f2;
                    pair_exchange(make_uint4(hw[0], hw[1], hw[2], hw[3]), make_uint4(hw[4], hw[5], hw[6], hw[7]), odd, f1, f2);
                    if (ok_first) *reinterpret_cast<uint4*>(y16 + row_first + n0 + o + (odd ? 8 : 0)) = f1;
                    if (ok_second) *reinterpret_cast<uint4*>(y16 + row_second + n0 + o + (odd ? 8 : 0)) = f2;
                }
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    if ((pass ? yc2 : yc1) == nullptr) continue;
                    const uint4 pk = p.e4m3_out ? encode16_e4m3_relu(v, pass ? p.rk2 : p.rk1)
                                     : p.epi.layerout ? encode16_exact<SFP33>(v, pass ? p.next_div2 : p.next_div)
                                                      : encode16_fast<SFP33>(v, pass ? sc2 : sc1);
                    if (pass) pk2[o / 16] = pk; else pk1[o / 16] = pk;
                }
            }
        }
        if (MODE == 1 && kQ16) continue;                        // stored above
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            uint8_t* yc = pass ? yc2 : yc1;
            if (MODE == 1 ? pass == 1 : yc == nullptr) continue;
            const uint4 pa = pass ? pk2[0] : pk1[0];
            if (two) {
                uint4 f1, f2;
                pair_exchange(pa, pass ? pk2[1] : pk1[1], odd, f1, f2);
                if (ok_first) *reinterpret_cast<uint4*>(yc + row_first + n0 + (odd ? 16 : 0)) = f1;
                if (ok_second) *reinterpret_cast<uint4*>(yc + row_second + n0 + (odd ? 16 : 0)) = f2;
            } else if (row_ok) {
                *reinterpret_cast<uint4*>(yc + row + n0) = pa;
            }
        }
    }
}

// The same epilogues in 16-column steps for the 16-epilogue-warp role split (88 registers per thread, 4 column
// groups of BLOCK_N / 4 columns): the residual of the next chunk is requested before the current one is
// processed, code pieces of two consecutive chunks pair up for the full-sector stores.
template <int BLOCK_N, int G, int MODE, bool SFP33, int OUT = -1>
__device__ __forceinline__ void epilogue_fast16(const Params& p, int tile, int next_tile, uint32_t tmem_acc, int quad, int cg,
                                                int lane, uint32_t s_mul, uint32_t s_add, uint32_t olut_rel) {
    constexpr int kCols = BLOCK_N / G;
    constexpr int kChunks = kCols / 16;
    const int Kout = p.Kout;
    const bool kE4m3 = OUT < 0 ? (p.e4m3_out != 0) : OUT == 1, kQ16 = OUT < 0 ? (p.epi.store_f16 != 0) : OUT == 2;
    const uint32_t tile_m = mdiv((uint32_t)tile, p.k_ntiles);
    const uint32_t m = tile_m * kBM + (uint32_t)(quad * 32 + lane);
    const int n_base = (int)((uint32_t)tile - tile_m * (uint32_t)p.n_tiles) * BLOCK_N + cg * kCols;
    const bool odd = (lane & 1) != 0;
    const uint32_t m_first = m & ~1u, m_second = m | 1u;
    const bool ok_first = m_first < p.M, ok_second = m_second < p.M, row_ok = m < p.M;
    const size_t row = (size_t)m * (size_t)Kout + n_base;
    const size_t row_first = (size_t)m_first * (size_t)Kout + n_base, row_second = (size_t)m_second * (size_t)Kout + n_base;
    const __half* resp = MODE == 2 ? reinterpret_cast<const __half*>(p.epi.residual) : nullptr;
    __half* y16 = MODE == 2 ? reinterpret_cast<__half*>(p.epi.y_f16) : nullptr;
    uint8_t* yc1 = p.epi.y_codes;
    uint8_t* yc2 = MODE == 2 ? p.epi.y_codes2 : nullptr;
    const float sc1 = p.sc1, sc2 = p.sc2;
    int nvalid = (Kout - n_base) >> 4;                           // chunks of this slab inside Kout
    nvalid = nvalid > kChunks ? kChunks : nvalid;
    if (MODE == 2 && resp != nullptr && next_tile < p.num_tiles) {
        const uint32_t next_m = mdiv((uint32_t)next_tile, p.k_ntiles);
        const uint32_t mn = next_m * kBM + (uint32_t)(quad * 32 + lane);
        const int nb = (int)((uint32_t)next_tile - next_m * (uint32_t)p.n_tiles) * BLOCK_N + cg * kCols;
        if (mn < p.M && nb < Kout) ptx::prefetch_l2(resp + (size_t)mn * Kout + nb);     // 32..128 bytes of one row
    }
    uint4 ra = make_uint4(0u, 0u, 0u, 0u), rb = ra, na = ra, nb4 = ra;
    if (MODE == 2 && resp != nullptr && nvalid > 0 && row_ok) {
        ra = __ldg(reinterpret_cast<const uint4*>(resp + row));
        rb = __ldg(reinterpret_cast<const uint4*>(resp + row) + 1);
    }
    uint4 prev1 = ra, prev2 = ra;
#pragma unroll
    for (int ch = 0; ch < kChunks; ++ch) {
        if (ch >= nvalid) break;                                 // warp-uniform
        uint32_t acc[16];
        ptx::tmem_ld16(tmem_acc + ((uint32_t)(quad * 32) << 16) + (uint32_t)(cg * kCols + ch * 16), acc);
        if (MODE == 2 && resp != nullptr && ch + 1 < kChunks) {
            na = nb4 = make_uint4(0u, 0u, 0u, 0u);
            if (ch + 1 < nvalid && row_ok) {
                na = __ldg(reinterpret_cast<const uint4*>(resp + row + (ch + 1) * 16));
                nb4 = __ldg(reinterpret_cast<const uint4*>(resp + row + (ch + 1) * 16) + 1);
            }
        }
        ptx::tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const float4 m4 = ptx::lds128_f4(s_mul + (uint32_t)(cg * kCols + ch * 16 + 4 * g) * 4u);
            const float4 a4 = ptx::lds128_f4(s_add + (uint32_t)(cg * kCols + ch * 16 + 4 * g) * 4u);
            v[4 * g + 0] = fmaf(__uint_as_float(acc[4 * g + 0]), m4.x, a4.x);
            v[4 * g + 1] = fmaf(__uint_as_float(acc[4 * g + 1]), m4.y, a4.y);
            v[4 * g + 2] = fmaf(__uint_as_float(acc[4 * g + 2]), m4.z, a4.z);
            v[4 * g + 3] = fmaf(__uint_as_float(acc[4 * g + 3]), m4.w, a4.w);
            if (MODE == 1 && (OUT == 0 || OUT == 2)) {               // the clamp to [0, 1] folds into the FMA
                v[4 * g + 0] = __saturatef(v[4 * g + 0]); v[4 * g + 1] = __saturatef(v[4 * g + 1]);
                v[4 * g + 2] = __saturatef(v[4 * g + 2]); v[4 * g + 3] = __saturatef(v[4 * g + 3]);
            }
        }
        uint4 pk1, pk2 = make_uint4(0u, 0u, 0u, 0u);
        if (MODE == 1 && kE4m3) {
            pk1 = encode16_e4m3_relu_prescaled(v);                   // the staged affine carries 1 / Ka_next
        } else if (MODE == 1 && kQ16) {
            uint32_t hq[8];
            encode16_q16<SFP33, OUT == 2>(v, olut_rel, hq);
            uint4 f1, f2;
            pair_exchange(make_uint4(hq[0], hq[1], hq[2], hq[3]), make_uint4(hq[4], hq[5], hq[6], hq[7]), odd, f1, f2);
            __half* yh = reinterpret_cast<__half*>(yc1);
            if (ok_first) *reinterpret_cast<uint4*>(yh + row_first + ch * 16 + (odd ? 8 : 0)) = f1;
            if (ok_second) *reinterpret_cast<uint4*>(yh + row_second + ch * 16 + (odd ? 8 : 0)) = f2;
            continue;
        } else if (MODE == 1) {
            int32_t t[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) t[i] = encode_relu_fast_raw16<SFP33>(OUT == 0 ? v[i] : __saturatef(v[i]));
            pk1 = make_uint4(ptx::pack_sat_u8x4(t[0], t[1], t[2], t[3]), ptx::pack_sat_u8x4(t[4], t[5], t[6], t[7]),
                             ptx::pack_sat_u8x4(t[8], t[9], t[10], t[11]), ptx::pack_sat_u8x4(t[12], t[13], t[14], t[15]));
        } else {
            if (resp != nullptr) {
                const uint32_t rw[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&rw[i]));
                    v[2 * i] += f.x; v[2 * i + 1] += f.y;
                }
                ra = na; rb = nb4;
            }
            if (p.epi.layerout) {                                // conv -> BN -> quantize_layerout -> ReLU
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = layerout_relu(v[i]);
            }
            if (y16 != nullptr) {
                uint32_t hw[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const __half2 h = __floats2half2_rn(fmaxf(v[2 * i], 0.0f), fmaxf(v[2 * i + 1], 0.0f));
                    hw[i] = *reinterpret_cast<const uint32_t*>(&h);
                }
                uint4 f1, f2;
                pair_exchange(make_uint4(hw[0], hw[1], hw[2], hw[3]), make_uint4(hw[4], hw[5], hw[6], hw[7]), odd, f1, f2);
                if (ok_first) *reinterpret_cast<uint4*>(y16 + row_first + ch * 16 + (odd ? 8 : 0)) = f1;
                if (ok_second) *reinterpret_cast<uint4*>(y16 + row_second + ch * 16 + (odd ? 8 : 0)) = f2;
            }
            pk1 = pk2;
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                if ((pass ? yc2 : yc1) == nullptr) continue;
                const uint4 pk = p.e4m3_out ? encode16_e4m3_relu(v, pass ? p.rk2 : p.rk1)
                                 : p.epi.layerout ? encode16_exact<SFP33>(v, pass ? p.next_div2 : p.next_div)
                                                  : encode16_fast<SFP33>(v, pass ? sc2 : sc1);
                if (pass) pk2 = pk; else pk1 = pk;
            }
        }
        // code stores: odd chunk -> pair with the previous chunk's piece (32 contiguous bytes per row and lane pair);
        // an even chunk without a successor goes out alone (16 bytes)
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            uint8_t* yc = pass ? yc2 : yc1;
            if (MODE == 1 ? pass == 1 : yc == nullptr) continue;
            const uint4 cur = pass ? pk2 : pk1;
            if (ch & 1) {
                uint4 f1, f2;
                pair_exchange(pass ? prev2 : prev1, cur, odd, f1, f2);
                if (ok_first) *reinterpret_cast<uint4*>(yc + row_first + (ch - 1) * 16 + (odd ? 16 : 0)) = f1;
                if (ok_second) *reinterpret_cast<uint4*>(yc + row_second + (ch - 1) * 16 + (odd ? 16 : 0)) = f2;
            } else if (ch + 1 >= nvalid) {
                if (row_ok) *reinterpret_cast<uint4*>(yc + row + ch * 16) = cur;
            } else {
                if (pass) prev2 = cur; else prev1 = cur;
            }
        }
    }
}

// The four output-side tensor maps (residual, float16 output, two code outputs; [M, Kout] matrices, boxes of
// 128 rows x 128 bytes) are only dereferenced by the STG instantiation.
struct OutMaps {
    CUtensorMap res, y16, c1, c2;
};

// HIFI (SLFP_CONV_SPLIT_OPERANDS): three passes over K into the same accumulator - pass 0: x_hi * w_hi, pass 1: x_hi * w_lo,
// pass 2: x_lo * w_hi - with (hi, lo) the float16 pair of a value.  The decode table holds hi in the low and lo in the
// high half of an entry, the decode warps pick the half per K block; the weight rows are [hi | lo].
template <int BLOCK_N, int GRAN, int DW, bool STG, bool HIFI = false, bool NODEC = false, bool A16 = false, bool ACC1 = false>
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_v2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                     const __grid_constant__ CUtensorMap tmap_x2, const __grid_constant__ OutMaps omaps, const Params p) {
    using C = Cfg<BLOCK_N, STG, NODEC, A16, ACC1>;
    constexpr uint32_t kAccBufs = (uint32_t)C::kAccBufs;
    static_assert(!NODEC || (GRAN == 64 && !HIFI), "e4m3 / float16 operands: whole 64-channel K blocks");
    static_assert(!STG || (BLOCK_N == 128 && DW == 8 && GRAN == 64), "staged epilogue: 128-column tiles, 16 epilogue warps");
    static_assert(!HIFI || !STG, "the split-operand mode uses the generic epilogue");
    using R = Roles<DW>;
    constexpr int kCodeStages = C::kCodeStages;
    constexpr int kDecWarps = R::kDecWarps, kEpiWarps = R::kEpiWarps, kEpiWarp0 = R::kEpiWarp0, kGroups = R::kGroups;
    constexpr int kDecGroupWarps = (kDecWarps / 4 <= C::kStages) ? 4 : 8;     // warps that decode one K block together
    // SW128 operand tiles need 1024-byte alignment.  The kernel has no static shared memory, so the dynamic
    // window starts at a link-time constant that honours __align__ (checked below): every shared address in
    // this kernel is then a compile-time offset and the table look-up needs no base-address add.
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_a = smem;                                   // [stages][128 rows][128 B]   (absent when A lives in TMEM)
    uint8_t* s_b = s_a + ((C::kATmem || NODEC) ? 0 : C::kStages * kABytes);   // [stages][BLOCK_N rows][128 B] (NODEC: 64 B)
    uint8_t* s_code = s_b + C::kStages * C::kBBytes;       // [kCodeStages][128 pixels][64 B]   (A16: 128 B rows)
    uint8_t* s_stage = s_code + kCodeStages * C::kCodeTile;  // STG: [2][4 groups][128][64 B] float16 + [2][4][128][32 B] codes
    uint32_t* s_lut = reinterpret_cast<uint32_t*>(s_stage + C::kStageBytes);
    uint32_t* s_olut = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(s_lut) + C::kLutB);   // [258][32] (store_f16)
    float* s_par = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s_olut) + C::kOutLutB);   // [2][BLOCK_N]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_par) + C::kParBytes);
    uint64_t* bar_cfull = s_bar;                           // [kCodeStages]  1 arrive.expect_tx
    uint64_t* bar_cempty = bar_cfull + kCodeStages;        // [kCodeStages]  8 decode warps
    uint64_t* bar_full = bar_cempty + kCodeStages;         // [stages]  8 decode warps + 1 weight-TMA arrive
    uint64_t* bar_empty = bar_full + C::kStages;           // [stages]  1 tcgen05.commit
    uint64_t* bar_tfull = bar_empty + C::kStages;          // [2]       1 tcgen05.commit
    uint64_t* bar_tempty = bar_tfull + 2;                  // [2]       8 epilogue warps
    uint64_t* bar_res = bar_tempty + 2;                    // [2]  STG: residual tile landed (1 arrive.expect_tx); 8 words reserved
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_res + 8);

    // warp index through a shuffle: tells the compiler it is warp-uniform, so role branches, barrier addresses and
    // the tcgen05 operands stay in uniform registers (no per-lane waterfall loops around UTCHMMA / UTCBAR)
    // Role index `warp` = physical warp rotated by four: the control roles (0..3: TMA producers, MMA issuer) run on the
    // PHYSICAL warps 24..27.  A scheduler picks the highest warp id among its eligible warps, so as warps 0..3 the
    // single-thread roles only got the issue slots the decode / epilogue warps left over (the producers needed
    // ~1 400 cycles per K block for ~50 instructions, tools/role_profile.py).  The rotation keeps warp % 4, i.e. the
    // TMEM lane quadrant a warp may touch, and warpgroup alignment for setmaxnreg.
    const int tid = threadIdx.x, pwarp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int warp = pwarp >= kWorkWarps ? pwarp - kWorkWarps : pwarp + kCtrlWarps;
    if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();     // would break the swizzle: fail loudly, never silently

    // ---- one-time setup ---------------------------------------------------------------------------
    // Decode tables: every code is decoded ONCE (a lane per code) and broadcast into its lane copies by shuffles - filling
    // all 8 192 words with their own decode_act_any (a divergent constant-table walk) was 1-2 us of every launch's prologue.
    const int pw0 = tid >> 5;
    if (!NODEC && pw0 < 8) {
        const float f = decode_act_any((uint32_t)(pw0 * 32 + lane), p.act_fmt, c_pow2frac);
        const __half hi = __float2half_rn(f);
        uint32_t e = (uint32_t)__half_as_ushort(hi);
        if (HIFI) e |= (uint32_t)__half_as_ushort(__float2half_rn(f - __half2float(hi))) << 16;
#pragma unroll 8
        for (int j = 0; j < 32; ++j) s_lut[(pw0 * 32 + j) * 32 + lane] = __shfl_sync(0xffffffffu, e, j);
    }
    if (C::kOutLutB != 0 && p.epi.store_f16 && pw0 >= 8 && pw0 < 8 + (kOutLutEntries + 31) / 32) {
        // raw post-ReLU codes 0..257 (the byte path saturates at 255) -> float16 of the value the consumer's table would decode
        const uint32_t c0 = (uint32_t)(pw0 - 8) * 32u, c = c0 + (uint32_t)lane;
        const uint32_t e = (uint32_t)__half_as_ushort(__float2half_rn(decode_act_any(c > 255u ? 255u : c, p.epi.next_fmt, c_pow2frac)));
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
            const uint32_t ej = __shfl_sync(0xffffffffu, e, j);
            if (c0 + (uint32_t)j < (uint32_t)kOutLutEntries && lane < C::kOutLutCopies) s_olut[(c0 + j) * C::kOutLutCopies + lane] = ej;
        }
    }
    if (warp == kWarpCode && lane == 0) {
        ptx::prefetch_tmap(&tmap_x);
        ptx::prefetch_tmap(&tmap_w);
        ptx::prefetch_tmap(&tmap_x2);
        for (int s = 0; s < kCodeStages; ++s) {
            ptx::mbar_init(ptx::smem_u32(&bar_cfull[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&bar_cempty[s]), kDecGroupWarps);   // the warps of one decode group
        }
        for (int s = 0; s < C::kStages; ++s) {
            ptx::mbar_init(ptx::smem_u32(&bar_full[s]), NODEC ? 1 : kDecGroupWarps + 1);  // decode group + the weight TMA's arrive (NODEC: one producer)
            ptx::mbar_init(ptx::smem_u32(&bar_empty[s]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(ptx::smem_u32(&bar_tfull[b]), 1);
            ptx::mbar_init(ptx::smem_u32(&bar_tempty[b]), (STG && p.stg_groups == 2) ? kEpiWarps / 2 : kEpiWarps);
        }
        for (int b = 0; b < 8; ++b) ptx::mbar_init(ptx::smem_u32(&bar_res[b]), 1);
        ptx::fence_mbar_init();
    }
    if (warp == kWarpMma) ptx::tmem_alloc<C::kTmemCols>(ptx::smem_u32(s_tmem));
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    // Programmatic dependent launch: everything above (decode table, barriers, tensor-memory allocation, tensor-map
    // prefetch) touches no global memory written by an earlier kernel, so with the launch attribute set (launch()) it
    // runs while the PREVIOUS kernel of the stream drains its last tiles; from here on we read its output.  Our own
    // dependents may be scheduled as soon as SMs free up (they block at the same point until this grid has completed
    // and flushed).  Without the launch attribute both instructions are no-ops.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const int my_tiles = ((int)blockIdx.x < p.num_tiles)
                             ? (p.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    // register re-balance (warpgroup granularity; first statement of every role so that ptxas allocates each
    // region against its own budget): producers / MMA / decode need few registers, the epilogue many
    if (warp < kDecWarp0) {
      ptx::setmaxnreg_dec<kRegsCtrl>();
      if (warp == kWarpCode || warp == kWarpCode2) {
        // =========================== code producers: TMA im2col ==========================================
        // TWO producer threads (warps 0 and 3), alternating K blocks.  One thread was the slowest stage of the
        // stage-1 layers and the stem: its ~50 dependent instructions per K block (barrier probe, expect_tx, address
        // arithmetic, up to four TMA issues) share a scheduler with six busy warps and took ~800 cycles per K block,
        // while the decode warps idled on the `full` barrier (tools/role_profile.py: 58-78 % of their life).
        // The whole warp runs the loop converged and one elected lane issues (like the MMA issuer): every operand of
        // the TMA / mbarrier instructions then lives in uniform registers - the per-lane form cost ~45 instructions
        // per K block in R2UR moves and an ELECT waterfall loop around each UTMALDG.
        {
            PROF_VARS;
            static_assert(kCodeStages % 2 == 0, "the two producers alternate K blocks: a stage must always be filled by the same one");
            const uint32_t mine = warp == kWarpCode ? 0u : 1u;
            uint32_t cs = 0, cphase = 0, g = 0;
            // no-decode variants, resident weights: with one N tile and a K-block count that divides the ring depth, stage s
            // always holds the weight tile of K block s % num_kb - after the first trip round the ring only the activation
            // tile is loaded (the width-folded stem re-read its 32 KB filter 170 times per CTA: a third of its L2 traffic)
            const bool w_res_nd = NODEC && p.n_tiles == 1 && p.num_kb <= kCodeStages && (kCodeStages % p.num_kb) == 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
                const uint32_t tile_m = mdiv((uint32_t)tile, p.k_ntiles);
                const int tile_n = (int)((uint32_t)tile - tile_m * (uint32_t)p.n_tiles);
                const uint32_t m0 = tile_m * kBM;
                const int n = (int)mdiv(m0, p.k_howo);
                const int rem = (int)(m0 - (uint32_t)n * (uint32_t)p.HoWo);
                const int ho = (int)mdiv((uint32_t)rem, p.k_wo), wo = rem - ho * p.Wo;
                const int w0 = wo * p.sw - p.pw, h0 = ho * p.sh - p.ph;
                int tap = 0, r = 0, s = 0, cb = 0;       // GRAN 64: cb = 64-channel block; GRAN 16: 16-channel block
                int kb = 0;                               // K block inside ONE pass over K (the split-operand mode makes three)
                for (int kbt = 0; kbt < p.num_kb; ++kbt, ++kb, ++g) {
                    if (HIFI && kb == p.kb_base) { kb = 0; tap = 0; r = 0; s = 0; cb = 0; }
                    if ((g & 1u) != mine) {                 // the other producer's K block: only advance the counters
                        if (GRAN == 64) {
                            if (++cb == p.cblocks) { cb = 0; ++tap; if (++s == p.S) { s = 0; ++r; } }
                        } else {
                            int valid = p.taps * p.c16s - kb * 4;
                            valid = valid > 4 ? 4 : valid;
                            for (int j = 0; j < valid; ++j)
                                if (++cb == p.c16s) { cb = 0; ++tap; if (++s == p.S) { s = 0; ++r; } }
                        }
                        if (++cs == (uint32_t)kCodeStages) { cs = 0; cphase ^= 1u; }
                        continue;
                    }
                    PROF(a, ptx::mbar_wait(ptx::smem_u32(NODEC ? &bar_empty[cs] : &bar_cempty[cs]), cphase ^ 1u, 1u | ((uint32_t)kb << 8) | ((uint32_t)ti << 16)));
                    const uint32_t full = ptx::smem_u32(NODEC ? &bar_full[cs] : &bar_cfull[cs]);
                    const uint32_t dst = ptx::smem_u32(s_code + cs * C::kCodeTile);
                    if (GRAN == 64) {
                        const bool load_b = NODEC && !(w_res_nd && g >= (uint32_t)kCodeStages);
                        if (ptx::elect_one()) {
                            ptx::mbar_arrive_expect_tx(full, (uint32_t)C::kCodeTile + (load_b ? (uint32_t)C::kBBytes : 0u));
                            if (load_b)                 // the weight tile of this K block rides on the same barrier
                                ptx::tma_load_2d(ptx::smem_u32(s_b + cs * C::kBBytes), &tmap_w, full, kbt * kBK, tile_n * BLOCK_N);
                            if (kb >= p.nkb1)           // concatenated second input: 1x1 window at (ho*sh2, wo*sw2)
                                ptx::tma_load_im2col_4d(dst, &tmap_x2, full, (kb - p.nkb1) * 64, wo * p.sw2, ho * p.sh2, n, 0, 0);
                            else if (p.plain_1x1) ptx::tma_load_2d(dst, &tmap_x, full, cb * 64, (int)m0);
                            else ptx::tma_load_im2col_4d(dst, &tmap_x, full, cb * 64, w0, h0, n, (uint16_t)(s * p.dw), (uint16_t)(r * p.dh));
                        }
                        __syncwarp();
                        if (++cb == p.cblocks) { cb = 0; ++tap; if (++s == p.S) { s = 0; ++r; } }
                    } else {
                        int valid = p.taps * p.c16s - kb * 4;          // 16-channel pieces left in K
                        valid = valid > 4 ? 4 : valid;
                        if (ptx::elect_one()) ptx::mbar_arrive_expect_tx(full, (uint32_t)(valid * (kCodeBytes / 4)));
                        __syncwarp();
                        for (int j = 0; j < valid; ++j) {
                            if (ptx::elect_one())
                                ptx::tma_load_im2col_4d(dst + (uint32_t)(j * (kCodeBytes / 4)), &tmap_x, full, cb * 16, w0, h0, n,
                                                        (uint16_t)(s * p.dw), (uint16_t)(r * p.dh));
                            __syncwarp();
                            if (++cb == p.c16s) { cb = 0; ++tap; if (++s == p.S) { s = 0; ++r; } }
                        }
                    }
                    if (++cs == (uint32_t)kCodeStages) { cs = 0; cphase ^= 1u; }
                }
            }
            if (mine == 0u && lane == 0) { PROF_FLUSH(8); }
        }
        __syncwarp();
    } else if (warp == kWarpWgt) {
        // =========================== weight producer: TMA tiles (B operand) ================================
        {
            PROF_VARS;
            uint32_t stage = 0, phase = 0;
            // Resident weights: with a single N tile and a K-block count that divides the ring depth, stage s always holds
            // K block s % num_kb of the same filter - after the first pass over the ring the producer only arrives on the
            // barrier.  Every TMA load costs >= ~190 cycles of the TMA unit whatever its size (tools/ubench/tma_box.cu), and
            // the space-to-depth stem, at 4 code loads + 1 weight load per K block, ran exactly at that floor.
            const bool w_res = !HIFI && p.n_tiles == 1 && p.num_kb <= C::kStages && (C::kStages % p.num_kb) == 0;
            uint32_t issued = 0;
            for (int ti = 0; ti < (NODEC ? 0 : my_tiles); ++ti) {
                const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
                const int n0 = (int)((uint32_t)tile - mdiv((uint32_t)tile, p.k_ntiles) * (uint32_t)p.n_tiles) * BLOCK_N;
                for (int kb = 0; kb < p.num_kb; ++kb, ++issued) {
                    PROF(a, ptx::mbar_wait(ptx::smem_u32(&bar_empty[stage]), phase ^ 1u, 2u | ((uint32_t)kb << 8) | ((uint32_t)ti << 16)));
                    const uint32_t full = ptx::smem_u32(&bar_full[stage]);
                    if (w_res && issued >= (uint32_t)C::kStages) {
                        if (ptx::elect_one()) ptx::mbar_arrive(full);
                    } else if (ptx::elect_one()) {
                        ptx::mbar_arrive_expect_tx(full, (uint32_t)C::kBBytes);
                        int wcol = kb * kBK;
                        if (HIFI) {                     // pass 0 / 2: w_hi, pass 1: w_lo (second half of the weight row)
                            const int pass = kb / p.kb_base, kbb = kb - pass * p.kb_base;
                            wcol = kbb * kBK + (pass == 1 ? p.w_lo_col : 0);
                        }
                        ptx::tma_load_2d(ptx::smem_u32(s_b + stage * C::kBBytes), &tmap_w, full, wcol, n0);
                    }
                    __syncwarp();
                    if (++stage == (uint32_t)C::kStages) { stage = 0; phase ^= 1u; }
                }
            }
            if (lane == 0) { PROF_FLUSH(12); }
        }
        __syncwarp();
    } else if (warp == kWarpMma) {
        // =========================== MMA issuer ===================================================
        // The whole warp runs the loop converged and ONE elected lane issues: with 7 warps per scheduler every
        // instruction of this warp waits its turn, and the per-lane form (divergent `lane == 0` branch: R2UR moves,
        // an ELECT waterfall loop around every tcgen05 instruction, descriptors rebuilt per MMA; ~75 instructions
        // per K block) made this thread the slowest stage of the pipeline - 51 % of its life between the first MMA
        // and the commit of a K block (tools/role_profile.py).  Descriptors differ only in their low word.
        constexpr uint32_t idesc = ptx::make_idesc(0u, kBM, BLOCK_N);
        constexpr uint32_t kDescHi = (NODEC && !A16) ? ((512u >> 4) | (1u << 14) | (4u << 29))      // SBO 512 B, version 1, SWIZZLE_64B
                                                     : ((1024u >> 4) | (1u << 14) | (2u << 29));    // SBO 1024 B, version 1, SWIZZLE_128B
        const uint32_t a_lo0 = ((ptx::smem_u32(NODEC ? s_code : s_a) >> 4) & 0x3fffu) | (1u << 16);
        const uint32_t b_lo0 = ((ptx::smem_u32(s_b) >> 4) & 0x3fffu) | (1u << 16);
        PROF_VARS;
        uint32_t stage = 0, phase = 0;
        for (int ti = 0; ti < my_tiles; ++ti) {
            const uint32_t buf = (uint32_t)ti % kAccBufs;
            PROF(a, ptx::mbar_wait(ptx::smem_u32(&bar_tempty[buf]), ((((uint32_t)ti / kAccBufs) & 1u)) ^ 1u, 3u | ((uint32_t)ti << 16)));
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * BLOCK_N;
            for (int kb = 0; kb < p.num_kb; ++kb) {
                PROF(b, ptx::mbar_wait(ptx::smem_u32(&bar_full[stage]), phase, 4u | ((uint32_t)kb << 8) | ((uint32_t)ti << 16)));
                ptx::tc_fence_after();
#ifdef SLFP_ROLE_PROFILE
                const long long prof_t = clock64();
#endif
                if (ptx::elect_one()) {
                    const uint32_t a_lo = a_lo0 + stage * (uint32_t)((NODEC ? C::kCodeTile : kABytes) >> 4);
                    const uint32_t b_lo = b_lo0 + stage * (uint32_t)(C::kBBytes >> 4);
                    const uint32_t a_tmem = tmem_base + (uint32_t)C::kATmemCol + stage * 32u;
#pragma unroll
                    for (int k = 0; k < ((NODEC && !A16) ? kBK / 32 : kBK / 16); ++k) {
                        const uint64_t bd = ((uint64_t)kDescHi << 32) | (uint64_t)(b_lo + 2u * k);
                        if (NODEC && !A16)              // e4m3 x e4m3, K = 32 per instruction: 32 bytes = 2 descriptor units per step
                            ptx::mma_f8_ss(d_tmem, ((uint64_t)kDescHi << 32) | (uint64_t)(a_lo + 2u * k), bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        else if (C::kATmem)
                            ptx::mma_f16_ts(d_tmem, a_tmem + k * 8, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        else
                            ptx::mma_f16_ss(d_tmem, ((uint64_t)kDescHi << 32) | (uint64_t)(a_lo + 2u * k), bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    ptx::mma_commit(ptx::smem_u32(&bar_empty[stage]));     // frees the stage
                    if (kb == p.num_kb - 1) ptx::mma_commit(ptx::smem_u32(&bar_tfull[buf]));   // accumulator ready
                }
                __syncwarp();
#ifdef SLFP_ROLE_PROFILE
                prof_c += (unsigned long long)(clock64() - prof_t);
#endif
                if (++stage == (uint32_t)C::kStages) { stage = 0; phase ^= 1u; }
            }
        }
        if (lane == 0) { PROF_FLUSH(16); }
      }
    } else if (warp < kDecWarp0 + kDecWarps) {
        ptx::setmaxnreg_dec<R::kRegsDec>();
        // =========================== decode: codes -> float16 A tile ==========================================
        // Decode warps work in GROUPS of four (one warp per TMEM lane quadrant = 32 pixel rows each); group g owns
        // the K blocks g, g + G, g + 2G, ... (G = groups), so G K blocks are in flight at once and a warp's
        // wait -> load -> look-up -> store latency chain is paid once per G K blocks instead of on every one.
        // thread = pixel row 32*(warp%4)+lane, all four 16-channel quarters of the K block.  The 64-byte code rows
        // are read through the TMA's 64B swizzle (chunk ^= (row >> 1) & 3): 8 consecutive rows hit 8 bank groups.
        // A waiter must never be two phases ahead of its mbarrier (a parity wait cannot tell phase n from n + 2):
        // the group of K block j waits on empty[j % S] for MMA(j - S) and is only guaranteed MMA(j - G - S), so the
        // number of groups G may not exceed the number of A stages S; with 3 stages the groups are 8 warps wide
        // (two warps per quadrant, two quarters each).
        constexpr int kWarpsPerGroup = (kDecWarps / 4 <= C::kStages) ? 4 : 8;
        constexpr int kDecGroups = kDecWarps / kWarpsPerGroup;
        constexpr int kQPW = 16 / kWarpsPerGroup;              // 16-channel quarters per warp: 4 or 2
        static_assert(kDecGroups <= C::kStages && kDecGroups <= kCodeStages, "groups may not outrun the barrier phases");
        static_assert(kCodeStages % kDecGroups == 0, "a code stage must always be consumed by the same decode group");
        if constexpr (!NODEC) {
        const int grp = (warp - kDecWarp0) / kWarpsPerGroup;
        const int q0 = (((warp - kDecWarp0) % kWarpsPerGroup) >> 2) * kQPW;   // first quarter of this warp
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lut_base = ptx::smem_u32(s_lut);        // 128-byte aligned: (code << 7) | lane*4 never carries
        const uint32_t lane4 = (uint32_t)lane * 4u;
        uint32_t c_off[kQPW], a_off[kQPW][2];
#pragma unroll
        for (int qi = 0; qi < kQPW; ++qi) {
            const int q = q0 + qi;
            c_off[qi] = GRAN == 64 ? (uint32_t)(row * 64 + ((q ^ ((row >> 1) & 3)) << 4)) : (uint32_t)(q * 2048 + row * 16);
            const uint32_t base = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
            a_off[qi][0] = base + (uint32_t)(((q * 2) ^ (row & 7)) << 4);
            a_off[qi][1] = base + (uint32_t)(((q * 2 + 1) ^ (row & 7)) << 4);
        }
        const uint32_t a_base = ptx::smem_u32(s_a);
        const uint32_t code_s = ptx::smem_u32(s_code);
        const uint32_t a_tmem_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)C::kATmemCol;
        const int k16_total = p.taps * p.c16s;
        const int total_kb = my_tiles * p.num_kb;
        PROF_VARS;
        int kb = grp % p.num_kb;                                // K block index inside its tile (for the K-padding test)
        for (int it = grp; it < total_kb; it += kDecGroups) {
            const uint32_t cs = (uint32_t)it % (uint32_t)kCodeStages, cphase = ((uint32_t)it / (uint32_t)kCodeStages) & 1u;
            const uint32_t stage = (uint32_t)it % (uint32_t)C::kStages, phase = ((uint32_t)it / (uint32_t)C::kStages) & 1u;
            PROF(a, ptx::mbar_wait(ptx::smem_u32(&bar_cfull[cs]), cphase, 5u | ((uint32_t)it << 8)));
            PROF(b, ptx::mbar_wait(ptx::smem_u32(&bar_empty[stage]), phase ^ 1u, 6u | ((uint32_t)it << 8)));   // MMA done with this A stage
#ifdef SLFP_ROLE_PROFILE
            const long long prof_t = clock64();
#endif
            if (C::kATmem) ptx::tc_fence_after();
            const uint32_t hsel = (HIFI && kb >= 2 * p.kb_base) ? 0x7632u : 0x5410u;
            const uint32_t a_dst = a_base + stage * kABytes;
#pragma unroll
            for (int qi = 0; qi < kQPW; ++qi) {
                const int q = q0 + qi;
                // 16-channel pieces beyond the last filter tap (K padding) are not loaded: zeros
                const int kbb = HIFI ? kb % p.kb_base : kb;
                const bool valid = GRAN == 64 || (kbb * 4 + q < k16_total);
                uint4 cw = make_uint4(0u, 0u, 0u, 0u);
                if (valid) cw = ptx::lds128_volatile(code_s + cs * kCodeBytes + c_off[qi]);
                const uint32_t w[4] = {cw.x, cw.y, cw.z, cw.w};
                uint32_t h[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t c = w[i];
                    // per code: a shift (IMAD.HI measured slower than SHF), one LOP3 ((x & 0x7f80) | lane*4), one
                    // LDS [reg + constant table base]; two table entries (float16 in the low half) merge with an IMAD
                    const uint32_t e0 = ptx::lds32_off(ptx::and_or(c << 7, 0x7f80u, lane4), lut_base);
                    const uint32_t e1 = ptx::lds32_off(ptx::and_or(c >> 1, 0x7f80u, lane4), lut_base);
                    const uint32_t e2 = ptx::lds32_off(ptx::and_or(c >> 9, 0x7f80u, lane4), lut_base);
                    const uint32_t e3 = ptx::lds32_off(ptx::and_or(c >> 17, 0x7f80u, lane4), lut_base);
                    if (HIFI) {                          // hi halves (passes 0, 1) or lo halves (pass 2) of two entries
                        h[2 * i] = __byte_perm(e0, e1, hsel);
                        h[2 * i + 1] = __byte_perm(e2, e3, hsel);
                    } else {
                        h[2 * i] = ptx::pack16_fma(e0, e1);
                        h[2 * i + 1] = ptx::pack16_fma(e2, e3);
                    }
                }
                if (C::kATmem) {
                    ptx::tmem_st8(a_tmem_lane + stage * 32u + (uint32_t)q * 8u, h);
                } else {
                    ptx::sts128(a_dst + a_off[qi][0], h[0], h[1], h[2], h[3]);
                    ptx::sts128(a_dst + a_off[qi][1], h[4], h[5], h[6], h[7]);
                }
            }
            if (C::kATmem) {
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
            } else {
                ptx::fence_proxy_async_smem();               // generic-proxy writes -> async proxy (UMMA)
            }
            __syncwarp();
            if (lane == 0) {
                ptx::mbar_arrive(ptx::smem_u32(&bar_cempty[cs]));   // the codes were consumed by the look-ups above
                ptx::mbar_arrive(ptx::smem_u32(&bar_full[stage]));
            }
            kb += kDecGroups;
            while (kb >= p.num_kb) kb -= p.num_kb;
#ifdef SLFP_ROLE_PROFILE
            prof_c += (unsigned long long)(clock64() - prof_t);
#endif
        }
        if (warp == kDecWarp0 && lane == 0) { PROF_FLUSH(20); }
        }  // !NODEC
    } else {
        ptx::setmaxnreg_inc<R::kRegsEpi>();
        // =========================== epilogue ===================================================================
        const int quad = warp & 3;                         // TMEM lane quadrant this warp may access
        const int half = (warp - kEpiWarp0) >> 2;
        const int etid = (warp - kEpiWarp0) * 32 + lane;
        const int mode = p.epi_mode;
        const int out_kind = p.e4m3_out ? 1 : (p.epi.store_f16 ? 2 : 0);      // what a codes-only epilogue writes (warp-uniform)
        const bool sfp33 = p.epi.next_fmt == SLFP_FMT_SFP33_RELU || (p.epi.layerout && p.epi.next_fmt == SLFP_FMT_SFP33);
        const uint32_t s_mul = ptx::smem_u32(s_par), s_add = s_mul + BLOCK_N * 4;
        // output table of store_f16: entry address = (raw code << 6) + this, raw code = (bits >> 18 | 19) - base
        const uint32_t olut_rel = (ptx::smem_u32(s_olut) + (uint32_t)(lane % C::kOutLutCopies) * 4u -
                                   (sfp33 ? 0x76Fu : 0xEDFu) * (uint32_t)(C::kOutLutCopies * 4)) | (C::kOutLutCopies == 16 ? 1u : 0u);
        int staged_n0 = -1;
        PROF_VARS;
        // per tile: stage the per-channel vectors when the channel tile changes, then wait for the accumulator
        auto tile_begin = [&](int ti, int tile) -> uint32_t {
            const uint32_t buf = (uint32_t)ti % kAccBufs;
            if (!STG && mode != 0) {
                // the fast modes read the folded affine from shared memory (mode 1: pre-scaled by 1/(16 Ka_next))
                const int n_tile0 = (int)((uint32_t)tile - mdiv((uint32_t)tile, p.k_ntiles) * (uint32_t)p.n_tiles) * BLOCK_N;
                if (n_tile0 != staged_n0) {
                    ptx::bar_sync(1, kEpiWarps * 32);          // every epilogue warp is done with the previous vectors
                    const float sc = mode == 1 ? (p.e4m3_out ? p.rk1 : p.sc1) : 1.0f;      // mode 2 keeps relu(y) itself for the float16 output
                    for (int i = etid; i < BLOCK_N; i += kEpiWarps * 32) {
                        const int n = n_tile0 + i;
                        s_par[i] = n < p.Kout ? __ldg(p.epi.ch_mul + n) * sc : 0.0f;
                        s_par[BLOCK_N + i] = n < p.Kout ? __ldg(p.epi.ch_add + n) * sc : 0.0f;
                    }
                    ptx::bar_sync(1, kEpiWarps * 32);
                    staged_n0 = n_tile0;
                }
            }
            PROF(a, ptx::mbar_wait_backoff(ptx::smem_u32(&bar_tfull[buf]), ((uint32_t)ti / kAccBufs) & 1u, 64, 7u | ((uint32_t)ti << 16)));
            ptx::tc_fence_after();
            return tmem_base + buf * BLOCK_N;
        };
        auto tile_end = [&](int ti) {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bar_tempty[(uint32_t)ti % kAccBufs]));
        };
        if constexpr (STG) {
            // ---- staged epilogue (mode 2 only): every global access of the epilogue is a TMA transfer ----------
            // A thread owns one output pixel (TMEM lane = row), so direct loads / stores touch 32 different rows per
            // instruction and the block tails (float16 residual in, float16 + codes out) were bound by L1TEX
            // wavefronts, not by HBM (profiles/r01_conv_v2.md).  Here the residual tile arrives by TMA one tile
            // ahead (double buffered), is updated IN PLACE to the float16 output and leaves by TMA store together
            // with the code tiles.  The TMA unit's cost is per box ROW (~4-5 cycles for any row <= 128 B, measured:
            // 32 / 64-byte rows made these transfers the bottleneck), so every box has 128-byte rows: the float16
            // tile is two [128 rows x 64 columns] halves, a code tile one [128 x 128] box, all with the 128-byte
            // swizzle (16-byte chunk ^= row & 7), which also keeps the per-row shared accesses conflict free.
            // Rows >= M and columns >= Kout are clipped by TMA.
            //
            // The tails are bound by the epilogue warps' instruction issue (profiles/r03_final.md: 479 SASS instructions per
            // warp and tile, of which only ~250 were arithmetic - the rest re-tested run-time flags per chunk and divided
            // tile indices by n_tiles five times per tile).  Hence: (1) the loop is instantiated per combination of
            // (residual, float16 output, second code tensor) with the flags as compile-time constants - one generic
            // instantiation keeps the run-time tests for e4m3 / layerout outputs; (2) the tile's (m, n) coordinates advance
            // incrementally (no division inside the loop).
            const int cg = half;                                                      // 32-column group 0..3
            const int r = quad * 32 + lane;
            // the staging TMA traffic is issued by ONE elected lane of the first epilogue warp, which reaches these points
            // converged: operands stay in uniform registers (no R2UR / ELECT waterfall around every UTMALDG / UTMASTG)
            const bool lead_warp = warp == kEpiWarp0;
#define SLFP_LEADER (lead_warp && ptx::elect_one())
            const bool has_res = p.epi.residual != nullptr, has_y16 = p.epi.y_f16 != nullptr;
            const bool has_c1 = p.epi.y_codes != nullptr, has_c2 = p.epi.y_codes2 != nullptr;
            constexpr uint32_t kHalfIo = kBM * 128;                                   // one [128 x 64] float16 half
            const uint32_t io0 = ptx::smem_u32(s_stage);                              // buffer b at + b * kIoBytes
            const uint32_t co1 = io0 + 2u * C::kIoBytes, co2 = co1 + C::kCoBytes;
            const uint32_t row_off = (uint32_t)(r * 128), sw = (uint32_t)(r & 7);
            const uint32_t io_t = (uint32_t)(cg >> 1) * kHalfIo + row_off;
            const uint32_t bres = ptx::smem_u32(&bar_res[0]);
            const float sc1 = p.sc1, sc2 = p.sc2;
            const uint32_t enc_sh = sfp33 ? 19u : 18u;
            const int32_t enc_base = sfp33 ? 0x76F : 0xEDF;
            const int n_tiles = p.n_tiles, Kout = p.Kout;
            const int step_m = (int)gridDim.x / n_tiles, step_n = (int)gridDim.x % n_tiles;
            auto run = [&](auto res_c, auto y16_c, auto c2_c, auto gen_c) {
                constexpr bool GEN = decltype(gen_c)::value;
                const bool h_res = GEN ? has_res : decltype(res_c)::value;
                const bool h_y16 = GEN ? has_y16 : decltype(y16_c)::value;
                const bool h_c1 = GEN ? has_c1 : true;
                const bool h_c2 = GEN ? has_c2 : decltype(c2_c)::value;
                const bool e4m3 = GEN ? (p.e4m3_out != 0) : false;
                const bool lay = GEN ? (p.epi.layerout != 0) : false;
                auto advance = [&](int& m, int& n) { m += step_m; n += step_n; if (n >= n_tiles) { n -= n_tiles; ++m; } };
                auto load_res = [&](int m, int n, uint32_t buf) {
                    const int c0 = n * BLOCK_N, r0 = m * kBM;
                    const bool two = c0 + 64 < Kout;
                    ptx::mbar_arrive_expect_tx(bres + buf * 8u, two ? 2u * kHalfIo : kHalfIo);
                    ptx::tma_load_2d(io0 + buf * C::kIoBytes, &omaps.res, bres + buf * 8u, c0, r0);
                    if (two) ptx::tma_load_2d(io0 + buf * C::kIoBytes + kHalfIo, &omaps.res, bres + buf * 8u, c0 + 64, r0);
                };
                // The folded per-channel affine of a tile's 128 columns is staged in shared memory ONE TILE AHEAD by 64 threads
                // (one float4 of mul or add each): read straight from global memory inside the tile it was the largest single
                // stall of the epilogue warps (long-scoreboard waits on 16 LDG.128 per thread and tile).  s_par: [2 buffers]
                // [mul 128 | add 128]; the writes for tile ti + 1 happen during tile ti, whose two CTA-wide epilogue barriers
                // order them before the reads.
                auto stage_affine = [&](int n, uint32_t buf) {
                    if (etid < 64) {
                        const int c = n * BLOCK_N + (etid & 31) * 4;
                        float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (c < Kout) v4 = __ldg(reinterpret_cast<const float4*>((etid < 32 ? p.epi.ch_mul : p.epi.ch_add) + c));
                        *reinterpret_cast<float4*>(s_par + buf * 2 * BLOCK_N + (etid >> 5) * BLOCK_N + (etid & 31) * 4) = v4;
                    }
                };
                int tm = (int)blockIdx.x / n_tiles, tn = (int)blockIdx.x % n_tiles;    // the only division: once per CTA
                if (my_tiles > 0) {
                    if (h_res && SLFP_LEADER) load_res(tm, tn, 0u);
                    stage_affine(tn, 0u);
                    ptx::bar_sync(1, kEpiWarps * 32);
                }
                for (int ti = 0; ti < my_tiles; ++ti) {
                    const uint32_t buf = (uint32_t)ti & 1u;
                    int nm = tm, nn = tn;
                    advance(nm, nn);                                                     // the next tile of this CTA
                    if (ti + 1 < my_tiles) stage_affine(nn, buf ^ 1u);
                    const uint32_t tacc = tile_begin(ti, 0);
                    const uint32_t s_mul_t = s_mul + buf * (uint32_t)(2 * BLOCK_N * 4) + (uint32_t)(cg * 32) * 4u, s_add_t = s_mul_t + BLOCK_N * 4;
                    const int n_slab = tn * BLOCK_N + cg * 32;
                    int nvalid = (Kout - n_slab) >> 4;                                   // 16-column chunks of this group inside Kout
                    nvalid = nvalid > 2 ? 2 : (nvalid < 0 ? 0 : nvalid);
                    const uint32_t io = io0 + buf * C::kIoBytes + io_t;
                    // both 16-column chunks leave TMEM at once; the accumulator buffer is released before any arithmetic
                    uint32_t acc[2][16];
                    uint32_t cw[2][2][4];                                                // packed codes [consumer][chunk]
                    const uint32_t tcol = tacc + ((uint32_t)(quad * 32) << 16) + (uint32_t)(cg * 32);
                    if (nvalid > 0) ptx::tmem_ld16(tcol, acc[0]);
                    if (nvalid > 1) ptx::tmem_ld16(tcol + 16u, acc[1]);
                    if (h_res) ptx::mbar_wait(bres + buf * 8u, ((uint32_t)ti >> 1) & 1u, 8u | ((uint32_t)ti << 16));
                    ptx::tmem_ld_wait();
                    tile_end(ti);
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) {
                        const bool live = ch < nvalid;                                   // group-uniform
                        const uint32_t ci = (uint32_t)((cg & 1) * 4 + 2 * ch);          // chunk index inside the 128-byte row
                        const uint32_t ioa = io + ((ci ^ sw) << 4), iob = io + (((ci + 1u) ^ sw) << 4);
                        float v[16];
                        if (live) {
                            uint4 ra = make_uint4(0u, 0u, 0u, 0u), rb = ra;
                            if (h_res) { ra = ptx::lds128_volatile(ioa); rb = ptx::lds128_volatile(iob); }
                            // per-channel affine: warp-uniform 16-byte shared-memory loads (broadcast) of the staged vectors
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                const float4 m4 = ptx::lds128_f4(s_mul_t + (uint32_t)(ch * 16 + 4 * g) * 4u);
                                const float4 a4 = ptx::lds128_f4(s_add_t + (uint32_t)(ch * 16 + 4 * g) * 4u);
                                v[4 * g + 0] = fmaf(__uint_as_float(acc[ch][4 * g + 0]), m4.x, a4.x);
                                v[4 * g + 1] = fmaf(__uint_as_float(acc[ch][4 * g + 1]), m4.y, a4.y);
                                v[4 * g + 2] = fmaf(__uint_as_float(acc[ch][4 * g + 2]), m4.z, a4.z);
                                v[4 * g + 3] = fmaf(__uint_as_float(acc[ch][4 * g + 3]), m4.w, a4.w);
                            }
                            if (h_res) {
                                const uint32_t rw[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&rw[i]));
                                    v[2 * i] += f.x; v[2 * i + 1] += f.y;
                                }
                            }
                            if (lay) {
#pragma unroll
                                for (int i = 0; i < 16; ++i) v[i] = layerout_relu(v[i]);
                            }
                        }
                        if (ch == 0) {
                            // The previous tile's stores must have drained the other float16 buffer before the next tile's
                            // residual lands in it, and the code staging tiles before this tile's codes are written.  Only the
                            // leader's warp waits here; everybody else goes on with the arithmetic and the in-place float16
                            // update (this tile's float16 buffer is not read by any pending store) and meets the leader at the
                            // barrier below, by which time the drain is long over.
                            if (SLFP_LEADER) {
                                PROF(b, ptx::bulk_wait_read0());
                                if (h_res && ti + 1 < my_tiles) load_res(nm, nn, buf ^ 1u);
                            }
                            __syncwarp();
                        }
                        if (!live) continue;
                        if (h_y16) {
                            uint32_t hw[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) hw[i] = ptx::pack_relu_f16x2(v[2 * i], v[2 * i + 1]);
                            ptx::sts128(ioa, hw[0], hw[1], hw[2], hw[3]);
                            ptx::sts128(iob, hw[4], hw[5], hw[6], hw[7]);
                        }
#pragma unroll
                        for (int pass = 0; pass < 2; ++pass) {
                            if (!(pass ? h_c2 : h_c1)) continue;
                            const float sc = pass ? sc2 : sc1;
                            if (e4m3) {
                                const uint4 q4 = encode16_e4m3_relu(v, pass ? p.rk2 : p.rk1);
                                cw[pass][ch][0] = q4.x; cw[pass][ch][1] = q4.y; cw[pass][ch][2] = q4.z; cw[pass][ch][3] = q4.w;
                                continue;
                            }
                            int32_t t[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i)      // encode_relu_fast_raw16 with run-time format constants
                                t[i] = (int32_t)(__float_as_uint(__saturatef(v[i] * sc)) >> enc_sh) - enc_base;
#pragma unroll
                            for (int i = 0; i < 4; ++i) cw[pass][ch][i] = ptx::pack_sat_u8x4(t[4 * i], t[4 * i + 1], t[4 * i + 2], t[4 * i + 3]);
                        }
                    }
                    PROF(c, ptx::bar_sync(1, kEpiWarps * 32));          // the leader has seen the code staging tiles drained
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) {
                        if (ch >= nvalid) continue;
#pragma unroll
                        for (int pass = 0; pass < 2; ++pass) {
                            if (!(pass ? h_c2 : h_c1)) continue;
                            ptx::sts128((pass ? co2 : co1) + row_off + (((uint32_t)(cg * 2 + ch) ^ sw) << 4), cw[pass][ch][0], cw[pass][ch][1],
                                        cw[pass][ch][2], cw[pass][ch][3]);
                        }
                    }
                    ptx::fence_proxy_async_smem();             // staging writes -> async proxy (TMA store)
                    PROF(c, ptx::bar_sync(1, kEpiWarps * 32));
                    if (SLFP_LEADER) {
                        const int c0 = tn * BLOCK_N, r0 = tm * kBM;
                        if (h_y16) {
                            ptx::tma_store_2d(&omaps.y16, io0 + buf * C::kIoBytes, c0, r0);
                            if (c0 + 64 < Kout) ptx::tma_store_2d(&omaps.y16, io0 + buf * C::kIoBytes + kHalfIo, c0 + 64, r0);
                        }
                        if (h_c1) ptx::tma_store_2d(&omaps.c1, co1, c0, r0);
                        if (h_c2) ptx::tma_store_2d(&omaps.c2, co2, c0, r0);
                        ptx::bulk_commit();
                    }
                    tm = nm; tn = nn;
                }
                if (SLFP_LEADER) ptx::bulk_wait0();
            };
            // Two-group form (one code tensor, the common block tail): the 16 epilogue warps split into two groups of 8 that work
            // on ALTERNATE tiles - group g owns accumulator buffer g, float16 staging tile g (residual in, float16 out, in place),
            // code staging tile g and named barrier 1 + g.  A thread takes 64 columns of its row (four 16-column chunks).  With one
            // group per tile the whole epilogue marched in lock step through wait -> TMEM load -> math -> barrier -> staging ->
            // barrier -> TMA store, and the ncu source view showed the issue slots 44-57 % busy with every phase's latency
            // exposed (profiles/r03_final.md); now one group's math overlaps the other's TMEM / TMA latencies.  The residual of
            // the group's NEXT tile can only be requested once this tile's float16 store has read the shared buffer (in-place
            // update), so its latency is covered by the other group's tile, and the tile after that is prefetched into L2.
            auto run2 = [&](auto res_c, auto y16_c) {
                constexpr bool h_res = decltype(res_c)::value, h_y16 = decltype(y16_c)::value;
                const int grp = (warp - kEpiWarp0) >> 3, gw = (warp - kEpiWarp0) & 7;
                const int chalf = gw >> 2;                                                // 64-column half of the tile
                const bool glead = gw == 0;
#define SLFP_GLEADER (glead && ptx::elect_one())
                const int gtid = gw * 32 + lane;
                const uint32_t iob = io0 + (uint32_t)grp * C::kIoBytes;                   // this group's float16 tile
                const uint32_t cob = grp ? co2 : co1;                                     // this group's code tile
                const uint32_t io_r = iob + (uint32_t)chalf * kHalfIo + row_off;
                const uint32_t brs = bres + (uint32_t)grp * 8u;
                const uint32_t par = s_mul + (uint32_t)grp * (uint32_t)(2 * BLOCK_N * 4); // [mul 128 | add 128] of the group's tile
                const uint32_t par_t = par + (uint32_t)(chalf * 64) * 4u;
                const int bar_id = 1 + grp;
                auto advance = [&](int& m, int& n) { m += step_m; n += step_n; if (n >= n_tiles) { n -= n_tiles; ++m; } };
                auto load_res = [&](int m, int n) {
                    const int c0 = n * BLOCK_N, r0 = m * kBM;
                    const bool two = c0 + 64 < Kout;
                    ptx::mbar_arrive_expect_tx(brs, two ? 2u * kHalfIo : kHalfIo);
                    ptx::tma_load_2d(iob, &omaps.res, brs, c0, r0);
                    if (two) ptx::tma_load_2d(iob + kHalfIo, &omaps.res, brs, c0 + 64, r0);
                };
                auto load_affine = [&](int n) -> float4 {
                    float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (gtid < 64) {
                        const int c = n * BLOCK_N + (gtid & 31) * 4;
                        if (c < Kout) v4 = __ldg(reinterpret_cast<const float4*>((gtid < 32 ? p.epi.ch_mul : p.epi.ch_add) + c));
                    }
                    return v4;
                };
                auto store_affine = [&](const float4& v4) {
                    if (gtid < 64) *reinterpret_cast<float4*>(s_par + grp * 2 * BLOCK_N + (gtid >> 5) * BLOCK_N + (gtid & 31) * 4) = v4;
                };
                int tm = (int)blockIdx.x / n_tiles, tn = (int)blockIdx.x % n_tiles;
                if (grp) advance(tm, tn);                                                 // group 1 starts with the CTA's second tile
                if (grp < my_tiles) {
                    if (h_res && SLFP_GLEADER) load_res(tm, tn);
                    store_affine(load_affine(tn));
                    ptx::bar_sync(bar_id, 256);
                }
                for (int ti = grp; ti < my_tiles; ti += 2) {
                    int nm = tm, nn = tn;
                    advance(nm, nn); advance(nm, nn);                                     // this group's next tile
                    const bool has_next = ti + 2 < my_tiles;
                    const float4 aff_next = has_next ? load_affine(nn) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (!h_res && h_y16) {
                        // no residual load orders the previous float16 store against this tile's in-place writes: wait for it here
                        if (SLFP_GLEADER) ptx::bulk_wait_read0();
                        ptx::bar_sync(bar_id, 256);
                    }
                    const uint32_t tacc = tile_begin(ti, 0);
                    const int n_half = tn * BLOCK_N + chalf * 64;
                    int nvalid = (Kout - n_half) >> 4;                                   // 16-column chunks of this thread inside Kout
                    nvalid = nvalid > 4 ? 4 : (nvalid < 0 ? 0 : nvalid);
                    const uint32_t tcol = tacc + ((uint32_t)(quad * 32) << 16) + (uint32_t)(chalf * 64);
                    uint32_t cw[4][4];                                                   // packed codes per chunk
                    if (h_res) ptx::mbar_wait(brs, ((uint32_t)ti >> 1) & 1u, 8u | ((uint32_t)ti << 16));
#pragma unroll
                    for (int pr = 0; pr < 2; ++pr) {
                        uint32_t acc[2][16];
                        if (2 * pr < nvalid) ptx::tmem_ld16(tcol + (uint32_t)(32 * pr), acc[0]);
                        if (2 * pr + 1 < nvalid) ptx::tmem_ld16(tcol + (uint32_t)(32 * pr + 16), acc[1]);
                        ptx::tmem_ld_wait();
                        if (pr == 1) tile_end(ti);                                       // the accumulator buffer is free again
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            const int ch = 2 * pr + c;
                            if (ch >= nvalid) continue;                                  // warp-uniform
                            const uint32_t ci = (uint32_t)(2 * ch);                      // 16-byte chunk inside the 128-byte row
                            const uint32_t ioa = io_r + ((ci ^ sw) << 4), iob2 = io_r + (((ci + 1u) ^ sw) << 4);
                            float v[16];
                            uint4 ra = make_uint4(0u, 0u, 0u, 0u), rb = ra;
                            if (h_res) { ra = ptx::lds128_volatile(ioa); rb = ptx::lds128_volatile(iob2); }
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                const float4 m4 = ptx::lds128_f4(par_t + (uint32_t)(ch * 16 + 4 * g) * 4u);
                                const float4 a4 = ptx::lds128_f4(par_t + (uint32_t)(BLOCK_N + ch * 16 + 4 * g) * 4u);
                                v[4 * g + 0] = fmaf(__uint_as_float(acc[c][4 * g + 0]), m4.x, a4.x);
                                v[4 * g + 1] = fmaf(__uint_as_float(acc[c][4 * g + 1]), m4.y, a4.y);
                                v[4 * g + 2] = fmaf(__uint_as_float(acc[c][4 * g + 2]), m4.z, a4.z);
                                v[4 * g + 3] = fmaf(__uint_as_float(acc[c][4 * g + 3]), m4.w, a4.w);
                            }
                            if (h_res) {
                                const uint32_t rw[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&rw[i]));
                                    v[2 * i] += f.x; v[2 * i + 1] += f.y;
                                }
                            }
                            if (h_y16) {
                                uint32_t hw[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) hw[i] = ptx::pack_relu_f16x2(v[2 * i], v[2 * i + 1]);
                                ptx::sts128(ioa, hw[0], hw[1], hw[2], hw[3]);
                                ptx::sts128(iob2, hw[4], hw[5], hw[6], hw[7]);
                            }
                            int32_t t[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                t[i] = (int32_t)(__float_as_uint(__saturatef(v[i] * sc1)) >> enc_sh) - enc_base;
#pragma unroll
                            for (int i = 0; i < 4; ++i) cw[ch][i] = ptx::pack_sat_u8x4(t[4 * i], t[4 * i + 1], t[4 * i + 2], t[4 * i + 3]);
                        }
                    }
                    // the code staging tile must have been read by the previous tile's store before it is rewritten
                    if (SLFP_GLEADER) ptx::bulk_wait_read0();
                    ptx::bar_sync(bar_id, 256);
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        if (ch >= nvalid) continue;
                        ptx::sts128(cob + row_off + (((uint32_t)(chalf * 4 + ch) ^ sw) << 4), cw[ch][0], cw[ch][1], cw[ch][2], cw[ch][3]);
                    }
                    if (has_next) store_affine(aff_next);                                // everybody is done reading this tile's vectors
                    ptx::fence_proxy_async_smem();                                       // staging writes -> async proxy (TMA store)
                    ptx::bar_sync(bar_id, 256);
                    if (SLFP_GLEADER) {
                        const int c0 = tn * BLOCK_N, r0 = tm * kBM;
                        if (h_y16) {
                            ptx::tma_store_2d(&omaps.y16, iob, c0, r0);
                            if (c0 + 64 < Kout) ptx::tma_store_2d(&omaps.y16, iob + kHalfIo, c0 + 64, r0);
                        }
                        ptx::tma_store_2d(&omaps.c1, cob, c0, r0);
                        ptx::bulk_commit();
                        if (h_res && has_next) {
                            if (h_y16) ptx::bulk_wait_read0();                           // in place: the store must have read the tile
                            load_res(nm, nn);
                            int pm = nm, pn = nn;
                            advance(pm, pn); advance(pm, pn);
                            if (ti + 4 < my_tiles) {                                     // the tile after that: towards L2
                                ptx::tma_prefetch_2d(&omaps.res, pn * BLOCK_N, pm * kBM);
                                if (pn * BLOCK_N + 64 < Kout) ptx::tma_prefetch_2d(&omaps.res, pn * BLOCK_N + 64, pm * kBM);
                            }
                        }
                    }
                    __syncwarp();
                    tm = nm; tn = nn;
                }
                if (SLFP_GLEADER) ptx::bulk_wait0();
#undef SLFP_GLEADER
            };
            using T = std::true_type;
            using F = std::false_type;
            if (p.stg_groups == 2) {                                                     // (host: one code tensor, no e4m3 / layerout)
                if (has_res && has_y16) run2(T{}, T{});
                else if (has_res) run2(T{}, F{});
                else if (has_y16) run2(F{}, T{});
                else run2(F{}, F{});
            } else
            if (p.e4m3_out || p.epi.layerout || !has_c1) run(F{}, F{}, F{}, T{});        // generic: run-time flags
            else if (has_res && has_y16 && !has_c2) run(T{}, T{}, F{}, F{});              // block tail inside a stage
            else if (has_res && !has_y16 && has_c2) run(T{}, F{}, T{}, F{});              // last tail of a stage (two consumers)
            else if (has_res && has_y16 && has_c2) run(T{}, T{}, T{}, F{});
            else if (has_res && !has_y16 && !has_c2) run(T{}, F{}, F{}, F{});
            else if (!has_res && has_y16 && !has_c2) run(F{}, T{}, F{}, F{});             // fused dual tail
            else if (!has_res && has_y16 && has_c2) run(F{}, T{}, T{}, F{});
            else if (!has_res && !has_y16 && has_c2) run(F{}, F{}, T{}, F{});
            else run(F{}, F{}, F{}, F{});
#undef SLFP_LEADER
        } else
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
            const int next_tile = ti + 1 < my_tiles ? tile + (int)gridDim.x : p.num_tiles;
            const uint32_t tacc = tile_begin(ti, tile);
            if (kGroups == 2) {
                if (mode == 1) {
#define SLFP_EPI1(SF, O) epilogue_fast<BLOCK_N, 2, 1, SF, O>(p, tile, next_tile, tacc, quad, half, lane, s_mul, s_add, olut_rel)
                    if (sfp33) { if (out_kind == 0) SLFP_EPI1(true, 0); else if (out_kind == 1) SLFP_EPI1(true, 1); else SLFP_EPI1(true, 2); }
                    else { if (out_kind == 0) SLFP_EPI1(false, 0); else if (out_kind == 1) SLFP_EPI1(false, 1); else SLFP_EPI1(false, 2); }
#undef SLFP_EPI1
                } else if (mode == 2) {
                    if (sfp33) epilogue_fast<BLOCK_N, 2, 2, true>(p, tile, next_tile, tacc, quad, half, lane, s_mul, s_add, olut_rel);
                    else epilogue_fast<BLOCK_N, 2, 2, false>(p, tile, next_tile, tacc, quad, half, lane, s_mul, s_add, olut_rel);
                } else {
                    epilogue_slab<BLOCK_N, 2>(p, tile, tacc, quad, half, lane);
                }
            } else {
                if (mode == 1) {
#define SLFP_EPI1(SF, O) epilogue_fast16<BLOCK_N, 4, 1, SF, O>(p, tile, next_tile, tacc, quad, half, lane, s_mul, s_add, olut_rel)
                    if (sfp33) { if (out_kind == 0) SLFP_EPI1(true, 0); else if (out_kind == 1) SLFP_EPI1(true, 1); else SLFP_EPI1(true, 2); }
                    else { if (out_kind == 0) SLFP_EPI1(false, 0); else if (out_kind == 1) SLFP_EPI1(false, 1); else SLFP_EPI1(false, 2); }
#undef SLFP_EPI1
                } else if (mode == 2) {
                    if (sfp33) epilogue_fast16<BLOCK_N, 4, 2, true>(p, tile, next_tile, tacc, quad, half, lane, s_mul, s_add, olut_rel);
                    else epilogue_fast16<BLOCK_N, 4, 2, false>(p, tile, next_tile, tacc, quad, half, lane, s_mul, s_add, olut_rel);
                } else {
                    epilogue_slab<BLOCK_N, 4>(p, tile, tacc, quad, half, lane);
                }
            }
            tile_end(ti);
        }
        if (warp == kEpiWarp0 && lane == 0) { PROF_FLUSH(24); }
    }

    // ---- teardown ---------------------------------------------------------------------------------
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == kWarpMma) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
    }
}

// ---- host side -----------------------------------------------------------------------------------------
template <typename PFN>
static PFN driver_fn(const char* name) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint(name, &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
        return reinterpret_cast<PFN>(ptr);
    cudaGetLastError();
    return nullptr;
}

template <int BLOCK_N, int GRAN, int DW, bool STG = false, bool HIFI = false, bool NODEC = false, bool A16 = false, bool ACC1 = false>
static int launch(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& tx2, const OutMaps& om, const Params& p, cudaStream_t st) {
    using C = Cfg<BLOCK_N, STG, NODEC, A16, ACC1>;
    auto kern = conv_igemm_v2_kernel<BLOCK_N, GRAN, DW, STG, HIFI, NODEC, A16, ACC1>;
    static DeviceOnce attr_once;
    bool& attr_done = attr_once.flag();
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
        if (e != cudaSuccess) return set_error((int)e, "conv_igemm_v2: smem attribute (%d B): %s", C::kSmemBytes, cudaGetErrorString(e));
        attr_done = true;
    }
    const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
    static const bool no_pdl = getenv("SLFP_NO_PDL") != nullptr;
    if (no_pdl) {
        kern<<<grid, kThreads, C::kSmemBytes, st>>>(tx, tw, tx2, om, p);
        return check_launch("conv_igemm_v2_kernel");
    }
    // programmatic stream serialization: this kernel's prologue may overlap the tail of the previous kernel in the
    // stream (it waits with griddepcontrol.wait before its first dependent access); also captured into CUDA graphs
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)kThreads);
    cfg.dynamicSmemBytes = C::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, tx, tw, tx2, om, p);
    if (le != cudaSuccess) return set_error((int)le, "conv_igemm_v2_kernel: %s", cudaGetErrorString(le));
    return check_launch("conv_igemm_v2_kernel");
}

}  // namespace v2

}  // namespace slfp

// Debug aid: install a host-mapped buffer (>= 2 x 8 bytes) that a timed-out barrier wait of the dense conv kernel
// records (tag << 32 | block << 8 | warp) into before it traps.  NULL removes it.
extern "C" int slfp_debug_set_buffer(void* host_mapped_device_ptr) {
    unsigned long long* p = (unsigned long long*)host_mapped_device_ptr;
    cudaError_t e = cudaMemcpyToSymbol(slfp::ptx::g_slfp_dbg, &p, sizeof(p));
    return e == cudaSuccess ? 0 : slfp::set_error((int)e, "slfp_debug_set_buffer: %s", cudaGetErrorString(e));
}

namespace slfp {

bool conv2d_fwd_dense_v2_supported(const SlfpConvDesc* d) {
    static const bool disabled = getenv("SLFP_CONV_V1") != nullptr;
    return !disabled && d->groups == 1 && d->c_phys % 16 == 0 && d->pad_h < 128 && d->pad_w < 128 &&
           (d->r - 1) * d->dil_h < 256 && (d->s - 1) * d->dil_w < 256 && d->stride_h <= 8 && d->stride_w <= 8;
}

int conv2d_fwd_dense_v2_impl(const SlfpConvDesc* d, const uint8_t* x_codes, const SlfpConvDesc* d2, const uint8_t* x2_codes,
                             const void* w_f16, const SlfpEpilogue* epi, cudaStream_t st);

int conv2d_fwd_dense_v2(const SlfpConvDesc* d, const uint8_t* x_codes, const void* w_f16, const SlfpEpilogue* epi,
                        cudaStream_t st) {
    return conv2d_fwd_dense_v2_impl(d, x_codes, nullptr, nullptr, w_f16, epi, st);
}

int conv2d_fwd_dense_v2_impl(const SlfpConvDesc* d, const uint8_t* x_codes, const SlfpConvDesc* d2, const uint8_t* x2_codes,
                             const void* w_f16, const SlfpEpilogue* epi, cudaStream_t st) {
    using namespace v2;
    const bool fast = epi->next_fmt == SLFP_FMT_SLFP34_RELU || epi->next_fmt == SLFP_FMT_SFP33_RELU;
    if (d->fmt != SLFP_FMT_SLFP34_ACT && d->fmt != SLFP_FMT_SFP33 && d->fmt != SLFP_FMT_SLFP34_RELU && d->fmt != SLFP_FMT_SFP33_RELU &&
        d->fmt != SLFP_FMT_SFP33_SFAST && d->fmt != SLFP_FMT_E4M3 && d->fmt != SLFP_FMT_F16Q)
        return set_error(SLFP_ERR_BAD_ARG, "conv2d_fwd: activation code format %d", d->fmt);
    if (epi->y_codes && (epi->k_phys_out % 16 != 0 || epi->k_phys_out < d->k))
        return set_error(SLFP_ERR_BAD_ARG, "conv2d_fwd: k_phys_out=%d", epi->k_phys_out);
    if (epi->y_codes && fast && !epi->relu)
        return set_error(SLFP_ERR_BAD_ARG, "conv2d_fwd: post-ReLU code formats need relu=1");
    const bool e4m3_out = epi->y_codes && epi->next_fmt == SLFP_FMT_E4M3;
    const bool a16 = d->fmt == SLFP_FMT_F16Q;       // the activation tensor is the float16 A operand itself
    const bool fold_w = (d->flags & SLFP_CONV_FOLD_W) != 0;
    if (a16 && (d->c_phys % 64 != 0 || (d->flags & ~SLFP_CONV_FOLD_W) != 0 || d2))
        return set_error(SLFP_ERR_BAD_ARG, "conv2d_fwd: SLFP_FMT_F16Q input needs c_phys %% 64 == 0, no operand flags, no second input");
    if (fold_w && (!a16 || d->c_phys != 64 || d->s != 1 || d->stride_h != 1 || d->stride_w != 1 || d->pad_h || d->pad_w || d->pad_h_extra ||
                   d->pad_w_extra || d->dil_h != 1 || d->dil_w != 1))
        return set_error(SLFP_ERR_BAD_ARG, "conv2d_fwd: SLFP_CONV_FOLD_W needs SLFP_FMT_F16Q, c_phys == 64, an R x 1 filter, stride 1, no padding");
    if (epi->store_f16 && (!epi->y_codes || !fast || epi->y_codes2 || epi->y_f16 || epi->y_f32 || epi->residual || epi->layerout))
        return set_error(SLFP_ERR_BAD_ARG, "conv2d_fwd: store_f16 is a form of the codes-only epilogue (post-ReLU format, one consumer)");
    const bool nodec = (d->flags & SLFP_CONV_E4M3_OPERANDS) != 0;
    if (nodec && (d->fmt != SLFP_FMT_E4M3 || d->c_phys % 64 != 0 || (d->flags & SLFP_CONV_SPLIT_OPERANDS) || d2))
        return set_error(SLFP_ERR_BAD_ARG, "conv2d_fwd: e4m3 operands need SLFP_FMT_E4M3 codes, c_phys %% 64 == 0, no split operands, no second input");
    if (epi->y_codes && !fast && !e4m3_out && epi->next_fmt != SLFP_FMT_SLFP34_ACT && epi->next_fmt != SLFP_FMT_SFP33)
        return set_error(SLFP_ERR_BAD_ARG, "conv2d_fwd: next_fmt=%d", epi->next_fmt);
    if ((((uintptr_t)x_codes | (uintptr_t)w_f16 | (uintptr_t)epi->y_f32 | (uintptr_t)epi->y_f16 |
          (uintptr_t)epi->y_codes | (uintptr_t)epi->y_codes2) & 15u) != 0)
        return set_error(SLFP_ERR_BAD_ARG, "conv2d_fwd: tensors must be 16-byte aligned");
    Params p;
    const int Ho = (d->h + 2 * d->pad_h + d->pad_h_extra - d->dil_h * (d->r - 1) - 1) / d->stride_h + 1;
    const int Wo = (d->w + 2 * d->pad_w + d->pad_w_extra - d->dil_w * (d->s - 1) - 1) / d->stride_w + 1;
    if (Ho <= 0 || Wo <= 0 || d->n <= 0) return set_error(SLFP_ERR_BAD_ARG, "conv2d_fwd: empty output");
    if (d->pad_h + d->pad_h_extra < 0 || d->pad_w + d->pad_w_extra < 0)
        return set_error(SLFP_ERR_BAD_ARG, "conv2d_fwd: negative end padding");
    const unsigned long long M64 = (unsigned long long)d->n * Ho * Wo;
    if (M64 >= (1ull << 31)) return set_error(SLFP_ERR_UNSUPPORTED, "conv2d_fwd: more than 2^31 output pixels");
    p.M = (uint32_t)M64;
    p.Kout = d->k;
    p.HoWo = Ho * Wo; p.Wo = Wo;
    p.S = d->s; p.sh = d->stride_h; p.sw = d->stride_w; p.ph = d->pad_h; p.pw = d->pad_w; p.dh = d->dil_h; p.dw = d->dil_w;
    p.taps = d->r * d->s;
    size_t pitch = slfp_conv_wpitch(d);
    p.nkb1 = (int)(pitch / kBK);
    p.sh2 = p.sw2 = 1;
    if (d2) {
        // second input concatenated along K: 1x1, stride only, same batch / output size / output channels / code format
        const int Ho2 = (d2->h - 1) / d2->stride_h + 1, Wo2 = (d2->w - 1) / d2->stride_w + 1;
        if (!x2_codes || d2->r != 1 || d2->s != 1 || d2->pad_h || d2->pad_w || d2->pad_h_extra || d2->pad_w_extra || d2->groups != 1 ||
            d2->c_phys % 64 != 0 || d->c_phys % 64 != 0 || d2->n != d->n || d2->k != d->k || Ho2 != Ho || Wo2 != Wo || d2->fmt != d->fmt ||
            ((uintptr_t)x2_codes & 15u))
            return set_error(SLFP_ERR_BAD_ARG, "conv2d_fwd_dual: the second input must be a 1x1 / no-padding conv with the same n, k, output size and code format");
        pitch += slfp_conv_wpitch(d2);
        p.sh2 = d2->stride_h; p.sw2 = d2->stride_w;
    }
    const bool hifi = (d->flags & SLFP_CONV_SPLIT_OPERANDS) != 0;
    if (hifi && d2) return set_error(SLFP_ERR_UNSUPPORTED, "conv2d_fwd_dual: no split-operand mode for the fused block tail");
    p.kb_base = (int)(pitch / kBK);
    p.num_kb = hifi ? 3 * p.kb_base : p.kb_base;
    p.w_lo_col = (int)pitch;
    const size_t wrow = hifi ? 2 * pitch : pitch;             // halves per weight row ([hi | lo] in the split-operand mode)
    p.cblocks = (d->c_phys % 64 == 0) ? d->c_phys / 64 : 0;
    p.c16s = d->c_phys / 16;
    p.act_fmt = d->fmt;
    p.epi = *epi;
    p.next_div = make_divk(epi->y_codes ? epi->next_k_div : 1.0f);
    p.next_div2 = make_divk(epi->y_codes2 ? epi->next_k_div2 : 1.0f);
    p.rk1 = (float)(1.0 / (double)(epi->y_codes ? epi->next_k_div : 1.0f));
    p.rk2 = (float)(1.0 / (double)(epi->y_codes2 ? epi->next_k_div2 : 1.0f));

    p.sc1 = (float)(1.0 / (16.0 * (double)(epi->y_codes ? epi->next_k_div : 1.0f)));
    p.sc2 = (float)(1.0 / (16.0 * (double)(epi->y_codes2 ? epi->next_k_div2 : 1.0f)));
    p.e4m3_out = e4m3_out ? 1 : 0;
    p.epi_mode = 0;
    {
        static const bool no_fast = getenv("SLFP_EPI_GENERIC") != nullptr;
        // layerout: the fast epilogues implement relu(quantize_layerout(y)) with exact 0 -> 0 (layerout == 2); the reference's
        // NaN at exact 0 (layerout == 1) stays with the generic epilogue
        const bool common = !no_fast && epi->layerout != 1 && epi->ch_mul && epi->ch_add && epi->relu && d->k % 16 == 0 && !epi->y_f32 &&
                            (!epi->y_codes || ((e4m3_out || (fast != (epi->layerout != 0))) && epi->k_phys_out == d->k && epi->next_k_div > 0.f)) &&
                            (!epi->y_codes2 || (epi->y_codes && epi->next_k_div2 > 0.f)) &&
                            (!epi->residual || epi->residual_f16);
        if (common) p.epi_mode = (epi->y_codes && !epi->y_codes2 && !epi->y_f16 && !epi->residual && !epi->layerout) ? 1 : 2;
    }
    // Staged (TMA) epilogue for the epilogue-bound mode-2 layers (block tails, short-K fused tails): 128-column tiles.
    static const int stg_max_kb = getenv("SLFP_STG_MAXKB") ? atoi(getenv("SLFP_STG_MAXKB")) : 8;
    if (epi->store_f16 && p.epi_mode != 1)
        return set_error(SLFP_ERR_UNSUPPORTED, "conv2d_fwd: store_f16 needs the codes-only fast epilogue (folded affine, relu, k %% 16 == 0)");
    const bool stg = !hifi && !nodec && p.epi_mode == 2 && !(epi->layerout && epi->y_codes) && p.cblocks && d->k > 64 && p.num_kb <= stg_max_kb &&
                     (((uintptr_t)epi->residual) & 15u) == 0;
    static const bool stg_one = getenv("SLFP_STG_ONE_GROUP") != nullptr;
    p.stg_groups = (stg && !stg_one && epi->y_codes && !epi->y_codes2 && !e4m3_out && !epi->layerout) ? 2 : 1;
    int bn = stg ? 128 : (d->k > 128 ? 256 : (d->k > 64 ? 128 : 64));
    p.m_tiles = (int)((p.M + kBM - 1) / kBM);
    // single-accumulator 256-column decode variant (ACC1): whole 64-channel K blocks, K-heavy layers (the epilogue of a tile is
    // not overlapped: >= 8 K blocks keep it below ~15 % of the tile)
    static const bool no_acc1 = getenv("SLFP_NO_ACC1") != nullptr;
    const bool acc1_ok = !no_acc1 && !hifi && !nodec && !a16 && !stg && p.cblocks && p.num_kb >= 8 && p.epi_mode != 0;
    if (epi->store_f16 && !a16 && !acc1_ok && bn > 128) bn = 128;   // the double-buffered 256-column decode variant has no room for the output table
    if (a16 && !stg && bn > 64) {
        // same wave-quantisation model for the no-decode float16 form: a K block costs about its MMA time
        // (256 / 128 / 64 cycles) plus a fixed ~60
        const int sms = num_sms();
        double best = 1e30;
        int best_bn = bn;
        for (int cand = bn; cand >= 64; cand >>= 1) {
            const long tiles = (long)p.m_tiles * ((d->k + cand - 1) / cand);
            const double t = (double)((tiles + sms - 1) / sms) * (cand + 60.0);
            if (t < best * 0.97) { best = t; best_bn = cand; }
        }
        bn = best_bn;
    }
    if (!stg && !nodec && !a16 && bn > 64) {
        // Tile width against wave quantisation (small-M layers: VGG-16 @2x2 / @4x4, ResNet-50 stage 4).  One CTA per SM works
        // through ceil(tiles / SMs) tiles; a tile costs ~num_kb x (cycles per K block), measured ~1 180 for the 256-column
        // form (A staged in shared memory) and ~640 for 128 / 64 columns (A in tensor memory) - the decode side, not the
        // MMA, sets the pace (profiles/r02_final.md).  Pick the width with the smallest modelled time; ties keep the wider tile.
        static const bool fixed = getenv("SLFP_BN_FIXED") != nullptr;
        const int sms = num_sms();
        double best = 1e30;
        int best_bn = bn;
        for (int cand = bn; cand >= 64 && !fixed; cand >>= 1) {
            const long tiles = (long)p.m_tiles * ((d->k + cand - 1) / cand);
            const double t = (double)((tiles + sms - 1) / sms) * (cand == 256 ? (acc1_ok ? 640.0 + 2200.0 / p.num_kb : 1180.0) : 640.0);
            if (t < best * 0.97) { best = t; best_bn = cand; }
        }
        bn = best_bn;
    }
    p.n_tiles = (d->k + bn - 1) / bn;
    p.num_tiles = p.m_tiles * p.n_tiles;
    p.k_ntiles = make_magic((uint32_t)p.n_tiles);
    p.k_howo = make_magic((uint32_t)p.HoWo);
    p.k_wo = make_magic((uint32_t)p.Wo);

    CUtensorMap tmap_x2;
    static auto enc_tiled = driver_fn<PFN_cuTensorMapEncodeTiled_v12000>("cuTensorMapEncodeTiled");
    static auto enc_im2col = driver_fn<PFN_cuTensorMapEncodeIm2col_v12000>("cuTensorMapEncodeIm2col");
    if (!enc_tiled || !enc_im2col) return set_error(SLFP_ERR_DRIVER, "conv2d_fwd: cuTensorMapEncode{Tiled,Im2col} not available");
    CUtensorMap tmap_w, tmap_x;
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)wrow, (cuuint64_t)d->k};
        const cuuint64_t gstr[1] = {(cuuint64_t)wrow * (nodec ? 1 : 2)};
        const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)bn};
        const cuuint32_t estr[2] = {1, 1};
        // e4m3 operands: byte rows of 64 with the 64-byte swizzle (the layout of the code tiles); float16: 128-byte rows
        CUresult cr = enc_tiled(&tmap_w, nodec ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w_f16),
                                gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                nodec ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return set_error(SLFP_ERR_DRIVER, "conv2d_fwd: cuTensorMapEncodeTiled failed (%d)", (int)cr);
    }
    p.plain_1x1 = (d->r == 1 && d->s == 1 && d->stride_h == 1 && d->stride_w == 1 && d->pad_h == 0 && d->pad_w == 0 &&
                   d->pad_h_extra == 0 && d->pad_w_extra == 0 && p.cblocks && getenv("SLFP_NO_PLAIN_1X1") == nullptr) ? 1 : 0;
    // SLFP_FMT_F16Q: 2-byte elements, 128-byte rows with the 128-byte swizzle (the MMA's canonical K-major operand layout)
    const cuuint64_t esz = a16 ? 2u : 1u;
    const CUtensorMapDataType x_type = a16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8;
    const CUtensorMapSwizzle x_swz64 = a16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    if (p.plain_1x1) {
        // 1x1 / stride 1: A is the [M, C] code matrix itself - a tiled 2-D load (rows beyond M zero-filled)
        const cuuint64_t gdim[2] = {(cuuint64_t)d->c_phys, (cuuint64_t)p.M};
        const cuuint64_t gstr[1] = {(cuuint64_t)d->c_phys * esz};
        const cuuint32_t box[2] = {64u, (cuuint32_t)kBM};
        const cuuint32_t estr[2] = {1, 1};
        CUresult cr = enc_tiled(&tmap_x, x_type, 2, const_cast<uint8_t*>(x_codes), gdim, gstr, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, x_swz64, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return set_error(SLFP_ERR_DRIVER, "conv2d_fwd: cuTensorMapEncodeTiled(codes) failed (%d)", (int)cr);
    } else {
        // NHWC codes as a (C, W, H, N) tensor; the pixel bounding box is the set of filter-window origins:
        // lower corner = -pad, upper corner = pad - (filter - 1) * dilation (relative to the tensor's far edge),
        // traversed with the convolution stride; the filter tap (r, s) is the per-load offset.
        const cuuint64_t gdim[4] = {(cuuint64_t)d->c_phys, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->n};
        // SLFP_CONV_FOLD_W: virtual pixels of 64 channels every 16 elements of a row of w + 3 physical pixels (they overlap)
        const cuuint64_t pix = fold_w ? 16u * esz : (cuuint64_t)d->c_phys * esz, rowb = fold_w ? (cuuint64_t)(d->w + 3) * pix : (cuuint64_t)d->w * pix;
        const cuuint64_t gstr[3] = {pix, rowb, rowb * d->h};
        const int lower[2] = {-d->pad_w, -d->pad_h};
        const int upper[2] = {d->pad_w + d->pad_w_extra - (d->s - 1) * d->dil_w, d->pad_h + d->pad_h_extra - (d->r - 1) * d->dil_h};
        const cuuint32_t estr[4] = {1, (cuuint32_t)d->stride_w, (cuuint32_t)d->stride_h, 1};
        CUresult cr = enc_im2col(&tmap_x, x_type, 4, const_cast<uint8_t*>(x_codes), gdim, gstr, lower, upper,
                                 (cuuint32_t)(p.cblocks ? 64 : 16), (cuuint32_t)kBM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 // a decode thread owns a pixel row: 64-byte rows are swizzled so that the 16-byte chunks
                                 // of 8 consecutive rows fall into 8 different bank groups
                                 p.cblocks ? x_swz64 : CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return set_error(SLFP_ERR_DRIVER, "conv2d_fwd: cuTensorMapEncodeIm2col failed (%d)", (int)cr);
    }
    // role split: the epilogue-heavy split (16 epilogue warps) when the per-tile epilogue work (elements x
    // instructions per element of the chosen epilogue) exceeds the per-tile decode work (codes x 3.5)
    static const char* force = getenv("SLFP_CONV_ROLES");          // "dec" / "epi": tuning override
    const double dec_work = (double)p.num_kb * kBM * kBK * 3.5;
    const double epi_work = (double)kBM * bn * (p.epi_mode == 1 ? 6.5 : (p.epi_mode == 2 ? 13.0 : 0.0));
    bool epi_heavy = p.epi_mode != 0 && epi_work > dec_work;
    if (force && force[0] == 'd') epi_heavy = false;
    if (force && force[0] == 'e' && p.epi_mode != 0) epi_heavy = true;
    tmap_x2 = tmap_x;
    if (d2) {
        const cuuint64_t gdim[4] = {(cuuint64_t)d2->c_phys, (cuuint64_t)d2->w, (cuuint64_t)d2->h, (cuuint64_t)d2->n};
        const cuuint64_t gstr[3] = {(cuuint64_t)d2->c_phys, (cuuint64_t)d2->c_phys * d2->w, (cuuint64_t)d2->c_phys * d2->w * d2->h};
        const int lower[2] = {0, 0}, upper[2] = {0, 0};
        const cuuint32_t estr[4] = {1, (cuuint32_t)d2->stride_w, (cuuint32_t)d2->stride_h, 1};
        CUresult cr = enc_im2col(&tmap_x2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<uint8_t*>(x2_codes), gdim, gstr, lower, upper, 64u,
                                 (cuuint32_t)kBM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return set_error(SLFP_ERR_DRIVER, "conv2d_fwd_dual: cuTensorMapEncodeIm2col failed (%d)", (int)cr);
    }
    OutMaps om;
    memset(&om, 0, sizeof(om));
    if (stg) {
        auto out_map = [&](CUtensorMap* m, const void* ptr, bool f16) -> bool {
            const cuuint64_t gdim[2] = {(cuuint64_t)d->k, (cuuint64_t)p.M};
            const cuuint64_t gstr[1] = {(cuuint64_t)d->k * (f16 ? 2u : 1u)};
            const cuuint32_t box[2] = {f16 ? 64u : 128u, (cuuint32_t)kBM};          // 128-byte rows
            const cuuint32_t estr[2] = {1, 1};
            return enc_tiled(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), gdim, gstr,
                             box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        };
        bool ok = true;
        if (epi->residual) ok = ok && out_map(&om.res, epi->residual, true);
        if (epi->y_f16) ok = ok && out_map(&om.y16, epi->y_f16, true);
        if (epi->y_codes) ok = ok && out_map(&om.c1, epi->y_codes, false);
        if (epi->y_codes2) ok = ok && out_map(&om.c2, epi->y_codes2, false);
        if (!ok) return set_error(SLFP_ERR_DRIVER, "conv2d_fwd: cuTensorMapEncodeTiled(outputs) failed");
        if (a16) return launch<128, 64, 8, true, false, true, true>(tmap_x, tmap_w, tmap_x2, om, p, st);
        return launch<128, 64, 8, true>(tmap_x, tmap_w, tmap_x2, om, p, st);
    }
    if (a16) {
        if (bn == 64) return launch<64, 64, 8, false, false, true, true>(tmap_x, tmap_w, tmap_x2, om, p, st);
        if (bn == 128) return launch<128, 64, 8, false, false, true, true>(tmap_x, tmap_w, tmap_x2, om, p, st);
        return launch<256, 64, 8, false, false, true, true>(tmap_x, tmap_w, tmap_x2, om, p, st);
    }
#define SLFP_V2_HIFI(BN)                                                                                         \
    if (hifi && bn == BN) {                                                                                      \
        if (p.cblocks) return launch<BN, 64, 16, false, true>(tmap_x, tmap_w, tmap_x2, om, p, st);               \
        return launch<BN, 16, 16, false, true>(tmap_x, tmap_w, tmap_x2, om, p, st);                              \
    }
    SLFP_V2_HIFI(64)
    SLFP_V2_HIFI(128)
    SLFP_V2_HIFI(256)
#undef SLFP_V2_HIFI
    if (nodec) {
        // no decode role: every non-control warp may as well help the epilogue (DW = 8: 16 epilogue warps)
        if (bn == 64) return launch<64, 64, 8, false, false, true>(tmap_x, tmap_w, tmap_x2, om, p, st);
        if (bn == 128) return launch<128, 64, 8, false, false, true>(tmap_x, tmap_w, tmap_x2, om, p, st);
        return launch<256, 64, 8, false, false, true>(tmap_x, tmap_w, tmap_x2, om, p, st);
    }
    if (bn == 256 && acc1_ok) return launch<256, 64, 16, false, false, false, false, true>(tmap_x, tmap_w, tmap_x2, om, p, st);
#define SLFP_V2_CASE(BN)                                                                                         \
    if (bn == BN) {                                                                                              \
        if (p.cblocks) return epi_heavy ? launch<BN, 64, 8>(tmap_x, tmap_w, tmap_x2, om, p, st) : launch<BN, 64, 16>(tmap_x, tmap_w, tmap_x2, om, p, st); \
        return epi_heavy ? launch<BN, 16, 8>(tmap_x, tmap_w, tmap_x2, om, p, st) : launch<BN, 16, 16>(tmap_x, tmap_w, tmap_x2, om, p, st);  \
    }
    SLFP_V2_CASE(64)
    SLFP_V2_CASE(128)
    SLFP_V2_CASE(256)
#undef SLFP_V2_CASE
    return set_error(SLFP_ERR_UNSUPPORTED, "conv2d_fwd: no tile for k=%d", d->k);
}

}  // namespace slfp
