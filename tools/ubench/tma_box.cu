// Micro-benchmark: TMA tiled-load throughput per SM as a function of the box shape and of the loads in flight.
//   One CTA per SM, one thread issues 2-D tiled loads of [ROWS rows x ROWB bytes] boxes from an L2-resident matrix
//   into a ring of D stages and waits for them in order; prints cycles per load and bytes per cycle per SM.
//   Shapes: the dense conv kernel's code tiles (128 rows x 64 B, 64B swizzle; 128 x 16 B for the stem) against
//   128-byte rows.  Build: nvcc -arch=sm_100a -O3 -o tma_box tma_box.cu -lcuda   (run: ./tma_box)
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

template <int D>
__global__ void __launch_bounds__(128, 1) k_tma(const __grid_constant__ CUtensorMap tm, int box_bytes, int rows_per_box, int total_rows,
                                                int loads, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[D];
    if (threadIdx.x == 0) {
        for (int i = 0; i < D; ++i) mbar_init(smem_u32(&bar[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nbox = total_rows / rows_per_box;
        int box = (blockIdx.x * 977) % nbox;
        const long long t0 = clock64();
        for (int i = 0; i < loads + D; ++i) {
            const int st = i % D;
            if (i >= D) mbar_wait(smem_u32(&bar[st]), ((i / D) - 1) & 1);
            if (i < loads) {
                mbar_expect(smem_u32(&bar[st]), box_bytes);
                tma2d(smem_u32(smem + st * 16384), &tm, smem_u32(&bar[st]), 0, box * rows_per_box);
                box += 131; if (box >= nbox) box -= nbox;
            }
        }
        out[blockIdx.x] = clock64() - t0;
    }
}

int main() {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    const int total_rows = 1 << 18;                          // x 256 B pitch = 64 MB: L2 resident after the warm-up
    uint8_t* x; cudaMalloc(&x, (size_t)total_rows * 256); cudaMemset(x, 1, (size_t)total_rows * 256);
    long long* out; cudaMalloc(&out, 148 * 8);
    struct Shape { const char* name; int rowb, rows, pitch; CUtensorMapSwizzle sw; };
    const Shape shapes[] = {
        {"128 rows x  16 B, pitch  16 (stem codes)        ", 16, 128, 16, CU_TENSOR_MAP_SWIZZLE_NONE},
        {"128 rows x  64 B, pitch  64 (64-ch codes)       ", 64, 128, 64, CU_TENSOR_MAP_SWIZZLE_64B},
        {"128 rows x  64 B, pitch 256 (256-ch codes)      ", 64, 128, 256, CU_TENSOR_MAP_SWIZZLE_64B},
        {"128 rows x 128 B, pitch 128                     ", 128, 128, 128, CU_TENSOR_MAP_SWIZZLE_128B},
        {"128 rows x 128 B, pitch 256                     ", 128, 128, 256, CU_TENSOR_MAP_SWIZZLE_128B},
        {" 64 rows x 128 B, pitch 128 (= 128 px x 64 B)   ", 128, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B},
        {" 32 rows x 256 B, pitch 256 (no swizzle)        ", 256, 32, 256, CU_TENSOR_MAP_SWIZZLE_NONE},
    };
    for (const Shape& s : shapes) {
        CUtensorMap tm;
        const cuuint64_t gdim[2] = {(cuuint64_t)s.rowb, (cuuint64_t)total_rows * 256 / s.pitch};
        const cuuint64_t gstr[1] = {(cuuint64_t)s.pitch};
        const cuuint32_t box[2] = {(cuuint32_t)s.rowb, (cuuint32_t)s.rows};
        const cuuint32_t es[2] = {1, 1};
        CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, x, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, s.sw,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) { printf("%s encode failed %d\n", s.name, (int)cr); continue; }
        const int bytes = s.rowb * s.rows, loads = 4000, rows_total = (int)gdim[1];
        auto run = [&](auto kern, int depth) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8);
            for (int it = 0; it < 2; ++it) kern<<<148, 128, 16384 * 8>>>(tm, bytes, s.rows, rows_total, loads, out);
            cudaDeviceSynchronize();
            long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
            printf("%s depth %d: %7.1f cycles / load  %6.1f B / cycle / SM\n", s.name, depth, avg / loads, bytes / (avg / loads));
        };
        run(k_tma<1>, 1); run(k_tma<2>, 2); run(k_tma<4>, 4); run(k_tma<8>, 8);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
    }
    return 0;
}
