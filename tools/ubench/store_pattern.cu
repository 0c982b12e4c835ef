// Micro-benchmark: achieved HBM write bandwidth of the epilogue's store patterns (one 16-byte piece per lane).
//   G = lanes that cooperate on one row: G=1 -> every lane writes 16 B of a DIFFERENT row (row pitch P bytes),
//   G=2 -> lane pairs write 32 contiguous bytes of a row, ... G=8 -> 128 contiguous bytes per row.
// Usage: store_pattern   (prints GB/s per pattern).  Build: nvcc -arch=sm_100a -O3 -o store_pattern store_pattern.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int G>
__global__ void __launch_bounds__(256) k_store(uint8_t* __restrict__ y, size_t rows, int pitch, int pieces_per_row) {
    // a warp handles 32/G rows x (pieces_per_row) pieces; instruction j writes pieces [j*G, j*G+G) of 32/G rows
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    constexpr int RPW = 32 / G;                    // rows per warp instruction
    const uint4 v = make_uint4(lane, 2, 3, 4);
    for (size_t r0 = warp * 32; r0 < rows; r0 += nwarps * 32) {        // 32 rows per warp tile (like the epilogue slab)
        for (int rr = 0; rr < 32; rr += RPW) {
            const size_t row = r0 + rr + lane / G;
            for (int p = 0; p < pieces_per_row; p += G) {
                const int piece = p + lane % G;
                if (row < rows) *reinterpret_cast<uint4*>(y + row * (size_t)pitch + (size_t)piece * 16) = v;
            }
        }
    }
}

int main() {
    const size_t rows = 802816;                    // 256 x 56 x 56 pixels
    for (int pitch : {256, 512}) {                 // bytes per row (256 ch codes / 256 ch f16)
        uint8_t* y; cudaMalloc(&y, rows * (size_t)pitch);
        const int ppr = pitch / 16 / 2;            // a warp covers half a row (its 128-column slab)
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        auto run = [&](auto kern, const char* name) {
            for (int it = 0; it < 3; ++it) kern<<<148 * 8, 256>>>(y, rows, pitch, ppr);
            cudaEventRecord(a);
            for (int it = 0; it < 10; ++it) kern<<<148 * 8, 256>>>(y, rows, pitch, ppr);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            printf("pitch %d  %-6s %8.1f GB/s (half rows written: %zu MB)\n", pitch, name, rows * (size_t)pitch / 2 / (ms / 10) / 1e6,
                   rows * (size_t)pitch / 2 >> 20);
        };
        run(k_store<1>, "G=1"); run(k_store<2>, "G=2"); run(k_store<4>, "G=4"); run(k_store<8>, "G=8");
        cudaFree(y);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
