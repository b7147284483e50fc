// Microbenchmark: sustained rate of the legacy warp-level tensor path (mma.sync.m16n8k16 bf16 -> fp32) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_rate hmma_rate.cu && ./hmma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
    uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = 5, b1 = 11;
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0.f;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    for (int wpb : {4, 8}) {
        const int blocks = 148 * 8 / (wpb / 4 > 0 ? 1 : 1), iters = 20000;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<<<blocks, wpb * 32>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); k<<<blocks, wpb * 32>>>(out, iters); cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 16 * 8 * 16 * 8.0 * iters * (double)blocks * wpb;
        printf("warps/block %d blocks %d: %.3f ms  %.1f TFLOP/s dense bf16 (mma.sync m16n8k16)\n", wpb, blocks, ms, flops / ms / 1e9);
    }
    return 0;
}
