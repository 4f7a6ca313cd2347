// icache.cu -- how fast does an SM stream straight-line code that every warp executes ONCE?
// (the situation of a fully unrolled NTT pass on a single 2^20 vector: ~7-14 warps per SM).
// Kernels of N independent integer instructions, 8 registers in flight; timed per launch, same
// kernel repeated (warm instruction cache) and alternating with a second kernel of the same size.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache icache.cu && ./icache
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

template <int N, int SALT>
__global__ void __launch_bounds__(128) body(uint32_t *out, uint32_t c) {
    uint32_t x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i + SALT;
#pragma unroll
    for (int i = 0; i < N; ++i) x[i & 7] = x[i & 7] * c + x[(i + 3) & 7] + (i * 2654435761u + SALT);
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int N>
void run(uint32_t *d, int ctas_per_sm) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = 148 * ctas_per_sm;
    float same = 0, alt = 0;
    for (int rep = 0; rep < 3; ++rep) {
        body<N, 0><<<grid, 128>>>(d, 3);
        body<N, 1><<<grid, 128>>>(d, 3);
    }
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) body<N, 0><<<grid, 128>>>(d, 3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&same, e0, e1);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) {
        body<N, 0><<<grid, 128>>>(d, 3);
        body<N, 1><<<grid, 128>>>(d, 3);
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&alt, e0, e1);
    printf("N=%6d instr (%4d KB)  %d CTA/SM x 4 warps: same kernel %.2f us/launch, alternating %.2f us/launch\n", N,
           N * 16 / 1024, ctas_per_sm, same * 1000 / 20, alt * 1000 / 20);
}

int main() {
    uint32_t *d;
    cudaMalloc(&d, 148 * 8 * 128 * 4);
    for (int c : {1, 2, 4}) {
        run<256>(d, c);
        run<512>(d, c);
        run<1024>(d, c);
        run<2048>(d, c);
        run<4096>(d, c);
        run<8192>(d, c);
    }
    return 0;
}
