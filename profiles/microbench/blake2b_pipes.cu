// blake2b_pipes.cu -- can BLAKE2b's 64-bit additions be moved from the ALU pipe (IADD3) to the
// FMA-heavy pipe (IMAD.WIDE with a multiplier of 1 that ptxas cannot fold)?  Each thread runs
// ITER chained compressions on register state; variants differ in how the adds are written.
//   V0: plain 64-bit C adds                    (all ALU: IADD3 / IADD3.X)
//   V1: c = c + d via mad.wide                 (2 of the 4 adds per G on the FMA pipe)
//   V2: V1 and a = (a + b) via mad.wide, + m on the ALU pipe
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o blake2b_pipes blake2b_pipes.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
typedef uint64_t u64;
typedef uint32_t u32;

__device__ __forceinline__ u64 rotr32(u64 x) { return (x >> 32) | (x << 32); }
__device__ __forceinline__ u64 rotr24(u64 x) {
    u32 lo = (u32)x, hi = (u32)(x >> 32);
    return ((u64)__byte_perm(lo, hi, 0x2107) << 32) | __byte_perm(lo, hi, 0x6543);
}
__device__ __forceinline__ u64 rotr16(u64 x) {
    u32 lo = (u32)x, hi = (u32)(x >> 32);
    return ((u64)__byte_perm(lo, hi, 0x1076) << 32) | __byte_perm(lo, hi, 0x5432);
}
__device__ __forceinline__ u64 rotr63(u64 x) {
    u32 lo = (u32)x, hi = (u32)(x >> 32);
    return ((u64)__funnelshift_l(lo, hi, 1) << 32) | __funnelshift_l(hi, lo, 1);
}
// a + b (mod 2^64) on the FMA pipe: (b + a_lo * one) as a 64-bit mad, then hi += a_hi * one
__device__ __forceinline__ u64 add_fma(u64 a, u64 b, u32 one) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32);
    u64 t;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"(a0), "r"(one), "l"(b));
    u32 t0 = (u32)t, t1 = (u32)(t >> 32);
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(t1) : "r"(a1), "r"(one));
    return ((u64)t1 << 32) | t0;
}

template <int V>
__device__ __forceinline__ void G(u64 &a, u64 &b, u64 &c, u64 &d, u64 x, u64 y, u32 one) {
    if (V >= 2) a = add_fma(a, b, one) + x; else a = a + b + x;
    d = rotr32(d ^ a);
    if (V >= 1) c = add_fma(d, c, one); else c = c + d;
    b = rotr24(b ^ c);
    a = a + b + y;
    d = rotr16(d ^ a);
    if (V >= 1) c = add_fma(d, c, one); else c = c + d;
    b = rotr63(b ^ c);
}

__constant__ unsigned char SIGMA[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};

template <int V>
__device__ __forceinline__ void compress(u64 h[8], const u64 m[16], u64 t, u32 one) {
    constexpr unsigned char S[12][16] = {
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
        {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
        {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
        {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
        {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
    u64 v[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = h[i];
    v[8] = 0x6a09e667f3bcc908ULL; v[9] = 0xbb67ae8584caa73bULL; v[10] = 0x3c6ef372fe94f82bULL; v[11] = 0xa54ff53a5f1d36f1ULL;
    v[12] = 0x510e527fade682d1ULL ^ t; v[13] = 0x9b05688c2b3e6c1fULL; v[14] = ~0x1f83d9abfb41bd6bULL; v[15] = 0x5be0cd19137e2179ULL;
#pragma unroll
    for (int r = 0; r < 12; ++r) {
        G<V>(v[0], v[4], v[8], v[12], m[S[r][0]], m[S[r][1]], one);
        G<V>(v[1], v[5], v[9], v[13], m[S[r][2]], m[S[r][3]], one);
        G<V>(v[2], v[6], v[10], v[14], m[S[r][4]], m[S[r][5]], one);
        G<V>(v[3], v[7], v[11], v[15], m[S[r][6]], m[S[r][7]], one);
        G<V>(v[0], v[5], v[10], v[15], m[S[r][8]], m[S[r][9]], one);
        G<V>(v[1], v[6], v[11], v[12], m[S[r][10]], m[S[r][11]], one);
        G<V>(v[2], v[7], v[8], v[13], m[S[r][12]], m[S[r][13]], one);
        G<V>(v[3], v[4], v[9], v[14], m[S[r][14]], m[S[r][15]], one);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[i + 8];
}

template <int V>
__global__ void __launch_bounds__(128) bench(u64 *out, u32 one, int iters) {
    u64 h[8], m[16];
    const u64 id = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = id * 0x9E3779B97F4A7C15ULL + i;
#pragma unroll
    for (int i = 0; i < 16; ++i) m[i] = id + 0x0123456789ABCDEFULL * (i + 1);
    for (int it = 0; it < iters; ++it) {
        compress<V>(h, m, 128 + it, one);
        m[it & 15] ^= h[0];
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= h[i];
    out[id] = s;
}

template <int V>
void run(u64 *d, u64 *hres) {
    const int grid = 148 * 8, iters = 64;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    bench<V><<<grid, 128>>>(d, 1, iters);
    cudaEventRecord(e0);
    bench<V><<<grid, 128>>>(d, 1, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(hres, d, 8 * 16, cudaMemcpyDeviceToHost);
    printf("V%d: %.3f ms, %.2f G compressions/s, check %016llx\n", V, ms, (double)grid * 128 * iters / ms / 1e6,
           (unsigned long long)(hres[0] ^ hres[5]));
}

int main() {
    u64 *d, h[16];
    cudaMalloc(&d, 148 * 8 * 128 * 8);
    run<0>(d, h);
    run<1>(d, h);
    run<2>(d, h);
    return 0;
}
