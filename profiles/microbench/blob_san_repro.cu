// Repro of a miscompiled conditional byte load (nvcc 12.9, -O3, sm_100a): one thread hashes a 427-byte message the way
// round 1's merkle_blob_leaf_kernel did and prints the message words of every block.  Expected digest word 0 for
// this message: 6231b0b06104202a (hashlib.blake2b); word 15 of the last block must be 0.  Observed on a B200:
// m15 = 01d8000001f00000 (bits of an address register), digest 16c23a6ec58ad9eb.  merkle.cu now reads the tail of
// the last block under ordinary branches.
#include <stdio.h>
#include <string.h>
#include "../../stark_brainfuck_b200/csrc/blake2b.cuh"
__global__ void k(const u8 *bytes, const u64 *off, u64 *out) {
    const u8 *msg = bytes + off[0];
    const u64 len = off[1] - off[0];
    u64 h[8];
    b2b_init(h);
    const u64 nblocks = len == 0 ? 1 : (len + 127) >> 7;
#pragma unroll 1
    for (u64 blk = 0; blk < nblocks; ++blk) {
        u64 m[16];
#pragma unroll
        for (int w = 0; w < 16; ++w) {
            u64 v = 0;
            for (int b = 7; b >= 0; --b) {
                const u64 q = blk * 128 + w * 8 + b;
                v = (v << 8) | (q < len ? msg[q] : 0);
            }
            m[w] = v;
        }
        const bool last = blk + 1 == nblocks;
        printf("blk %llu m0 %016llx m15 %016llx t %llu last %d\n", blk, m[0], m[15], last ? len : (blk + 1) * 128, (int)last);
        b2b_compress(h, m, last ? len : (blk + 1) * 128, last);
        printf("   h0 %016llx h7 %016llx\n", h[0], h[7]);
    }
    for (int i = 0; i < 8; ++i) out[i] = h[i];
}
int main() {
    const int L = 427;
    unsigned char hb[L + 8];
    for (int i = 0; i < L + 8; ++i) hb[i] = (unsigned char)(i * 7 + 3);
    unsigned long long ho[2] = {0, L};
    u8 *db; u64 *doff, *dout;
    cudaMalloc(&db, L + 8); cudaMalloc(&doff, 16); cudaMalloc(&dout, 64);
    cudaMemcpy(db, hb, L + 8, cudaMemcpyHostToDevice); cudaMemcpy(doff, ho, 16, cudaMemcpyHostToDevice);
    k<<<1, 1>>>(db, doff, dout);
    unsigned long long r[8];
    cudaMemcpy(r, dout, 64, cudaMemcpyDeviceToHost);
    printf("digest0 %016llx\n", r[0]);
    return 0;
}
