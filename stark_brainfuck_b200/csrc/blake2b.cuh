// blake2b.cuh -- BLAKE2b-512 (RFC 7693), unkeyed, 64-byte digest: what hashlib.blake2b(msg)
// computes in code/merkle.py:31-32 and :38-39.  One hash per thread; the 64-bit state lives
// in registers, rotations map to PRMT (16/24), a register swap (32) and SHF (63).
#pragma once
#include "gl64.cuh"

__device__ __forceinline__ u64 b2b_rotr32(u64 x) { return (x >> 32) | (x << 32); }
__device__ __forceinline__ u64 b2b_rotr24(u64 x) {
    u32 lo = (u32)x, hi = (u32)(x >> 32);
    return ((u64)__byte_perm(lo, hi, 0x2107) << 32) | __byte_perm(lo, hi, 0x6543);
}
__device__ __forceinline__ u64 b2b_rotr16(u64 x) {
    u32 lo = (u32)x, hi = (u32)(x >> 32);
    return ((u64)__byte_perm(lo, hi, 0x1076) << 32) | __byte_perm(lo, hi, 0x5432);
}
__device__ __forceinline__ u64 b2b_rotr63(u64 x) {
    u32 lo = (u32)x, hi = (u32)(x >> 32);
    return ((u64)__funnelshift_l(lo, hi, 1) << 32) | __funnelshift_l(hi, lo, 1);
}

#define B2B_IV0 0x6a09e667f3bcc908ULL
#define B2B_IV1 0xbb67ae8584caa73bULL
#define B2B_IV2 0x3c6ef372fe94f82bULL
#define B2B_IV3 0xa54ff53a5f1d36f1ULL
#define B2B_IV4 0x510e527fade682d1ULL
#define B2B_IV5 0x9b05688c2b3e6c1fULL
#define B2B_IV6 0x1f83d9abfb41bd6bULL
#define B2B_IV7 0x5be0cd19137e2179ULL

__device__ __forceinline__ void b2b_init(u64 h[8]) {
    h[0] = B2B_IV0 ^ 0x01010040ULL;  // digest_length 64, key 0, fanout 1, depth 1
    h[1] = B2B_IV1;
    h[2] = B2B_IV2;
    h[3] = B2B_IV3;
    h[4] = B2B_IV4;
    h[5] = B2B_IV5;
    h[6] = B2B_IV6;
    h[7] = B2B_IV7;
}

#define B2B_G(a, b, c, d, x, y)  \
    do {                         \
        a = a + b + (x);         \
        d = b2b_rotr32(d ^ a);   \
        c = c + d;               \
        b = b2b_rotr24(b ^ c);   \
        a = a + b + (y);         \
        d = b2b_rotr16(d ^ a);   \
        c = c + d;               \
        b = b2b_rotr63(b ^ c);   \
    } while (0)

#define B2B_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
    do {                                                                                \
        B2B_G(v0, v4, v8, v12, m[s0], m[s1]);                                           \
        B2B_G(v1, v5, v9, v13, m[s2], m[s3]);                                           \
        B2B_G(v2, v6, v10, v14, m[s4], m[s5]);                                          \
        B2B_G(v3, v7, v11, v15, m[s6], m[s7]);                                          \
        B2B_G(v0, v5, v10, v15, m[s8], m[s9]);                                          \
        B2B_G(v1, v6, v11, v12, m[s10], m[s11]);                                        \
        B2B_G(v2, v7, v8, v13, m[s12], m[s13]);                                         \
        B2B_G(v3, v4, v9, v14, m[s14], m[s15]);                                         \
    } while (0)

// one compression; t = total bytes hashed so far including this block (< 2^64)
__device__ __forceinline__ void b2b_compress(u64 h[8], const u64 m[16], u64 t, bool last) {
    u64 v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
    u64 v8 = B2B_IV0, v9 = B2B_IV1, v10 = B2B_IV2, v11 = B2B_IV3;
    u64 v12 = B2B_IV4 ^ t, v13 = B2B_IV5, v14 = last ? ~B2B_IV6 : B2B_IV6, v15 = B2B_IV7;
    B2B_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    B2B_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
    B2B_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4);
    B2B_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8);
    B2B_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13);
    B2B_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9);
    B2B_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11);
    B2B_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10);
    B2B_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5);
    B2B_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0);
    B2B_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    B2B_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
    h[0] ^= v0 ^ v8;
    h[1] ^= v1 ^ v9;
    h[2] ^= v2 ^ v10;
    h[3] ^= v3 ^ v11;
    h[4] ^= v4 ^ v12;
    h[5] ^= v5 ^ v13;
    h[6] ^= v6 ^ v14;
    h[7] ^= v7 ^ v15;
}
