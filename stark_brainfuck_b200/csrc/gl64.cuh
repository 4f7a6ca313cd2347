// gl64.cuh -- arithmetic in F_p, p = 2^64 - 2^32 + 1, and in F_p[X]/(X^3 - X + 1).
//
// Semantics follow code/algebra.py:89-108 and code/extension_field.py:55-86 of the
// reference: all values are canonical integers in [0, p).  The reference reduces Python
// big ints with `% p`; here the 128-bit product is folded with the identities
// 2^64 = 2^32 - 1 and 2^96 = -1 (mod p), which yields the same canonical value.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GL_HD __host__ __device__ __forceinline__
#else
#define GL_HD inline
#endif

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint8_t u8;

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL  // 2^64 - p = 2^32 - 1
#define GL_HALF 0x7FFFFFFF80000001ULL  // 2^-1 mod p  (code/fri.py:127 `two.inverse()`)

GL_HD u64 gl_add(u64 a, u64 b) {  // a, b < p
    u64 s = a + b;
    // wrapped past 2^64, or landed in [p, 2^64): subtract p, i.e. add 2^64 - p
    return (s < a || s >= GL_P) ? s + GL_EPS : s;
}
GL_HD u64 gl_sub(u64 a, u64 b) {
    u64 d = a - b;
    return (a < b) ? d - GL_EPS : d;  // + p (mod 2^64)
}
GL_HD u64 gl_neg(u64 a) { return a ? GL_P - a : 0; }

GL_HD void gl_mul_wide(u64 a, u64 b, u64 &lo, u64 &hi) {
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umul64hi(a, b);
#else
    unsigned __int128 t = (unsigned __int128)a * b;
    lo = (u64)t;
    hi = (u64)(t >> 64);
#endif
}

// x = hi*2^64 + lo  ->  canonical x mod p
GL_HD u64 gl_reduce128(u64 lo, u64 hi) {
    u64 hh = hi >> 32, hl = hi & GL_EPS;
    u64 t0 = lo - hh;  // hh * 2^96 = -hh
    if (lo < hh) t0 -= GL_EPS;
    u64 t1 = (hl << 32) - hl;  // hl * 2^64 = hl * (2^32 - 1)
    u64 r = t0 + t1;
    if (r < t1) r += GL_EPS;
    return r >= GL_P ? r - GL_P : r;
}
GL_HD u64 gl_mul(u64 a, u64 b) {
    u64 lo, hi;
    gl_mul_wide(a, b, lo, hi);
    return gl_reduce128(lo, hi);
}
GL_HD u64 gl_pow(u64 a, u64 e) {
    u64 acc = 1;
    while (e) {
        if (e & 1) acc = gl_mul(acc, a);
        a = gl_mul(a, a);
        e >>= 1;
    }
    return acc;
}
GL_HD u64 gl_inv(u64 a) { return gl_pow(a, GL_P - 2); }  // inverse(0) == 0 like code/algebra.py:101-103

// a^e from a table of repeated squarings sq[b] = a^(2^b)
GL_HD u64 gl_pow_sq(const u64 *sq, u64 e) {
    u64 acc = 1;
    for (int b = 0; e; ++b, e >>= 1)
        if (e & 1) acc = gl_mul(acc, sq[b]);
    return acc;
}

// ---- cubic extension, X^3 = X - 1 --------------------------------------------------
struct xfe {
    u64 c[3];
};
GL_HD xfe x_add(const xfe &a, const xfe &b) { return {{gl_add(a.c[0], b.c[0]), gl_add(a.c[1], b.c[1]), gl_add(a.c[2], b.c[2])}}; }
GL_HD xfe x_sub(const xfe &a, const xfe &b) { return {{gl_sub(a.c[0], b.c[0]), gl_sub(a.c[1], b.c[1]), gl_sub(a.c[2], b.c[2])}}; }
GL_HD xfe x_mul(const xfe &a, const xfe &b) {
    // schoolbook d0..d4, then X^3 = X - 1, X^4 = X^2 - X  (code/extension_field.py:65-66)
    u64 d0 = gl_mul(a.c[0], b.c[0]);
    u64 d1 = gl_add(gl_mul(a.c[0], b.c[1]), gl_mul(a.c[1], b.c[0]));
    u64 d2 = gl_add(gl_add(gl_mul(a.c[0], b.c[2]), gl_mul(a.c[1], b.c[1])), gl_mul(a.c[2], b.c[0]));
    u64 d3 = gl_add(gl_mul(a.c[1], b.c[2]), gl_mul(a.c[2], b.c[1]));
    u64 d4 = gl_mul(a.c[2], b.c[2]);
    return {{gl_sub(d0, d3), gl_sub(gl_add(d1, d3), d4), gl_add(d2, d4)}};
}
GL_HD xfe x_mul_base(const xfe &a, u64 s) { return {{gl_mul(a.c[0], s), gl_mul(a.c[1], s), gl_mul(a.c[2], s)}}; }
