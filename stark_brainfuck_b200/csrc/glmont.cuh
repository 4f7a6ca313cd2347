// glmont.cuh -- Goldilocks arithmetic with Montgomery-form multipliers (NTT passes of ntt4.cuh, quotient.cu,
// combine.cu, misc.cu, dist.cu).
//
// Values are u64 residues mod p = 2^64 - 2^32 + 1 in one of two states:
//   "lazy"       any u64 (stands for its residue)
//   "canonical"  <= p   (weakly canonical; the library's outputs go through canon() -> < p)
// Twiddles are kept in Montgomery form wm = w * 2^64 mod p (< p), so that
//   mont_mul(a, wm) = a * w mod p     for ANY u64 a, result canonical (< p)
// with a multiplication-free reduction (p^-1 = 2^32 + 1 mod 2^64):
//   x = a * wm = xh * 2^64 + xl,  a' = xl * (2^32 + 1) mod 2^64,  b = hi64(a' * p),  r = xh - b (+ p on borrow)
// which costs 9 carry-chain instructions after the 64x64 -> 128 multiply (4 IMAD.WIDE), and
// whose canonical result lets the butterfly's add/sub run without a separate reduction:
//   ladd(a, b): a lazy, b canonical -> lazy      lsub(a, b): a lazy, b canonical -> lazy
// The functions compile for the host as well (plain C) so that the pass logic can be checked
// on a machine without a GPU (tests/ntt4_hostcheck.cpp, tests/glmont_hostcheck.cpp).
#pragma once
#include "gl64.cuh"

GL_HD u64 gl_to_mont(u64 w) { return gl_mul(w, GL_EPS); }  // 2^64 = EPS (mod p)

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ u64 m_pack(u32 lo, u32 hi) {
    u64 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void m_unpack(u64 x, u32 &lo, u32 &hi) {
    asm("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(x));
}
#endif

GL_HD u64 mont_mul(u64 a, u64 wm) {
#if defined(__CUDA_ARCH__)
    const u64 xl = a * wm;
    const u64 xh = __umul64hi(a, wm);
    u32 x0, x1, x2, x3, r0, r1;
    m_unpack(xl, x0, x1);
    m_unpack(xh, x2, x3);
    // a' = (a1 : x0), a1 = x1 + x0 (carry e);  b = a' - (a' >> 32) - e = a' - (a1 + e);  r = xh - b
    asm("{\n .reg .u32 a1, t, b0, b1, m;\n"
        " add.cc.u32  a1, %3, %2;\n"
        " addc.u32    t, a1, 0;\n"   // a1 + e never wraps: e = 1 implies a1 <= 2^32 - 2
        " sub.cc.u32  b0, %2, t;\n"
        " subc.u32    b1, a1, 0;\n"
        " sub.cc.u32  %0, %4, b0;\n"
        " subc.cc.u32 %1, %5, b1;\n"
        " subc.u32    m, 0, 0;\n"    // 0xFFFFFFFF (= EPS) on borrow
        " sub.cc.u32  %0, %0, m;\n"  // r + 2^64 - EPS = r + p
        " subc.u32    %1, %1, 0;\n}"
        : "=&r"(r0), "=&r"(r1)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3));
    return m_pack(r0, r1);
#else
    const unsigned __int128 x = (unsigned __int128)a * wm;
    const u64 xl = (u64)x, xh = (u64)(x >> 64);
    const u64 a2 = xl + (xl << 32);
    const u64 e = a2 < xl;
    const u64 b = a2 - (a2 >> 32) - e;
    u64 r = xh - b;
    if (xh < b) r -= GL_EPS;
    return r;
#endif
}

// a + b: a lazy, b <= p; lazy result
GL_HD u64 ladd(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
    u32 a0, a1, b0, b1, r0, r1;
    m_unpack(a, a0, a1);
    m_unpack(b, b0, b1);
    asm("{\n .reg .u32 c, m;\n"
        " add.cc.u32  %0, %2, %4;\n"
        " addc.cc.u32 %1, %3, %5;\n"
        " addc.u32    c, 0, 0;\n"
        " neg.s32     m, c;\n"       // EPS on carry: (a + b - 2^64) + EPS = a + b - p < 2^64
        " add.cc.u32  %0, %0, m;\n"
        " addc.u32    %1, %1, 0;\n}"
        : "=&r"(r0), "=&r"(r1)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return m_pack(r0, r1);
#else
    u64 s = a + b;
    if (s < a) s += GL_EPS;
    return s;
#endif
}

// a - b: a lazy, b <= p; lazy result
GL_HD u64 lsub(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
    u32 a0, a1, b0, b1, r0, r1;
    m_unpack(a, a0, a1);
    m_unpack(b, b0, b1);
    asm("{\n .reg .u32 m;\n"
        " sub.cc.u32  %0, %2, %4;\n"
        " subc.cc.u32 %1, %3, %5;\n"
        " subc.u32    m, 0, 0;\n"
        " sub.cc.u32  %0, %0, m;\n"  // (a - b + 2^64) - EPS = a - b + p >= 0
        " subc.u32    %1, %1, 0;\n}"
        : "=&r"(r0), "=&r"(r1)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return m_pack(r0, r1);
#else
    u64 d = a - b;
    if (a < b) d -= GL_EPS;
    return d;
#endif
}

// the representative in [0, p)
GL_HD u64 lcanon(u64 x) { return x >= GL_P ? x - GL_P : x; }

// ---- cubic extension with Montgomery multiplications (quotient.cu, combine.cu) ----------------------
// a * b * 2^-64 in F_p[X]/(X^3 - X + 1): a has lazy coefficients, b canonical ones; lazy result.
// With b = (value * 2^64) this is the plain product; with a plain b the factor 2^-64 is compensated
// by the caller (quotient.cu pre-scales each monomial's coefficient by 2^(64 * degree)).
// X^3 = X - 1, X^4 = X^2 - X  (code/extension_field.py:65-66):
//   c0 = a0b0 - (a1b2 + a2b1),  c1 = a0b1 + a1b0 + (a1b2 + a2b1) - a2b2,  c2 = a0b2 + a1b1 + a2b0 + a2b2
GL_HD xfe x_mul_mont(const xfe &a, const xfe &b) {
    const u64 m12 = mont_mul(a.c[1], b.c[2]), m21 = mont_mul(a.c[2], b.c[1]), m22 = mont_mul(a.c[2], b.c[2]);
    xfe r;
    r.c[0] = lsub(lsub(mont_mul(a.c[0], b.c[0]), m12), m21);
    r.c[1] = lsub(ladd(ladd(ladd(mont_mul(a.c[0], b.c[1]), mont_mul(a.c[1], b.c[0])), m12), m21), m22);
    r.c[2] = ladd(ladd(ladd(mont_mul(a.c[0], b.c[2]), mont_mul(a.c[1], b.c[1])), mont_mul(a.c[2], b.c[0])), m22);
    return r;
}
// acc (lazy) + a * b * 2^-64, same operand states
GL_HD void x_fma_mont(xfe &acc, const xfe &a, const xfe &b) {
    const u64 m12 = mont_mul(a.c[1], b.c[2]), m21 = mont_mul(a.c[2], b.c[1]), m22 = mont_mul(a.c[2], b.c[2]);
    acc.c[0] = lsub(lsub(ladd(acc.c[0], mont_mul(a.c[0], b.c[0])), m12), m21);
    acc.c[1] = lsub(ladd(ladd(ladd(ladd(acc.c[1], mont_mul(a.c[0], b.c[1])), mont_mul(a.c[1], b.c[0])), m12), m21), m22);
    acc.c[2] = ladd(ladd(ladd(ladd(acc.c[2], mont_mul(a.c[0], b.c[2])), mont_mul(a.c[1], b.c[1])), mont_mul(a.c[2], b.c[0])),
                    m22);
}
