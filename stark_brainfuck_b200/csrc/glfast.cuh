// glfast.cuh -- device-only Goldilocks arithmetic with lazy reduction in the plain (non-Montgomery)
// domain, for kernels whose twiddles are computed on the fly (dist.cu); the NTT passes use glmont.cuh.
//
// Representation: a u64 in [0, 2^64) standing for its residue mod p = 2^64 - 2^32 + 1
// ("lazy"); a value is "canonical" when it is < p.  2^64 = EPS (mod p), EPS = 2^32 - 1.
//   fadd(a, b): a + b,  needs ONE canonical operand, lazy result      (6 integer instructions)
//   fsub(a, b): a - b,  needs b canonical,           lazy result      (5)
//   fmul(a, b): a * b,  any operands,                lazy result      (4 IMAD.WIDE + ~15)
//   canon(x)  : the canonical representative                          (4-5)
// Carry chains are written in PTX (add.cc/addc/subc) so that a carry costs one IADD3.X
// instead of a compare-and-select; values only leave the kernels through canon().
#pragma once
#include "gl64.cuh"

__device__ __forceinline__ u64 f_pack(u32 lo, u32 hi) {
    u64 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void f_unpack(u64 x, u32 &lo, u32 &hi) {
    asm("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(x));
}

__device__ __forceinline__ u64 fadd(u64 a, u64 b) {
    u32 a0, a1, b0, b1, r0, r1;
    f_unpack(a, a0, a1);
    f_unpack(b, b0, b1);
    // s = a + b; on carry out add EPS (= the all-ones carry mask).  With one operand < p the
    // wrapped sum is < p, so the correction cannot carry again.
    asm("{\n .reg .u32 c, m;\n"
        " add.cc.u32  %0, %2, %4;\n"
        " addc.cc.u32 %1, %3, %5;\n"
        " addc.u32    c, 0, 0;\n"   // carry out as 0/1 (never read an add-chain carry with subc:
        " neg.s32     m, c;\n"      //  ptxas keeps the borrow convention CF = !borrow for sub chains)
        " add.cc.u32  %0, %0, m;\n"
        " addc.u32    %1, %1, 0;\n}"
        : "=&r"(r0), "=&r"(r1)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return f_pack(r0, r1);
}

__device__ __forceinline__ u64 fsub(u64 a, u64 b) {
    u32 a0, a1, b0, b1, r0, r1;
    f_unpack(a, a0, a1);
    f_unpack(b, b0, b1);
    // d = a - b; on borrow subtract EPS (i.e. add p mod 2^64).  b < p keeps the wrapped
    // difference above EPS, so the correction cannot borrow again.
    asm("{\n .reg .u32 m;\n"
        " sub.cc.u32  %0, %2, %4;\n"
        " subc.cc.u32 %1, %3, %5;\n"
        " subc.u32    m, 0, 0;\n"
        " sub.cc.u32  %0, %0, m;\n"
        " subc.u32    %1, %1, 0;\n}"
        : "=&r"(r0), "=&r"(r1)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return f_pack(r0, r1);
}

__device__ __forceinline__ u64 canon(u64 x) {
    u32 x0, x1, r0, r1;
    f_unpack(x, x0, x1);
    // p = 0xFFFFFFFF_00000001:  x >= p  <=>  x1 == 0xFFFFFFFF and x0 != 0,  and then x - p = x0 - 1
    asm("{\n .reg .pred q;\n"
        " setp.eq.u32     q, %3, 0xFFFFFFFF;\n"
        " setp.ne.and.u32 q, %2, 0, q;\n"
        " mov.u32         %0, %2;\n"
        " @q add.u32      %0, %2, 0xFFFFFFFF;\n"
        " selp.u32        %1, 0, %3, q;\n}"
        : "=&r"(r0), "=r"(r1)
        : "r"(x0), "r"(x1));
    return f_pack(r0, r1);
}

__device__ __forceinline__ u64 fmul(u64 a, u64 b) {
    u32 a0, a1, b0, b1;
    f_unpack(a, a0, a1);
    f_unpack(b, b0, b1);
    // 64x64 -> 128 by four 32x32 -> 64 multiply-adds; no partial sum can overflow 64 bits
    const u64 lo = (u64)a0 * b0;
    const u64 m1 = (u64)a0 * b1 + (lo >> 32);
    const u64 m2 = (u64)a1 * b0 + (u32)m1;
    const u64 hi = (u64)a1 * b1 + (m1 >> 32) + (m2 >> 32);
    const u32 x0 = (u32)lo, x1 = (u32)m2, x2 = (u32)hi, x3 = (u32)(hi >> 32);
    u32 r0, r1;
    // x3*2^96 + x2*2^64 + (x1:x0)  =  (x1:x0) - x3 + x2*EPS   (2^96 = -1, 2^64 = EPS)
    asm("{\n .reg .u32 m, t0, t1;\n"
        " sub.cc.u32  %0, %2, %5;\n"
        " subc.cc.u32 %1, %3, 0;\n"
        " subc.u32    m, 0, 0;\n"
        " sub.cc.u32  %0, %0, m;\n"
        " subc.u32    %1, %1, 0;\n"
        " sub.cc.u32  t0, 0, %4;\n"
        " subc.u32    t1, %4, 0;\n"
        " add.cc.u32  %0, %0, t0;\n"
        " addc.cc.u32 %1, %1, t1;\n"
        " addc.u32    t0, 0, 0;\n"
        " neg.s32     m, t0;\n"
        " add.cc.u32  %0, %0, m;\n"
        " addc.u32    %1, %1, 0;\n}"
        : "=&r"(r0), "=&r"(r1)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3));
    return f_pack(r0, r1);
}

// a^e from repeated squarings sq[b] = a^(2^b); lazy result
__device__ __forceinline__ u64 fpow_sq(const u64 *sq, u64 e) {
    u64 acc = 1;
    for (int b = 0; e; ++b, e >>= 1)
        if (e & 1) acc = fmul(acc, sq[b]);
    return acc;
}
