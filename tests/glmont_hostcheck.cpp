// Host check of the Montgomery / lazy arithmetic the quotient, combination and point-evaluation kernels are
// built from (stark_brainfuck_b200/csrc/glmont.cuh): the same functions compile for the host, so their
// contracts are checked here against plain 128-bit arithmetic without a GPU.
//   g++ -O2 -std=c++17 -I stark_brainfuck_b200/csrc tests/glmont_hostcheck.cpp -o glmont_hostcheck
#include <stdio.h>
#include <stdlib.h>

#include "glmont.cuh"

typedef unsigned __int128 u128;

static u64 rng_state = 0x2545F4914F6CDD1DULL;
static u64 rnd() {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return rng_state;
}
static u64 any_u64() {  // lazy operands: any 64-bit pattern, edge values over-represented
    switch (rnd() & 7) {
        case 0: return GL_P - 1 - (rnd() & 3);
        case 1: return GL_P + (rnd() & 0xFFFFFFFF) % 0xFFFFFFFFULL;  // in [p, 2^64)
        case 2: return rnd() & 3;
        case 3: return ~(u64)0 - (rnd() & 3);
        default: return rnd();
    }
}
static u64 canonical() { return any_u64() % GL_P; }
static u64 mulmod(u64 a, u64 b) { return (u64)((u128)(a % GL_P) * (b % GL_P) % GL_P); }
static u64 addmod(u64 a, u64 b) { return (u64)(((u128)(a % GL_P) + (b % GL_P)) % GL_P); }
static u64 submod(u64 a, u64 b) { return (u64)(((u128)(a % GL_P) + GL_P - (b % GL_P)) % GL_P); }

struct X3 {
    u64 c[3];
};
static X3 xmul_ref(const X3 &a, const X3 &b) {  // X^3 = X - 1 (code/extension_field.py:65-66)
    u64 d[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) d[i + j] = addmod(d[i + j], mulmod(a.c[i], b.c[j]));
    return {{submod(d[0], d[3]), submod(addmod(d[1], d[3]), d[4]), addmod(d[2], d[4])}};
}

int main() {
    long fails = 0;
    const u64 R = GL_EPS;  // 2^64 mod p
    const u64 Rinv = gl_inv(R);
    for (int it = 0; it < 2000000; ++it) {
        const u64 a = any_u64(), b = canonical();
        // mont_mul(a, w * 2^64) = a * w for ANY u64 a, result <= p
        const u64 r = mont_mul(a, gl_to_mont(b));
        if (r > GL_P || r % GL_P != mulmod(a, b)) ++fails;
        // with a plain second operand the product loses one factor 2^64 (the quotient kernel's bookkeeping)
        const u64 q = mont_mul(a, b);
        if (q > GL_P || q % GL_P != mulmod(mulmod(a, b), Rinv)) ++fails;
        // lazy add / subtract: first operand any u64, second <= p
        if (ladd(a, b) % GL_P != addmod(a, b) || lsub(a, b) % GL_P != submod(a, b)) ++fails;
        if (ladd(a, GL_P) % GL_P != a % GL_P || lsub(a, GL_P) % GL_P != a % GL_P) ++fails;
        if (lcanon(a) != a % GL_P) ++fails;
    }
    for (int it = 0; it < 300000; ++it) {
        xfe a, bm, acc;
        X3 ar, br, accr;
        for (int j = 0; j < 3; ++j) {
            a.c[j] = any_u64();
            ar.c[j] = a.c[j] % GL_P;
            br.c[j] = canonical();
            bm.c[j] = gl_to_mont(br.c[j]);
            acc.c[j] = any_u64();
            accr.c[j] = acc.c[j] % GL_P;
        }
        const X3 want = xmul_ref(ar, br);
        const xfe got = x_mul_mont(a, bm);
        x_fma_mont(acc, a, bm);
        for (int j = 0; j < 3; ++j) {
            if (got.c[j] % GL_P != want.c[j]) ++fails;
            if (acc.c[j] % GL_P != addmod(accr.c[j], want.c[j])) ++fails;
        }
        // e multiplications by a plain value == one multiplication by the cached power c_e (quotient.cu):
        // c_1 = x, c_(k+1) = lcanon(c_k * x * 2^-64)
        xfe x, c = {}, chain = a;
        for (int j = 0; j < 3; ++j) x.c[j] = canonical();
        const int e = 1 + (int)(rnd() % 8);
        c = x;
        for (int k = 1; k < e; ++k) {
            c = x_mul_mont(c, x);
            for (int j = 0; j < 3; ++j) c.c[j] = lcanon(c.c[j]);
        }
        for (int k = 0; k < e; ++k) chain = x_mul_mont(chain, x);
        const xfe once = x_mul_mont(a, c);
        for (int j = 0; j < 3; ++j)
            if (once.c[j] % GL_P != chain.c[j] % GL_P) ++fails;
    }
    printf("%ld failed\n", fails);
    return fails != 0;
}
