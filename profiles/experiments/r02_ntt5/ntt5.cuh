// ntt5.cuh -- two-pass transform for plain ntt / intt of 2^16 .. 2^22 points (code/ntt.py:4-42), the shape of
// BASELINE config 2 (one 2^20 base-field vector).  Same arithmetic as ntt4.cuh (Montgomery twiddles, lazy
// add/sub), different schedule, built around what profiles/r01_final_ntt4.summary.txt showed for a single
// vector: 256 CTAs on 148 SMs (1.73 waves), 43 lane-instructions per butterfly of which 16 are index arithmetic
// of the generic pass, and a load -> compute -> store sequence per CTA that nothing overlaps.
//
//   n = R1 * R2,  x[r C + c]  (r < R1 rows, c < C = R2 columns)
//   pass 1   column transforms of length R1, times w^(c k1), in place in the work buffer
//   pass 2   row transforms of length R2 over the contiguous c, output X[k1 + R1 k2]
// A CTA owns Tc ADJACENT columns with Tc = ceil(C / #SM): 2^20 -> 147 CTAs of 7 columns, one wave, every SM
// busy for the whole pass.  R/16 threads work on one column and synchronise among themselves only (named
// barrier per column), so the columns of a CTA drift apart and fill each other's stalls.
//   R = 16 * B,  B = 16 * Q,  Q = R / 256 in {1, 2, 4, 8}:
//   step 1   thread b:         16-point DFT over a  (r = B a + b)        * w_R^(b k1)
//   step 2   thread (k1, b0):  16-point DFT over b1 (b = Q b1 + b0)      * w_B^(b0 k2)
//   step 3   thread (k1, q):   Q-point DFT over b0 for k2 = q + Q j      -> X[k1 + 16 k2 + 256 m]
// All shapes are template parameters: strides and digit positions fold into immediates.  Steps 1 and 2 work in
// place in shared memory (a thread overwrites exactly what it read).  The column is stored as 16 rows of
// B + Q elements: the Q pad elements make the 64-bit accesses of all three steps conflict-free per half warp.
// Global traffic is done in TILE ORDER (lanes along the Tc adjacent columns: 8 Tc contiguous bytes per row)
// through shared memory by all 8 B threads of the CTA (n5_tile_thread: four instructions per element): pass 1 loads
// with cp.async, both passes store their tile the same way; pass 2 reads its contiguous rows straight into
// registers.
// The inter-pass twiddle w^(c k) is a product of two table entries (W^i, i < 1024, and W^(1024 i)); the
// second table carries n^-1 for the inverse transform, which therefore costs nothing extra.
// Phase functions are __host__ __device__: tests/ntt5_hostcheck.cpp runs them thread by thread on the CPU.
#pragma once
#include "ntt4.cuh"

constexpr int N5_SLOTS = 8;  // column slots per CTA (shared-memory layout); Tc <= N5_SLOTS

template <int LOG_R>
struct N5 {
    static constexpr int R = 1 << LOG_R, B = R / 16, LOG_B = LOG_R - 4, T3 = LOG_R - 8, Q = 1 << T3;
    static constexpr int ROW = B + Q;       // padded length of one k1 row
    static constexpr int COL = 16 * ROW;    // elements of one column
    static constexpr int CS = COL + 2;      // column stride: = 2 (mod 16), tile-order accesses conflict-free
    static constexpr size_t smem_elems() { return (size_t)N5_SLOTS * CS + R + B; }  // tile | tw1 | tw2
};

struct Pass5Params {
    const u64 *in;
    u64 *out;
    u64 in_plane_stride, out_plane_stride;
    u32 C, log_C;      // columns of this pass; output index = row * C + col in both passes
    u32 Tc;            // columns per CTA
    u32 last;          // 0: pass 1 (input row * C + col, inter-pass twiddle); 1: pass 2 (input col * R + row, canonical output)
    const u64 *tw1;    // [k1][b]  w_R^(k1 b)    R entries
    const u64 *tw2;    // [k2][b0] w_B^(k2 b0)   B entries
    const u64 *tw_lo;  // W^i, i < 1024                       (pass 1)
    const u64 *tw_hi;  // W^(1024 i) (* n^-1 for the inverse)  (pass 1)
    u64 w16[8];        // w_16^e, e < 8
    // all table entries in Montgomery form
};

template <int LOG_R>
GL_HD u32 n5_pos(u32 row) {  // shared-memory position of row `row` of a column
    return row + (row >> N5<LOG_R>::LOG_B) * N5<LOG_R>::Q;
}

// ---- step 1: thread b of column slot `col`; v = rows B a + b, a < 16 (bit-reversed for the DIT network) ----
template <int LOG_R>
GL_HD void n5_step1_compute(u64 (&v)[16], const u32 b, const u64 *tw1, const u64 *w16) {
    using N = N5<LOG_R>;
    dft4_dit<4>(v, w16, 1);
    v[0] = canon4(v[0]);
#pragma unroll
    for (int k = 1; k < 16; ++k) v[k] = mont_mul(v[k], tw1[(u32)k * N::B + b]);
}
template <int LOG_R>
GL_HD void n5_step1_load_smem(u64 (&v)[16], const u64 *Sc, const u32 b) {
    using N = N5<LOG_R>;
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = Sc[(u32)bitrev4_c(j, 4) * N::ROW + b];
}
template <int LOG_R>
GL_HD void n5_step1_store(const u64 (&v)[16], u64 *Sc, const u32 b) {
    using N = N5<LOG_R>;
#pragma unroll
    for (int k = 0; k < 16; ++k) Sc[(u32)k * N::ROW + b] = v[k];
}

// ---- step 2: thread t = k1 * Q + b0, in place ------------------------------------------------------------------
template <int LOG_R>
GL_HD void n5_step2(u64 *Sc, const u32 t, const u64 *tw2, const u64 *w16) {
    using N = N5<LOG_R>;
    const u32 k1 = t >> N::T3, b0 = t & (N::Q - 1);
    u64 *base = Sc + k1 * N::ROW + b0;
    u64 v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = base[(u32)bitrev4_c(j, 4) * N::Q];
    dft4_dit<4>(v, w16, 1);
    if (N::Q > 1) {
        v[0] = canon4(v[0]);
#pragma unroll
        for (int k = 1; k < 16; ++k) v[k] = mont_mul(v[k], tw2[(u32)k * N::Q + b0]);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) base[(u32)k * N::Q] = v[k];
}

// ---- step 3: thread t = k1 + 16 q reads its 16 / Q groups of Q values, Q-point DFTs, output scaling ----------------
// v[j * Q + m] = X[k1 + 16 (q + Q j) + 256 m]
template <int LOG_R>
GL_HD void n5_step3_load(u64 (&v)[16], const u64 *Sc, const u32 t) {
    using N = N5<LOG_R>;
    const u32 k1 = t & 15, q = t >> 4;
#pragma unroll
    for (int j = 0; j < 16 / N::Q; ++j)
#pragma unroll
        for (int m = 0; m < N::Q; ++m)
            v[j * N::Q + m] = Sc[k1 * N::ROW + N::Q * (q + N::Q * j) + (u32)bitrev4_c(m, N::T3)];
}
template <int LOG_R>
GL_HD void n5_step3_compute_store(u64 (&v)[16], u64 *Sc, const u32 t, const Pass5Params &P, const u32 colg) {
    using N = N5<LOG_R>;
    const u32 k1 = t & 15, q = t >> 4;
#pragma unroll
    for (int j = 0; j < 16 / N::Q; ++j) {
        if (N::Q > 1) {
            u64 g[N::Q];
#pragma unroll
            for (int m = 0; m < N::Q; ++m) g[m] = v[j * N::Q + m];
            dft4_dit<N::T3>(g, P.w16, 16 >> N::T3);
#pragma unroll
            for (int m = 0; m < N::Q; ++m) v[j * N::Q + m] = g[m];
        }
#pragma unroll
        for (int m = 0; m < N::Q; ++m) {
            const u32 k = k1 + 16 * (q + N::Q * j) + 256 * m;
            u64 x = v[j * N::Q + m];
            if (!P.last) {
                const u32 e = colg * k;  // < n <= 2^22
                x = mont_mul(x, mont_mul(P.tw_lo[e & 1023], P.tw_hi[e >> 10]));
            } else {
                x = canon4(x);
            }
            Sc[n5_pos<LOG_R>(k)] = x;
        }
    }
}

// ---- tile-order global <-> shared ----------------------------------------------------------------------------
// The CTA has 8 B threads.  Thread tid moves column slot tid & 7 of the rows (tid >> 3) + B i, i < 16: a warp touches
// 4 adjacent rows x 8 adjacent columns (8 Tc contiguous bytes per row), and because (tid >> 3) < B the shared-memory
// position of row (tid >> 3) + B i is i * ROW + (tid >> 3) -- both addresses advance by a constant per step.
template <int LOG_R>
GL_HD void n5_tile_thread(const u32 tid, u32 &slot, u32 &row0) {
    slot = tid & (N5_SLOTS - 1);
    row0 = tid >> 3;
}
