// ntt4.cuh -- one PASS of the batched NTT (replaces the radix-2 recursion of code/ntt.py:4-23 for
// one digit of the index; the plan over passes lives in ntt4_plan.h / ntt.cu).
//
// A pass computes, for every column of its tile, the R-point transform over the rows
//     X[k] = sum_r x[r] * w_R^(r k),      R = 2^t * E^a,  E = 16 (a <= 2) or 8 (a <= 3),  t < log2 E
// Design constraint that shaped it (profiles/microbench/icache.cu): an SM streams
// code that each warp executes once at ~1.5 B/clk, so a fully unrolled pass (100 KB of SASS) is
// bound by instruction fetch.  Here the executed code is three small loop bodies (~25 KB):
//   tail   global -> registers -> 2^t-point transform over the top digit -> * w_R^(low*k) ->
//          shared memory               (fused: zero padding, row part of the coset scale)
//   core   `a` times: shared -> registers -> E-point transform (DIT, twiddles from the
//          kernel-parameter constant bank) -> * w^(lo*k) -> shared, in place
//   out    shared -> * inter-pass twiddle w^(col*k) (four interleaved running products per
//          thread) | * n^-1 offset^-k | canonicalise -> global
// E = 16 keeps the instruction count lowest (large batches); E = 8 doubles the number of threads
// per tile, which is what a single vector needs to fill 148 SMs (small batches).
// The tile lives in shared memory as [col][row] with one pad element per E rows and a column
// stride = 4 (mod 16) elements, which makes every 64-bit access pattern of the three phases
// conflict-free (DESIGN.md has the lane -> bank tables).
// All multiplications are Montgomery multiplications by canonical table values (glmont.cuh).
// The phase functions are __host__ __device__: tests/ntt4_hostcheck.cpp runs them thread by
// thread on the CPU against the oracle.
#pragma once
#include "glmont.cuh"

enum : u32 {
    P4_FIRST = 1,    // bounds check against n_in (zero padding)
    P4_OUT_MUL = 2,  // last pass: multiply the output by out_mul (n^-1)
    P4_LAST = 4,     // last pass of the plan: rows are the contiguous input dimension
    P4_INPUT = 8,    // first pass of the plan: reads the caller's vector
};

struct Pass4Params {
    const u64 *in;
    u64 *out;
    u64 in_plane_stride, out_plane_stride;
    u64 in_blk_stride, out_blk_stride;
    u64 in_row_stride, in_col_stride, out_row_stride;  // output columns are always contiguous
    const u64 *tw_tail;    // [k][low]: w_R^(k*low), k < 2^t, low < R/2^t          (null when t == 0)
    const u64 *tw_core[2]; // step s < a-1: [k][lo] = w_M^(k*lo), k < E, lo < L = E^(a-1-s), M = E*L
    const u64 *in_scale;   // (scale^in_row_stride)^row, row < R, or null
    const u64 *out_scale;  // (scale^out_row_stride)^k, k < R, or null      (last pass, inverse coset)
    const u64 *tw_lo;      // W^i, i < 1024, W = omega^tw_mul                 (non-last passes)
    const u64 *tw_hi;      // W^(1024 i)
    const u64 *col_scale;  // scale^col, col < number of columns, or null    (first pass, forward coset)
    u64 out_mul;           // n^-1
    u64 n_in;
    u32 flags;
    u32 log_R, log_T, a;  // R = 2^log_R rows, T = 2^log_T columns per tile, a core steps (t = log_R - a log_E)
    u32 log_E;            // 4 or 3
    u32 cs;               // shared-memory column stride (elements)
    u64 s_sq[32];         // scale^(2^b)
    u64 w16[8];           // z^e, e < 8, z a primitive 16th root with z^(16/M) = w_M (M = 2, 4, 8, 16)
    // all table entries and out_mul, s_sq, w16 are in Montgomery form (glmont.cuh)
};

// The caller's input is read once and the final output written once: streaming cache hints keep
// them from displacing the intermediate vector and the twiddle tables in L2.  The intermediate
// (written by a non-last pass, read by the next one) uses the default policy and stays resident.
GL_HD u64 load_stream(const u64 *p) {
#if defined(__CUDA_ARCH__)
    return __ldcs(reinterpret_cast<const unsigned long long *>(p));
#else
    return *p;
#endif
}
GL_HD void store_stream(u64 *p, u64 v) {
#if defined(__CUDA_ARCH__)
    __stcs(reinterpret_cast<unsigned long long *>(p), v);
#else
    *p = v;
#endif
}

GL_HD constexpr int bitrev4_c(int x, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}

GL_HD u64 canon4(u64 x) {
#if defined(__CUDA_ARCH__)
    u32 x0, x1, r0, r1;
    m_unpack(x, x0, x1);
    // x >= p  <=>  x1 == 0xFFFFFFFF and x0 != 0,  and then x - p = x0 - 1
    asm("{\n .reg .pred q;\n"
        " setp.eq.u32     q, %3, 0xFFFFFFFF;\n"
        " setp.ne.and.u32 q, %2, 0, q;\n"
        " mov.u32         %0, %2;\n"
        " @q add.u32      %0, %2, 0xFFFFFFFF;\n"
        " selp.u32        %1, 0, %3, q;\n}"
        : "=&r"(r0), "=r"(r1)
        : "r"(x0), "r"(x1));
    return m_pack(r0, r1);
#else
    return lcanon(x);
#endif
}

// In-register M-point transform, decimation in time.  On entry v[j] = x[bitrev(j)], all
// canonical; on exit v[k] = X[k] (lazy).  w[e * ws] = w_M^e in Montgomery form, e < M/2.
// One flat loop of M/2 butterflies per stage with compile-time bounds: every index into v
// (registers) and w (constant bank) is a literal after unrolling.
template <int LOG_M, int S>
GL_HD void dit4_stage(u64 (&v)[1 << LOG_M], const u64 *w, const int ws) {
    constexpr int M = 1 << LOG_M, half = 1 << (S - 1);
#pragma unroll
    for (int i = 0; i < M / 2; ++i) {
        const int j = i & (half - 1);
        const int lo = ((i >> (S - 1)) << S) + j, hi = lo + half;
        const u64 u = v[lo];
        u64 t = v[hi];
        if (j == 0) {
            if (S > 1) t = canon4(t);
        } else {
            t = mont_mul(t, w[(j << (LOG_M - S)) * ws]);
        }
        v[lo] = ladd(u, t);
        v[hi] = lsub(u, t);
    }
}
template <int LOG_M, int S = 1>
GL_HD void dft4_dit(u64 (&v)[1 << LOG_M], const u64 *w, const int ws) {
    if constexpr (S <= LOG_M) {
        dit4_stage<LOG_M, S>(v, w, ws);
        dft4_dit<LOG_M, S + 1>(v, w, ws);
    }
}

GL_HD u64 mont_pow_sq4(const u64 *sq, u64 e) {
    u64 acc = GL_EPS;  // 1 in Montgomery form
    for (int b = 0; e; ++b, e >>= 1)
        if (e & 1) acc = mont_mul(acc, sq[b]);
    return acc;
}

// shared-memory index of (col, row): one pad element per E rows, column stride cs
GL_HD u32 sidx4(u32 col, u32 row, u32 cs, u32 log_E) { return col * cs + row + (row >> log_E); }

// column stride = 16/T (mod 16) elements for tiles of T = 2 or 4 columns: a 64-bit wavefront
// (16 lanes = T columns x 16/T consecutive rows) then falls into 16 distinct bank pairs
static inline u32 pass4_cs(u32 log_R, u32 log_T, u32 log_E) {
    const u32 rows = (1u << log_R) + ((1u << log_R) >> log_E);
    if (log_T == 0) return rows;
    const u32 want = 16u >> log_T;
    return rows + ((want + 16 - rows % 16) % 16);
}
static inline u32 pass4_threads(u32 log_R, u32 log_T, u32 log_E) {
    const u32 nt = (1u << (log_R + log_T)) >> log_E;
    return nt < 32 ? 32 : nt;
}
static inline size_t pass4_smem_elems(u32 log_R, u32 log_T, u32 log_E) {
    return (size_t)pass4_cs(log_R, log_T, log_E) << log_T;
}
// entries of the core twiddle tables of steps 0 .. a-2 (staged back to back in shared memory)
GL_HD u32 pass4_core_table_elems(u32 log_E, u32 a, u32 s) {
    return s + 1 < a ? 1u << (log_E * (a - s)) : 0;
}

// ---- tail: global -> 2^t-point transform over the top digit -> twiddle -> shared --------------
// `tables_ready()` blocks until the staged tables can be read (device: mbarrier wait; host: no-op);
// it is called after the global loads have been issued.  Index arithmetic is 32-bit (a plane has
// at most 2^30 elements) and strength-reduced: element j of a group sits j' * L rows further.
template <int TL, int LE, class Ready>
GL_HD void pass4_tail(const Pass4Params &P, const u32 tid, const u32 nthreads, const u32 bx, const u32 by,
                      const u32 bz, const u64 *tw_tail_s, u64 *S, Ready tables_ready) {
    constexpr int M = 1 << TL;
    const u32 log_L = P.log_R - TL, L = 1u << log_L;  // rows per value of the top digit (>= E)
    const u32 T = 1u << P.log_T;
    const u32 ngroups = L << P.log_T;
    const bool rowfast = (P.flags & P4_LAST) && P.log_T > 0;
    const bool check = (P.flags & P4_FIRST) != 0;
    const u32 n_in = (u32)P.n_in, rs = (u32)P.in_row_stride, cst = (u32)P.in_col_stride;
    const u32 Lrs = L * rs;           // global distance between consecutive values of the top digit
    const u32 Lp = L + (L >> LE);     // the same in (padded) shared memory
    const u32 blk0 = by * (u32)P.in_blk_stride + (bx << P.log_T) * cst;
    const u64 *plane0 = P.in + (u64)bz * P.in_plane_stride;
    // A thread owns NG = E / M groups (g = tid + i * nthreads).  All E loads are issued before any
    // arithmetic, so a pass pays one global-memory round trip, not NG of them.
    constexpr int NG = (1 << LE) >> TL;
    u64 v[NG][M];
    u32 idx0[NG];
    const bool full = (u32)NG * nthreads == ngroups;  // false only for tiles smaller than one warp's worth
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) {
        const u32 g = tid + (u32)gi * nthreads;
        u32 col, low;
        if (rowfast) {  // rows are the contiguous global dimension: lanes run along the rows
            low = g & (L - 1);
            col = g >> log_L;
        } else {
            col = g & (T - 1);
            low = g >> P.log_T;
        }
        idx0[gi] = g < ngroups ? blk0 + col * cst + low * rs : 0;  // groups past the tile read element 0
    }
    if (full && !check && (P.flags & P4_INPUT)) {  // the caller's vector: read once
#pragma unroll
        for (int gi = 0; gi < NG; ++gi)
#pragma unroll
            for (int j = 0; j < M; ++j) v[gi][j] = load_stream(plane0 + idx0[gi] + (u32)bitrev4_c(j, TL) * Lrs);
    } else if (full && !check) {
#pragma unroll
        for (int gi = 0; gi < NG; ++gi)
#pragma unroll
            for (int j = 0; j < M; ++j) v[gi][j] = plane0[idx0[gi] + (u32)bitrev4_c(j, TL) * Lrs];
    } else {
        // zero padding beyond n_in (code/fri.py:28-29): clamp the address, then clear by index
        const u32 last = check ? n_in - 1 : 0xFFFFFFFFu;  // n_in == 0 never launches
#pragma unroll
        for (int gi = 0; gi < NG; ++gi)
#pragma unroll
            for (int j = 0; j < M; ++j) {
                const u32 idx = idx0[gi] + (u32)bitrev4_c(j, TL) * Lrs;
                v[gi][j] = plane0[idx < last ? idx : last];
            }
#pragma unroll
        for (int gi = 0; gi < NG; ++gi)
#pragma unroll
            for (int j = 0; j < M; ++j) {
                const u32 idx = idx0[gi] + (u32)bitrev4_c(j, TL) * Lrs;
                if (idx > last || tid + (u32)gi * nthreads >= ngroups) v[gi][j] = 0;
            }
    }
    bool waited = false;
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) {
        const u32 g = tid + (u32)gi * nthreads;
        if (g < ngroups) {
            u32 col, low;
            if (rowfast) {
                low = g & (L - 1);
                col = g >> log_L;
            } else {
                col = g & (T - 1);
                low = g >> P.log_T;
            }
            if (P.in_scale) {
#pragma unroll
                for (int j = 0; j < M; ++j)
                    v[gi][j] = mont_mul(v[gi][j], P.in_scale[(u32)bitrev4_c(j, TL) * L + low]);
            }
            if (TL > 0) {
                dft4_dit<TL>(v[gi], P.w16, 16 >> TL);
                v[gi][0] = canon4(v[gi][0]);
                if (!waited) {
                    tables_ready();
                    waited = true;
                }
#pragma unroll
                for (int k = 1; k < M; ++k) v[gi][k] = mont_mul(v[gi][k], tw_tail_s[(u32)k * L + low]);
            }
            u64 *dst = S + col * P.cs + low + (low >> LE);
#pragma unroll
            for (int k = 0; k < M; ++k) dst[(u32)k * Lp] = v[gi][k];
        }
    }
    if (!waited) tables_ready();  // every thread passes the wait exactly once
}

// ---- core: one in-place E-point step over the digit of weight L = E^(a-1-s) ----------------------
template <int LE>
GL_HD void pass4_core(const Pass4Params &P, const u32 s, const u32 tid, const u32 nthreads, const u64 *tw_core_s,
                      u64 *S) {
    constexpr int E = 1 << LE;
    const u32 log_L = LE * (P.a - 1 - s), L = 1u << log_L;
    const u32 Lp = L + (L >> LE);  // padded distance between consecutive digit values (L = 1 or a multiple of E)
    const u32 T = 1u << P.log_T;
    const u32 ngroups = (1u << (P.log_R + P.log_T)) >> LE;
    // tables of the steps are staged back to back: step s starts after E * L_s' entries of every s' < s
    const u64 *tw = tw_core_s;
    for (u32 q = 0; q < s; ++q) tw += pass4_core_table_elems(LE, P.a, q);
#pragma unroll 1
    for (u32 g = tid; g < ngroups; g += nthreads) {
        const u32 col = g & (T - 1);
        const u32 q = g >> P.log_T;
        const u32 lo = q & (L - 1), hi = q >> log_L;
        const u32 row0 = ((hi << LE) << log_L) + lo;
        u64 *base = S + col * P.cs + row0 + (row0 >> LE);
        u64 v[E];
#pragma unroll
        for (int j = 0; j < E; ++j) v[j] = base[(u32)bitrev4_c(j, LE) * Lp];
        dft4_dit<LE>(v, P.w16, 16 >> LE);
        if (log_L > 0) {  // not the last step: twiddle w_(E L)^(lo k); products are canonical
            v[0] = canon4(v[0]);
#pragma unroll
            for (int k = 1; k < E; ++k) v[k] = mont_mul(v[k], tw[((u32)k << log_L) + lo]);
        }
#pragma unroll
        for (int k = 0; k < E; ++k) base[(u32)k * Lp] = v[k];
    }
}

// ---- out: shared -> output scaling -> global -------------------------------------------------------
// Thread (col, hi) owns the E positions hi*E + i it wrote in the last core step; they hold
// X[k], k = krev(hi) + (R/E) i, an arithmetic progression, so the inter-pass twiddle
// W^(colg k) is a running product.
template <int LE>
GL_HD void pass4_out(const Pass4Params &P, const u32 tid, const u32 nthreads, const u32 bx, const u32 by, const u32 bz,
                     const u64 *S) {
    constexpr u32 E = 1u << LE;
    const u32 T = 1u << P.log_T;
    const u32 ngroups = (1u << (P.log_R + P.log_T)) >> LE;
    const u32 t = P.log_R - LE * P.a;
    const bool last = (P.flags & P4_LAST) != 0;
    const u32 log_h = LE * (P.a - 1);
    const u32 log_kstride = P.log_R - LE;
    const u64 ostep = P.out_row_stride << log_kstride;  // global distance between consecutive i
#pragma unroll 1
    for (u32 g = tid; g < ngroups; g += nthreads) {
        const u32 col = g & (T - 1);
        const u32 hi = g >> P.log_T;
        // position hi*E + i  <->  k = kT + 2^t * (k' + E^(a-1) * i),  hi = kT * E^(a-1) + (k_1, .., k_{a-1})
        // where k' = k_1 + E k_2 + ... reverses the base-E digits that the earlier core steps produced
        const u32 kT = hi >> log_h;
        u32 digits = hi & ((1u << log_h) - 1), k1 = 0;
        for (u32 d = 1; d < P.a; ++d) {
            k1 = (k1 << LE) | (digits & (E - 1));
            digits >>= LE;
        }
        const u32 kbase = kT + (k1 << t);
        const u32 colg = (bx << P.log_T) + col;
        u64 *out = P.out + (u64)bz * P.out_plane_stride + (u64)by * P.out_blk_stride + colg +
                   (u64)kbase * P.out_row_stride;
        const u64 *src = S + col * P.cs + hi * (E + 1);  // rows hi*E .. hi*E+E-1: no pad inside
        if (!last) {
            // X[k] * W^(colg k) (* scale^colg): c[j] runs over k = kbase + kstride * (4 i + j)
            const u64 e0 = (u64)colg * kbase, e1 = (u64)colg << log_kstride;
            u64 c0 = mont_mul(P.tw_lo[e0 & 1023], P.tw_hi[e0 >> 10]);
            if (P.col_scale) c0 = mont_mul(c0, P.col_scale[colg]);
            const u64 step1 = mont_mul(P.tw_lo[e1 & 1023], P.tw_hi[e1 >> 10]);
            const u64 step2 = mont_mul(step1, step1);
            const u64 step4 = mont_mul(step2, step2);
            u64 c[4];
            c[0] = c0;
            c[1] = mont_mul(c0, step1);
            c[2] = mont_mul(c0, step2);
            c[3] = mont_mul(c[1], step2);
#pragma unroll 1
            for (u32 i = 0; i < E / 4; ++i) {
#pragma unroll
                for (u32 j = 0; j < 4; ++j) {
                    out[0] = mont_mul(src[4 * i + j], c[j]);
                    out += ostep;
                    c[j] = mont_mul(c[j], step4);
                }
            }
        } else if (P.out_scale) {
            const u64 cc = mont_mul(P.out_mul, mont_pow_sq4(P.s_sq, (u64)by * P.out_blk_stride + colg));
            const u64 *os = P.out_scale + kbase;
#pragma unroll 1
            for (u32 i = 0; i < E; ++i) {
                store_stream(out, mont_mul(mont_mul(src[i], cc), os[(u64)i << log_kstride]));
                out += ostep;
            }
        } else if (P.flags & P4_OUT_MUL) {
#pragma unroll 4
            for (u32 i = 0; i < E; ++i) {
                store_stream(out, mont_mul(src[i], P.out_mul));
                out += ostep;
            }
        } else {
#pragma unroll 4
            for (u32 i = 0; i < E; ++i) {
                store_stream(out, canon4(src[i]));
                out += ostep;
            }
        }
    }
}
