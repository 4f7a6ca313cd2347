// ntt4_plan.h -- host-side plan of a transform of length 2^log_n as 1-3 passes
// (ntt4.cuh): Cooley-Tukey on the index digits, most significant input digit first.  Pure
// host code without CUDA calls, shared by ntt.cu and the CPU check (tests/ntt4_hostcheck.cpp).
//
//   pass 0 .. npass-2  rows = current top digit (stride = number of lower-digit columns),
//                      result stays in place (tile shape kept), times the inter-pass twiddle
//   last pass          rows = lowest input digit (contiguous), writes the natural-order output
// Tables are described, not built: (base, count, layout); the caller materialises them in
// Montgomery form.
#pragma once
#include <stdlib.h>

#include "ntt4.cuh"

struct Tab4 {
    bool used = false;
    bool two_d = false;  // entry [k*L + lo] = base^(k*lo), L = 2^log_r2, instead of base^i
    u64 base = 1;
    u32 log_count = 0;
    u32 log_r2 = 0;
};

struct Pass4Plan {
    Pass4Params P;  // table pointers, in/out and plane strides are filled in by the caller
    Tab4 tw_tail, tw_core[2], in_scale, out_scale, tw_lo, tw_hi, col_scale;
    u32 log_R = 0, log_T = 0, tail = 0;
    bool first = false, last = false;
    u32 grid_x = 1, grid_y = 1;
};

static inline bool plan4_supported(u32 log_n) { return log_n >= 4 && log_n <= 30; }

// w: the root actually used (omega, or omega^-1 for the inverse); scale: offset or offset^-1.
// Returns the number of passes.
// log_E: 4 = 16-point core steps (fewest instructions), 3 = 8-point core steps (twice the threads).
static inline int plan4(u32 log_n, u64 n_in, u64 w, u64 scale, bool inverse, bool do_scale, u32 log_T_multi, u32 log_E,
                        Pass4Plan plan[3]) {
    const u64 n = (u64)1 << log_n;
    u32 lg[3] = {0, 0, 0};
    int npass;
    if (log_n <= 11) {
        npass = 1;
        lg[0] = log_n;
    } else if (log_n <= 22) {
        npass = 2;
        lg[0] = (log_n + 1) / 2;
        lg[1] = log_n / 2;
    } else {
        npass = 3;
        lg[0] = (log_n + 2) / 3;
        lg[1] = (log_n - lg[0] + 1) / 2;
        lg[2] = log_n - lg[0] - lg[1];
    }
    u64 s_sq[32];
    u64 ss = scale;
    for (int b = 0; b < 32; ++b) {
        s_sq[b] = gl_to_mont(ss);
        ss = gl_mul(ss, ss);
    }
    const u64 ninv = gl_inv(n % GL_P);  // code/ntt.py:39
    const u32 LOG_T_MULTI = log_T_multi;  // columns per tile in multi-pass plans (2: 32-byte segments)

    u32 rest = log_n;
    for (int ps = 0; ps < npass; ++ps) {
        Pass4Plan &pl = plan[ps];
        pl = Pass4Plan();
        Pass4Params &P = pl.P;
        const u32 log_R = lg[ps];
        rest -= log_R;
        const bool first = ps == 0, last = ps + 1 == npass;
        const u32 a = log_R / log_E, t = log_R - log_E * a;  // R = 2^t * E^a (log_R <= 11: a <= 2 resp. 3)
        pl.log_R = log_R;
        pl.tail = t;
        pl.log_T = npass == 1 ? 0 : LOG_T_MULTI;
        pl.first = first;
        pl.last = last;
        P.in = nullptr;
        P.out = nullptr;
        P.in_plane_stride = P.out_plane_stride = 0;
        P.tw_tail = P.tw_core[0] = P.tw_core[1] = P.in_scale = P.out_scale = P.tw_lo = P.tw_hi = P.col_scale = nullptr;
        P.n_in = n_in;
        P.flags = (first && n_in < n ? P4_FIRST : 0) | (last ? P4_LAST : 0) | (first ? P4_INPUT : 0);  // bounds checks only when padding
        P.log_R = log_R;
        P.log_T = pl.log_T;
        P.a = a;
        P.log_E = log_E;
        P.cs = pass4_cs(log_R, pl.log_T, log_E);
        P.out_mul = GL_EPS;  // 1
        for (int b = 0; b < 32; ++b) P.s_sq[b] = s_sq[b];
        const u64 wR = gl_pow(w, n >> log_R);  // primitive R-th root
        const u64 w16 = gl_pow(wR, (u64)1 << (log_R - 4));
        for (u32 e = 0; e < 8; ++e) P.w16[e] = gl_to_mont(gl_pow(w16, e));
        if (t > 0) {  // w_R^(k * low), k < 2^t, low < R / 2^t
            pl.tw_tail.used = true;
            pl.tw_tail.two_d = true;
            pl.tw_tail.base = wR;
            pl.tw_tail.log_count = log_R;
            pl.tw_tail.log_r2 = log_R - t;
        }
        for (u32 st = 0; st + 1 < a; ++st) {  // core step st: w_M^(k * lo), k < E, lo < L = E^(a-1-st), M = E L
            const u32 log_M = log_E * (a - st);
            pl.tw_core[st].used = true;
            pl.tw_core[st].two_d = true;
            pl.tw_core[st].base = gl_pow(wR, (u64)1 << (log_R - log_M));
            pl.tw_core[st].log_count = log_M;
            pl.tw_core[st].log_r2 = log_M - log_E;
        }
        if (!last) {
            const u64 ncols = (u64)1 << rest;
            P.in_row_stride = P.out_row_stride = ncols;
            P.in_col_stride = 1;
            P.in_blk_stride = P.out_blk_stride = first ? 0 : ((u64)1 << (log_n - lg[0]));
            const u64 W = first ? w : gl_pow(w, (u64)1 << lg[0]);  // twiddle w^(tw_mul * col * k)
            pl.tw_lo.used = true;
            pl.tw_lo.base = W;
            pl.tw_lo.log_count = 10;
            pl.tw_hi.used = true;
            pl.tw_hi.base = gl_pow(W, 1024);
            pl.tw_hi.log_count = rest + log_R > 10 ? rest + log_R - 10 : 0;  // exponents col * k < 2^(rest + log_R)
            if (first && do_scale && !inverse) {
                pl.in_scale.used = true;
                pl.in_scale.base = gl_pow(scale, ncols);
                pl.in_scale.log_count = log_R;
                pl.col_scale.used = true;
                pl.col_scale.base = scale;
                pl.col_scale.log_count = rest;
            }
            pl.grid_x = (u32)(ncols >> pl.log_T);
            pl.grid_y = first ? 1 : (1u << lg[0]);
        } else if (npass == 1) {
            P.in_row_stride = P.out_row_stride = 1;
            P.in_col_stride = 0;
            P.in_blk_stride = P.out_blk_stride = 0;
            if (do_scale && !inverse) {
                pl.in_scale.used = true;
                pl.in_scale.base = scale;
                pl.in_scale.log_count = log_R;
            }
        } else {
            // rows = lowest input digit j1 (contiguous), tile columns = adjacent values of the
            // FIRST pass's output digit, blocks = middle digit (3-pass plans)
            P.in_row_stride = 1;
            if (npass == 2) {
                P.in_col_stride = (u64)1 << log_R;
                P.in_blk_stride = P.out_blk_stride = 0;
                P.out_row_stride = (u64)1 << lg[0];
                pl.grid_x = (1u << lg[0]) >> pl.log_T;
            } else {
                P.in_col_stride = (u64)1 << (lg[1] + lg[2]);
                P.in_blk_stride = (u64)1 << lg[2];
                P.out_blk_stride = (u64)1 << lg[0];
                P.out_row_stride = (u64)1 << (lg[0] + lg[1]);
                pl.grid_x = (1u << lg[0]) >> pl.log_T;
                pl.grid_y = 1u << lg[1];
            }
        }
        if (last && inverse) {
            P.out_mul = gl_to_mont(ninv);
            P.flags |= P4_OUT_MUL;
            if (do_scale) {
                pl.out_scale.used = true;
                pl.out_scale.base = gl_pow(scale, P.out_row_stride);
                pl.out_scale.log_count = log_R;
            }
        }
    }
    return npass;
}
