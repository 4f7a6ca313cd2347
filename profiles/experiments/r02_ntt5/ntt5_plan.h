// ntt5_plan.h -- host-side plan of the two-pass schedule of ntt5.cuh (pure host code, shared by ntt.cu and
// tests/ntt5_hostcheck.cpp).  Tables are described (Tab4), the caller materialises them in Montgomery form.
#pragma once
#include "ntt4_plan.h"
#include "ntt5.cuh"

struct Pass5Plan {
    Pass5Params P;
    Tab4 tw1, tw2, tw_lo, tw_hi;
    u32 log_R = 0, grid_x = 1, threads = 32;
    size_t smem = 0;
};

// plain transforms (no coset scale, no zero padding) of 2^16 .. 2^22 points
static inline bool plan5_supported(u32 log_n, u64 n_in, bool do_scale) {
    return log_n >= 16 && log_n <= 22 && n_in == ((u64)1 << log_n) && !do_scale;  // (8 B = 1024 threads at R = 2^11)
}

// w: the root actually used (omega, or omega^-1 for the inverse).  n_sm: SMs to spread a single vector over.
static inline void plan5(u32 log_n, u64 w, bool inverse, u32 n_planes, u32 n_sm, Pass5Plan plan[2]) {
    const u64 n = (u64)1 << log_n;
    const u32 lg[2] = {(log_n + 1) / 2, log_n / 2};
    for (int ps = 0; ps < 2; ++ps) {
        Pass5Plan &pl = plan[ps];
        pl = Pass5Plan();
        Pass5Params &P = pl.P;
        const u32 log_R = lg[ps], log_C = log_n - log_R;
        const u32 B = (1u << log_R) >> 4;
        pl.log_R = log_R;
        P.in = nullptr;
        P.out = nullptr;
        P.in_plane_stride = P.out_plane_stride = 0;
        P.C = 1u << log_C;
        P.log_C = log_C;
        P.last = ps;
        // one wave with every SM busy for a single vector; full sectors per row for batches
        u32 tc = n_planes == 1 ? (P.C + n_sm - 1) / n_sm : N5_SLOTS;
        if (tc < 1) tc = 1;
        if (tc > N5_SLOTS) tc = N5_SLOTS;
        P.Tc = tc;
        pl.grid_x = (P.C + tc - 1) / tc;
        pl.threads = N5_SLOTS * B;  // all eight slots' threads move the tile; slots >= Tc idle while it is transformed
        pl.smem = sizeof(u64) * ((size_t)N5_SLOTS * (16 * (B + (B >> 4)) + 2) + ((size_t)1 << log_R) + B) + 16;
        const u64 wR = gl_pow(w, n >> log_R);  // primitive R-th root of this pass
        pl.tw1.used = pl.tw1.two_d = true;
        pl.tw1.base = wR;
        pl.tw1.log_count = log_R;
        pl.tw1.log_r2 = log_R - 4;
        pl.tw2.used = pl.tw2.two_d = true;
        pl.tw2.base = gl_pow(wR, 16);
        pl.tw2.log_count = log_R - 4;
        pl.tw2.log_r2 = log_R - 8;
        const u64 w16 = gl_pow(wR, (u64)1 << (log_R - 4));
        for (u32 e = 0; e < 8; ++e) P.w16[e] = gl_to_mont(gl_pow(w16, e));
        P.tw1 = P.tw2 = P.tw_lo = P.tw_hi = nullptr;
        if (ps == 0) {
            pl.tw_lo.used = true;
            pl.tw_lo.base = w;
            pl.tw_lo.log_count = 10;
            pl.tw_hi.used = true;
            pl.tw_hi.base = gl_pow(w, 1024);
            pl.tw_hi.log_count = log_n - 10;
            pl.tw_hi.mul = inverse ? gl_inv(n % GL_P) : 1;  // code/ntt.py:39, carried by the inter-pass twiddle
        }
    }
}
