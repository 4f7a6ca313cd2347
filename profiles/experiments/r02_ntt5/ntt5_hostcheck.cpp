// tests/ntt5_hostcheck.cpp -- CPU check of the two-pass schedule of stark_brainfuck_b200/csrc/ntt5.cuh: the
// kernel's __host__ __device__ phase functions are run thread by thread (phases separated where the kernel has
// its barriers) under the host plan of ntt5_plan.h and compared with the CPU oracle (oracle/liboracle.so).
// Test infrastructure: built and run by tests/test_ntt4_host.py; never part of the product.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "ntt5_plan.h"

extern "C" {
int orc_ntt(u64 omega, const u64 *in, u64 *out, u64 n);
int orc_intt(u64 omega, const u64 *in, u64 *out, u64 n);
}

static std::vector<u64> build_table(const Tab4 &t) {
    std::vector<u64> tab;
    if (!t.used) return tab;
    const u64 cnt = (u64)1 << t.log_count;
    tab.resize(cnt < 2 ? 2 : cnt);
    const u64 r2n = (u64)1 << t.log_r2;
    for (u64 i = 0; i < tab.size(); ++i) {
        const u64 e = t.two_d ? (i >> t.log_r2) * (i & (r2n - 1)) : i;
        tab[i] = gl_to_mont(gl_mul(t.mul, gl_pow(t.base, e)));
    }
    return tab;
}

template <int LOG_R>
static void emulate_pass(const Pass5Plan &pl, u32 n_planes) {
    using N = N5<LOG_R>;
    const Pass5Params &P = pl.P;
    std::vector<u64> S((size_t)N5_SLOTS * N::CS);
    std::vector<u64> regs((size_t)P.Tc * N::B * 16);
    for (u32 plane = 0; plane < n_planes; ++plane)
        for (u32 bx = 0; bx < pl.grid_x; ++bx) {
            for (auto &x : S) x = 0xDEADBEEFDEADBEEFULL;
            const u32 c0 = bx * P.Tc;
            const u32 ncols = P.C - c0 < P.Tc ? P.C - c0 : P.Tc;
            const u64 *in = P.in + (u64)plane * P.in_plane_stride;
            u64 *out = P.out + (u64)plane * P.out_plane_stride;
            if (!P.last)
                for (u32 tid = 0; tid < (u32)N::B * N5_SLOTS; ++tid) {
                    u32 slot, row0;
                    n5_tile_thread<LOG_R>(tid, slot, row0);
                    if (slot >= ncols) continue;
                    for (u32 i = 0; i < 16; ++i)
                        S[slot * N::CS + i * N::ROW + row0] = in[((u64)(row0 + N::B * i) << P.log_C) + c0 + slot];
                }
            for (u32 col = 0; col < ncols; ++col)
                for (u32 t = 0; t < (u32)N::B; ++t) {
                    u64 v[16];
                    u64 *Sc = S.data() + col * N::CS;
                    if (!P.last) {
                        n5_step1_load_smem<LOG_R>(v, Sc, t);
                    } else {
                        const u64 *src = in + (u64)(c0 + col) * N::R + t;
                        for (int j = 0; j < 16; ++j) v[j] = src[(u32)bitrev4_c(j, 4) * N::B];
                    }
                    n5_step1_compute<LOG_R>(v, t, P.tw1, P.w16);
                    n5_step1_store<LOG_R>(v, Sc, t);
                }
            for (u32 col = 0; col < ncols; ++col)
                for (u32 t = 0; t < (u32)N::B; ++t) n5_step2<LOG_R>(S.data() + col * N::CS, t, P.tw2, P.w16);
            for (u32 col = 0; col < ncols; ++col)
                for (u32 t = 0; t < (u32)N::B; ++t) {
                    u64 v[16];
                    n5_step3_load<LOG_R>(v, S.data() + col * N::CS, t);
                    memcpy(&regs[((size_t)col * N::B + t) * 16], v, sizeof(v));
                }
            for (u32 col = 0; col < ncols; ++col)
                for (u32 t = 0; t < (u32)N::B; ++t) {
                    u64 v[16];
                    memcpy(v, &regs[((size_t)col * N::B + t) * 16], sizeof(v));
                    n5_step3_compute_store<LOG_R>(v, S.data() + col * N::CS, t, P, c0 + col);
                }
            for (u32 tid = 0; tid < (u32)N::B * N5_SLOTS; ++tid) {
                u32 slot, row0;
                n5_tile_thread<LOG_R>(tid, slot, row0);
                if (slot >= ncols) continue;
                for (u32 i = 0; i < 16; ++i)
                    out[((u64)(row0 + N::B * i) << P.log_C) + c0 + slot] = S[slot * N::CS + i * N::ROW + row0];
            }
        }
}

static u64 rng_state = 0x9E3779B97F4A7C15ULL;
static u64 rnd() {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return rng_state;
}

static int run_case(u32 log_n, bool inverse, u32 n_planes, u32 n_sm) {
    const u64 n = (u64)1 << log_n;
    u64 omega = 1753635133440165772ULL;  // code/algebra.py:129
    for (u32 i = 0; i < 32 - log_n; ++i) omega = gl_mul(omega, omega);
    const u64 w = inverse ? gl_inv(omega) : omega;
    Pass5Plan plan[2];
    plan5(log_n, w, inverse, n_planes, n_sm, plan);
    std::vector<u64> in(n * n_planes), out(n * n_planes, 0x1111111111111111ULL), work(n * n_planes, 0x2222);
    for (auto &x : in) {
        x = rnd() % GL_P;
        if ((rnd() & 15) == 0) x = GL_P - 1 - (rnd() & 3);
        if ((rnd() & 15) == 0) x = rnd() & 3;
    }
    std::vector<std::vector<u64>> keep;
    for (int ps = 0; ps < 2; ++ps) {
        Pass5Plan &pl = plan[ps];
        auto bind = [&](const Tab4 &t, const u64 *&dst) {
            if (!t.used) return;
            keep.push_back(build_table(t));
            dst = keep.back().data();
        };
        bind(pl.tw1, pl.P.tw1);
        bind(pl.tw2, pl.P.tw2);
        bind(pl.tw_lo, pl.P.tw_lo);
        bind(pl.tw_hi, pl.P.tw_hi);
        pl.P.in = ps == 0 ? in.data() : work.data();
        pl.P.out = ps == 0 ? work.data() : out.data();
        pl.P.in_plane_stride = pl.P.out_plane_stride = n;
        switch (pl.log_R) {
            case 8: emulate_pass<8>(pl, n_planes); break;
            case 9: emulate_pass<9>(pl, n_planes); break;
            case 10: emulate_pass<10>(pl, n_planes); break;
            case 11: emulate_pass<11>(pl, n_planes); break;
            default: return 1;
        }
    }
    std::vector<u64> want(n);
    for (u32 q = 0; q < n_planes; ++q) {
        if (inverse) orc_intt(omega, in.data() + q * n, want.data(), n); else orc_ntt(omega, in.data() + q * n, want.data(), n);
        if (memcmp(want.data(), out.data() + q * n, n * 8)) {
            u64 bad = 0;
            for (u64 i = 0; i < n; ++i) bad += want[i] != out[q * n + i];
            printf("MISMATCH log_n %u inverse %d plane %u: %llu of %llu differ\n", log_n, (int)inverse, q,
                   (unsigned long long)bad, (unsigned long long)n);
            return 1;
        }
    }
    return 0;
}

int main(int argc, char **argv) {
    const u32 max_log = argc > 1 ? (u32)atoi(argv[1]) : 20;
    int failed = 0, total = 0;
    for (u32 log_n = 16; log_n <= max_log; ++log_n)
        for (int inverse = 0; inverse < 2; ++inverse) {
            failed += run_case(log_n, inverse, 1, 148);
            ++total;
        }
    failed += run_case(16, false, 3, 148);  // batches: 8 columns per CTA
    failed += run_case(17, true, 2, 148);
    failed += run_case(16, false, 1, 7);    // few SMs: more columns per CTA than slots -> capped
    total += 3;
    printf("%d cases, %d failed\n", total, failed);
    return failed ? 1 : 0;
}
