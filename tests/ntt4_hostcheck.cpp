// tests/ntt4_hostcheck.cpp -- CPU check of the NTT pass logic (ntt4.cuh).
//
// Runs the __host__ __device__ phase functions of stark_brainfuck_b200/csrc/ntt4.cuh thread
// by thread (phases separated exactly where the kernel has its barriers) under the host
// plan of ntt4_plan.h, and compares the result with the CPU oracle (oracle/liboracle.so).
// Test infrastructure: built and run by tests/test_ntt4_host.py; never part of the product.
//
//   g++ -O2 -std=c++17 -I stark_brainfuck_b200/csrc tests/ntt4_hostcheck.cpp -L oracle -loracle
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "ntt4_plan.h"

extern "C" {
int orc_coset_evaluate(u64 offset, u64 omega, const u64 *coeffs, u64 m, u64 *out, u64 n);
int orc_coset_interpolate(u64 offset, u64 omega, const u64 *values, u64 *out, u64 n);
}

static std::vector<u64> build_table(const Tab4 &t) {
    std::vector<u64> tab;
    if (!t.used) return tab;
    const u64 cnt = (u64)1 << t.log_count;
    tab.resize(cnt < 2 ? 2 : cnt);
    if (t.two_d) {
        const u64 r2n = (u64)1 << t.log_r2;
        for (u64 i = 0; i < cnt; ++i) tab[i] = gl_to_mont(gl_pow(t.base, (i >> t.log_r2) * (i & (r2n - 1))));
    } else {
        u64 acc = 1;
        for (u64 i = 0; i < tab.size(); ++i) {
            tab[i] = gl_to_mont(acc);
            acc = gl_mul(acc, t.base);
        }
    }
    return tab;
}

template <int TL, int LE>
static void emulate_pass(const Pass4Plan &pl, u32 n_planes) {
    const Pass4Params &P = pl.P;
    const u32 nthreads = pass4_threads(P.log_R, P.log_T, LE);
    std::vector<u64> S(pass4_smem_elems(P.log_R, P.log_T, LE));
    // the TMA bulk copies: tables of the core steps back to back
    std::vector<u64> core;
    for (u32 st = 0; st < 2; ++st)
        if (pass4_core_table_elems(LE, P.a, st)) core.insert(core.end(), P.tw_core[st], P.tw_core[st] + pass4_core_table_elems(LE, P.a, st));
    for (u32 bz = 0; bz < n_planes; ++bz)
        for (u32 by = 0; by < pl.grid_y; ++by)
            for (u32 bx = 0; bx < pl.grid_x; ++bx) {
                for (auto &x : S) x = 0xDEADBEEFDEADBEEFULL;
                for (u32 t = 0; t < nthreads; ++t)
                    pass4_tail<TL, LE>(P, t, nthreads, bx, by, bz, P.tw_tail, S.data(), [] {});
                for (u32 s = 0; s < P.a; ++s)
                    for (u32 t = 0; t < nthreads; ++t) pass4_core<LE>(P, s, t, nthreads, core.data(), S.data());
                for (u32 t = 0; t < nthreads; ++t) pass4_out<LE>(P, t, nthreads, bx, by, bz, S.data());
            }
}

static int dispatch(const Pass4Plan &pl, u32 n_planes) {
    if (pl.P.log_E == 4) {
        switch (pl.tail) {
            case 0: emulate_pass<0, 4>(pl, n_planes); return 0;
            case 1: emulate_pass<1, 4>(pl, n_planes); return 0;
            case 2: emulate_pass<2, 4>(pl, n_planes); return 0;
            case 3: emulate_pass<3, 4>(pl, n_planes); return 0;
        }
    } else {
        switch (pl.tail) {
            case 0: emulate_pass<0, 3>(pl, n_planes); return 0;
            case 1: emulate_pass<1, 3>(pl, n_planes); return 0;
            case 2: emulate_pass<2, 3>(pl, n_planes); return 0;
        }
    }
    return -1;
}

static u64 rng_state = 0x9E3779B97F4A7C15ULL;
static u64 rnd() {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return rng_state;
}

static int run_case(u32 log_n, u64 n_in, u64 offset, bool inverse, u32 n_planes, u32 log_T = 2, u32 log_E = 4) {
    const u64 n = (u64)1 << log_n;
    u64 omega = 1753635133440165772ULL;  // code/algebra.py:129
    for (u32 i = 0; i < 32 - log_n; ++i) omega = gl_mul(omega, omega);
    const u64 w = inverse ? gl_inv(omega) : omega;
    const u64 scale = inverse ? gl_inv(offset) : offset;
    Pass4Plan plan[3];
    const int npass = plan4(log_n, n_in, w, scale, inverse, offset != 1, log_T, log_E, plan);

    std::vector<u64> in(n_in * n_planes), out(n * n_planes, 0x1111111111111111ULL), work(n * n_planes, 0x2222);
    for (auto &x : in) {
        x = rnd() % GL_P;
        if ((rnd() & 15) == 0) x = GL_P - 1 - (rnd() & 3);  // edge values
        if ((rnd() & 15) == 0) x = rnd() & 3;
    }
    std::vector<std::vector<u64>> keep;
    for (int ps = 0; ps < npass; ++ps) {
        Pass4Plan &pl = plan[ps];
        auto bind = [&](const Tab4 &t, const u64 *&dst) {
            if (!t.used) return;
            keep.push_back(build_table(t));
            dst = keep.back().data();
        };
        bind(pl.tw_tail, pl.P.tw_tail);
        bind(pl.tw_core[0], pl.P.tw_core[0]);
        bind(pl.tw_core[1], pl.P.tw_core[1]);
        bind(pl.in_scale, pl.P.in_scale);
        bind(pl.out_scale, pl.P.out_scale);
        bind(pl.tw_lo, pl.P.tw_lo);
        bind(pl.tw_hi, pl.P.tw_hi);
        bind(pl.col_scale, pl.P.col_scale);
        pl.P.in = pl.first ? in.data() : work.data();
        pl.P.in_plane_stride = pl.first ? n_in : n;
        pl.P.out = pl.last ? out.data() : work.data();
        pl.P.out_plane_stride = n;
        if (dispatch(pl, n_planes)) {
            printf("no kernel for log_R %u\n", pl.log_R);
            return 1;
        }
    }
    std::vector<u64> ref(n);
    for (u32 q = 0; q < n_planes; ++q) {
        if (inverse)
            orc_coset_interpolate(offset, omega, in.data() + q * n_in, ref.data(), n);
        else
            orc_coset_evaluate(offset, omega, in.data() + q * n_in, n_in, ref.data(), n);
        if (memcmp(ref.data(), out.data() + q * n, n * 8)) {
            u64 bad = 0, first = n;
            for (u64 i = 0; i < n; ++i)
                if (ref[i] != out[q * n + i]) {
                    if (first == n) first = i;
                    ++bad;
                }
            printf("MISMATCH log_n=%u n_in=%llu offset=%llu inverse=%d plane=%u: %llu wrong, first at %llu\n", log_n,
                   (unsigned long long)n_in, (unsigned long long)offset, (int)inverse, q, (unsigned long long)bad,
                   (unsigned long long)first);
            return 1;
        }
    }
    return 0;
}

int main(int argc, char **argv) {
    const u32 max_log = argc > 1 ? (u32)atoi(argv[1]) : 16;
    int fails = 0, cases = 0;
    for (u32 log_n = 4; log_n <= max_log; ++log_n) {
        if (!plan4_supported(log_n)) continue;
        const u64 n = (u64)1 << log_n;
        for (int inverse = 0; inverse < 2; ++inverse)
            for (u64 offset : {(u64)1, (u64)7}) {
                const u64 nin_list[3] = {n, n / 4, n / 4 + 3};
                for (u64 n_in : nin_list) {
                    if (inverse && n_in != n) continue;
                    ++cases;
                    fails += run_case(log_n, n_in, offset, inverse != 0, log_n <= 12 ? 2 : 1);
                    ++cases;  // 8-point core steps (small batches)
                    fails += run_case(log_n, n_in, offset, inverse != 0, log_n <= 12 ? 2 : 1, 2, 3);
                    if (log_n >= 12 && n_in == n) {  // narrower tiles (used for small batches)
                        cases += 2;
                        fails += run_case(log_n, n_in, offset, inverse != 0, 1, 0);
                        fails += run_case(log_n, n_in, offset, inverse != 0, 1, 1);
                    }
                }
            }
    }
    // one 3-pass plan (8+8+7)
    if (max_log >= 23) {
        ++cases;
        fails += run_case(23, (u64)1 << 21, 7, false, 1);
        ++cases;
        fails += run_case(23, (u64)1 << 23, 7, true, 1);
        ++cases;  // 8-point core steps: 2^23 = 2^8 * 2^8 * 2^7 with a = 2 and tails 2^2, 2^2, 2^1
        fails += run_case(23, (u64)1 << 21, 7, false, 1, 2, 3);
    }
    printf("%d cases, %d failed\n", cases, fails);
    return fails ? 1 : 0;
}
