// CPU check of the constraint programs of stark_brainfuck_b200/csrc/quotient_prog.h: the host compiler (expanded
// polynomial -> greedy Horner tree -> stack program) and the interpreter the kernel runs, compiled with g++, against
// a direct monomial-by-monomial evaluation (code/multivariate.py:105-116) in plain 128-bit arithmetic.
//
//   quotient_hostcheck < programs.txt
// input, whitespace separated (written by tests/test_ntt4_host.py from tests/golden/air.json and random programs):
//   n_cases, then per case:  width n_constraints max_factors n_points (even) unused
//                            kinds[width]  mono_off[n_constraints + 1]  coeffs[3 * n_mono]  factors[n_mono * max_factors]
//                            n_points x (2 * width x 3) variable values  (upper coefficients of base columns are 0)
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "quotient_prog.h"

static u64 rd() {
    unsigned long long x;
    if (scanf("%llu", &x) != 1) {
        fprintf(stderr, "short input\n");
        exit(2);
    }
    return (u64)x;
}

template <int K>
struct Mem {
    const u64 *vals[K];  // per point of the thread: [variable][3]
    std::vector<u64> words;
    u64 var(u32 v, int j, int k) const { return vals[k][3 * v + j]; }
    u64 get(u32 w, int k) const { return words.at(w * K + k); }
    void put(u32 w, int k, u64 x) { words.at(w * K + k) = x; }
};

int main() {
    const u64 n_cases = rd();
    u64 failed = 0, checked = 0, ops = 0;
    for (u64 cs = 0; cs < n_cases; ++cs) {
        const u32 width = (u32)rd(), nc = (u32)rd(), mf = (u32)rd(), n_points = (u32)rd();
        (void)rd();
        std::vector<u32> kinds(width), off(nc + 1);
        for (auto &k : kinds) k = (u32)rd();
        for (auto &o : off) o = (u32)rd();
        const u32 n_mono = off[nc];
        std::vector<u64> coeffs(3 * (size_t)n_mono);
        for (auto &c : coeffs) c = rd();
        std::vector<u32> factors((size_t)n_mono * mf);
        for (auto &f : factors) f = (u32)rd();
        std::vector<u64> consts, code;
        std::vector<u32> prog_off;
        u32 max_need = 0;
        char why[160];
        if (q_compile(width, nc, off.data(), coeffs.data(), factors.data(), mf, kinds, consts, code, prog_off, max_need, why,
                      sizeof(why))) {
            printf("case %llu: compile failed: %s\n", (unsigned long long)cs, why);
            ++failed;
            continue;
        }
        ops += code.size();
        std::vector<u64> vals((size_t)n_points * 6 * width);
        for (auto &v : vals) v = rd();
        auto direct = [&](u32 c, const u64 *pv) {  // sum_m coeff_m prod_f var^e
            xfe want = {{0, 0, 0}};
            for (u32 m = off[c]; m < off[c + 1]; ++m) {
                xfe t = {{coeffs[3 * m] % GL_P, coeffs[3 * m + 1] % GL_P, coeffs[3 * m + 2] % GL_P}};
                for (u32 f = 0; f < mf; ++f) {
                    const u32 fac = factors[(size_t)m * mf + f], e = fac & 0xFF, v = fac >> 8;
                    for (u32 k = 0; k < e; ++k) t = x_mul(t, xfe{{pv[3 * v], pv[3 * v + 1], pv[3 * v + 2]}});
                }
                want = x_add(want, t);
            }
            return want;
        };
        auto compare = [&](u32 c, u32 pt, const xfe &got) {
            const xfe want = direct(c, vals.data() + (size_t)pt * 6 * width);
            ++checked;
            for (int j = 0; j < 3; ++j)
                if (lcanon(got.c[j]) != want.c[j]) {
                    if (failed < 10) printf("case %llu constraint %u point %u coefficient %d differs\n", (unsigned long long)cs, c, pt, j);
                    ++failed;
                    return;
                }
        };
        const size_t stack_words = 3 * (size_t)(max_need ? max_need : 1);
        for (u32 c = 0; c < nc; ++c) {
            for (u32 pt = 0; pt < n_points; ++pt) {  // one point per thread
                Mem<1> mem{{vals.data() + (size_t)pt * 6 * width}, std::vector<u64>(stack_words, 0xDEADBEEFDEADBEEFull)};
                xfe acc[1];
                q_run<1>(code.data() + prog_off[c], consts.data(), mem, acc);
                compare(c, pt, acc[0]);
            }
            for (u32 pt = 0; pt + 1 < n_points; pt += 2) {  // two points per thread
                Mem<2> mem{{vals.data() + (size_t)pt * 6 * width, vals.data() + (size_t)(pt + 1) * 6 * width},
                           std::vector<u64>(2 * stack_words, 0xDEADBEEFDEADBEEFull)};
                xfe acc[2];
                q_run<2>(code.data() + prog_off[c], consts.data(), mem, acc);
                compare(c, pt, acc[0]);
                compare(c, pt + 1, acc[1]);
            }
        }
    }
    printf("%llu evaluations, %llu program words, %llu failed\n", (unsigned long long)checked, (unsigned long long)ops,
           (unsigned long long)failed);
    return failed != 0;
}
