// CPU check of the constraint programs of stark_brainfuck_b200/csrc/quotient_prog.h: the host compiler (expanded
// polynomial -> greedy Horner tree -> stack program) and the interpreter the kernel runs, compiled with g++, against
// a direct monomial-by-monomial evaluation (code/multivariate.py:105-116) in plain 128-bit arithmetic.
//
//   quotient_hostcheck < programs.txt
// input, whitespace separated (written by tests/test_ntt4_host.py from tests/golden/air.json and random programs):
//   n_cases, then per case:  width n_constraints max_factors n_points stage
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

struct Mem {
    const u64 *vals;  // [variable][3]
    std::vector<u64> words;
    u64 var(u32 v, int j) const { return vals[3 * v + j]; }
    u64 get(u32 w) const { return words.at(w); }
    void put(u32 w, u64 x) { words.at(w) = x; }
};

int main() {
    const u64 n_cases = rd();
    u64 failed = 0, checked = 0, ops = 0;
    for (u64 cs = 0; cs < n_cases; ++cs) {
        const u32 width = (u32)rd(), nc = (u32)rd(), mf = (u32)rd(), n_points = (u32)rd();
        const bool stage = rd() != 0;
        std::vector<u32> kinds(width), off(nc + 1);
        for (auto &k : kinds) k = (u32)rd();
        for (auto &o : off) o = (u32)rd();
        const u32 n_mono = off[nc];
        std::vector<u64> coeffs(3 * (size_t)n_mono);
        for (auto &c : coeffs) c = rd();
        std::vector<u32> factors((size_t)n_mono * mf);
        for (auto &f : factors) f = (u32)rd();
        std::vector<u64> consts;
        std::vector<u32> code, prog_off;
        u32 max_words = 0;
        char why[160];
        if (q_compile(width, nc, off.data(), coeffs.data(), factors.data(), mf, kinds, stage, consts, code, prog_off, max_words,
                      why, sizeof(why))) {
            printf("case %llu: compile failed: %s\n", (unsigned long long)cs, why);
            ++failed;
            continue;
        }
        ops += code.size();
        std::vector<u64> vals(6 * (size_t)width);
        for (u32 pt = 0; pt < n_points; ++pt) {
            for (auto &v : vals) v = rd();
            for (u32 c = 0; c < nc; ++c) {
                // direct evaluation: sum_m coeff_m prod_f var^e
                xfe want = {{0, 0, 0}};
                for (u32 m = off[c]; m < off[c + 1]; ++m) {
                    xfe t = {{coeffs[3 * m] % GL_P, coeffs[3 * m + 1] % GL_P, coeffs[3 * m + 2] % GL_P}};
                    for (u32 f = 0; f < mf; ++f) {
                        const u32 fac = factors[(size_t)m * mf + f], e = fac & 0xFF, v = fac >> 8;
                        for (u32 k = 0; k < e; ++k) t = x_mul(t, xfe{{vals[3 * v], vals[3 * v + 1], vals[3 * v + 2]}});
                    }
                    want = x_add(want, t);
                }
                Mem mem{vals.data(), std::vector<u64>(max_words ? max_words : 1, 0xDEADBEEFDEADBEEFull)};
                const u32 *pc = q_stage(code.data() + prog_off[c], mem);
                const xfe got = q_run(pc, consts.data(), mem);
                ++checked;
                for (int j = 0; j < 3; ++j)
                    if (lcanon(got.c[j]) != want.c[j]) {
                        if (failed < 10) printf("case %llu constraint %u point %u coefficient %d differs\n", (unsigned long long)cs, c, pt, j);
                        ++failed;
                        break;
                    }
            }
        }
    }
    printf("%llu evaluations, %llu program words, %llu failed\n", (unsigned long long)checked, (unsigned long long)ops,
           (unsigned long long)failed);
    return failed != 0;
}
