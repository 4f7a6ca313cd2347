// quotient_prog.h -- the constraint programs of quotient.cu: the host compiler (expanded polynomial -> greedy
// multivariate Horner tree -> accumulator + stack program) and the interpreter the kernel runs per point.
// Host + device: tests/quotient_hostcheck.cpp compiles both with g++ and checks them against a direct
// monomial-by-monomial evaluation (code/multivariate.py:105-116) on the reference's real AIR (tests/golden/air.json).
#pragma once
#include <stdio.h>

#include <algorithm>
#include <vector>

#include "glmont.cuh"

#define Q_THREADS 128
#define Q_MAX_WORDS 96   // staged variable words + stack words per thread (1 KB of shared memory per word and CTA)
#define Q_MAX_STACK 16

// The constraint program (host-compiled, see b2s_quotients): one u32 per operation,
//     opcode | Q_A (the accumulator holds an extension-field value) | Q_S (so does the operand) | Q_D | arg << 8
enum : u32 {
    Q_LOADC = 0,   // acc = constant[arg]
    Q_ADDC = 1,    // acc += constant[arg]
    Q_MUL = 2,     // acc *= variable; arg = its first staged word, or with Q_D the variable's index (read from global)
    Q_PUSH = 3,    // stack[arg] = acc
    Q_ADDPOP = 4,  // acc += stack[arg]
    Q_END = 5,
    Q_A = 16,
    Q_S = 32,
    Q_D = 64,
    Q_EXT_VAR = 0x80000000u,  // load-list word: the variable is a genuine extension-field column (three words)
};


// ---- interpreter ---------------------------------------------------------------------------------------------
// Mem: var(v, j) = coefficient j of variable v at this point (global memory), get(w) / put(w, x) = the thread's
// word w of staged variables + stack (shared memory).
template <class Mem>
GL_HD const u32 *q_stage(const u32 *pc, Mem &mem) {
    const u32 n_loads = *pc++;
    u32 w = 0;
#pragma unroll 4
    for (u32 k = 0; k < n_loads; ++k) {
        const u32 lw = pc[k], v = lw & 0xFFFFFF;
        mem.put(w++, mem.var(v, 0));
        if (lw & Q_EXT_VAR) {
            mem.put(w++, mem.var(v, 1));
            mem.put(w++, mem.var(v, 2));
        }
    }
    return pc + n_loads;
}

template <class Mem>
GL_HD xfe q_run(const u32 *pc, const u64 *consts, Mem &mem) {
    xfe acc = {{0, 0, 0}};
    for (;;) {
        const u32 op = *pc++;
        const u32 arg = op >> 8;
        const u32 code = op & 15;
        if (code == Q_MUL) {
            u64 x0, x1 = 0, x2 = 0;
            if (op & Q_D) {
                x0 = mem.var(arg, 0);
                if (op & Q_S) {
                    x1 = mem.var(arg, 1);
                    x2 = mem.var(arg, 2);
                }
            } else {
                x0 = mem.get(arg);
                if (op & Q_S) {
                    x1 = mem.get(arg + 1);
                    x2 = mem.get(arg + 2);
                }
            }
            if (!(op & Q_S)) {
                acc.c[0] = mont_mul(acc.c[0], x0);
                if (op & Q_A) {
                    acc.c[1] = mont_mul(acc.c[1], x0);
                    acc.c[2] = mont_mul(acc.c[2], x0);
                }
            } else if (!(op & Q_A)) {  // base-field value times an extension-field variable
                const u64 a = acc.c[0];
                acc.c[0] = mont_mul(a, x0);
                acc.c[1] = mont_mul(a, x1);
                acc.c[2] = mont_mul(a, x2);
            } else {
                acc = x_mul_mont(acc, xfe{{x0, x1, x2}});
            }
        } else if (code == Q_ADDC) {
            const u64 *k = consts + 3 * (u64)arg;  // host-scaled constants are canonical
            acc.c[0] = ladd(acc.c[0], k[0]);
            if (op & Q_S) {
                acc.c[1] = ladd(acc.c[1], k[1]);
                acc.c[2] = ladd(acc.c[2], k[2]);
            }
        } else if (code == Q_LOADC) {
            const u64 *k = consts + 3 * (u64)arg;
            acc.c[0] = k[0];
            acc.c[1] = (op & Q_S) ? k[1] : 0;
            acc.c[2] = (op & Q_S) ? k[2] : 0;
        } else if (code == Q_PUSH) {
            mem.put(arg, lcanon(acc.c[0]));
            if (op & Q_A) {
                mem.put(arg + 1, lcanon(acc.c[1]));
                mem.put(arg + 2, lcanon(acc.c[2]));
            }
        } else if (code == Q_ADDPOP) {
            acc.c[0] = ladd(acc.c[0], mem.get(arg));
            if (op & Q_S) {
                acc.c[1] = ladd(acc.c[1], mem.get(arg + 1));
                acc.c[2] = ladd(acc.c[2], mem.get(arg + 2));
            }
        } else {
            return acc;
        }
    }
}

// ---- host: expanded constraint -> Horner tree -> program ------------------------------------------------------
struct QMono {
    std::vector<u32> e;  // exponent of every variable
    u64 coef[3];         // scaled by 2^(64 * degree)
};
struct QNode {
    int var = -1;  // < 0: leaf
    int q = -1, r = -1;
    u32 cidx = 0;     // leaf: its constant
    bool ext = false; // leaf: the constant has extension-field coefficients
    u32 need = 0;     // stack slots
};
struct QCompiler {
    u32 n_vars;
    const std::vector<u32> &kinds;  // per codeword: non-zero = lifted base-field column
    u32 width;
    std::vector<QNode> nodes;
    std::vector<u64> &consts;  // three words per constant, shared by the call
    std::vector<u32> code;
    std::vector<u32> slot;  // per variable: first staged word, or ~0 (read from global)
    u32 stack0 = 0;         // first stack word
    bool direct = false;

    QCompiler(u32 nv, const std::vector<u32> &k, u32 w, std::vector<u64> &cs) : n_vars(nv), kinds(k), width(w), consts(cs) {}
    bool var_base(u32 v) const { return kinds[v >= width ? v - width : v] != 0; }

    int build(std::vector<QMono> ms) {
        std::vector<u32> cnt(n_vars, 0);
        for (const QMono &m : ms)
            for (u32 v = 0; v < n_vars; ++v) cnt[v] += m.e[v] != 0;
        int best = -1;
        u64 best_w = 0;
        for (u32 v = 0; v < n_vars; ++v) {
            if (!cnt[v]) continue;
            // multiplications saved by taking v out of cnt monomials at once, then the count itself
            const u64 w = ((u64)(cnt[v] - 1) * (var_base(v) ? 1 : 9) << 20) | cnt[v];
            if (best < 0 || w > best_w) {
                best = (int)v;
                best_w = w;
            }
        }
        QNode n;
        if (best < 0) {  // constants only (equal exponent vectors are summed)
            u64 c[3] = {0, 0, 0};
            for (const QMono &m : ms)
                for (int j = 0; j < 3; ++j) c[j] = gl_add(c[j], m.coef[j]);
            n.cidx = (u32)(consts.size() / 3);
            n.ext = (c[1] | c[2]) != 0;
            consts.insert(consts.end(), c, c + 3);
            nodes.push_back(n);
            return (int)nodes.size() - 1;
        }
        std::vector<QMono> Q, R;
        for (QMono &m : ms) {
            if (m.e[best]) {
                --m.e[best];
                Q.push_back(std::move(m));
            } else {
                R.push_back(std::move(m));
            }
        }
        n.var = best;
        n.q = build(std::move(Q));
        n.r = R.empty() ? -1 : build(std::move(R));
        const u32 nq = nodes[n.q].need;
        if (n.r < 0 || nodes[n.r].var < 0) {
            n.need = nq;  // a constant is added in place
        } else {
            const u32 nr = nodes[n.r].need;
            n.need = nq >= nr ? std::max(nq, nr + 1) : std::max(nr, nq + 1);
        }
        nodes.push_back(n);
        return (int)nodes.size() - 1;
    }
    // emits code that leaves the node's value in the accumulator; `sp` = stack slots in use; returns "is extension"
    bool mul(bool acc_ext, u32 v) {
        const bool vx = !var_base(v);
        const u32 arg = direct ? v : slot[v];
        code.push_back(Q_MUL | (acc_ext ? Q_A : 0) | (vx ? Q_S : 0) | (direct ? Q_D : 0) | (arg << 8));
        return acc_ext || vx;
    }
    bool emit(int id, u32 sp) {
        const QNode &n = nodes[id];
        if (n.var < 0) {
            code.push_back(Q_LOADC | (n.ext ? Q_S : 0) | (n.cidx << 8));
            return n.ext;
        }
        if (n.r < 0) return mul(emit(n.q, sp), (u32)n.var);
        const QNode &r = nodes[n.r];
        if (r.var < 0) {
            const bool a = mul(emit(n.q, sp), (u32)n.var);
            code.push_back(Q_ADDC | (r.ext ? Q_S : 0) | (r.cidx << 8));
            return a || r.ext;
        }
        const u32 at = stack0 + 3 * sp;
        bool first, second;
        if (nodes[n.q].need >= r.need) {
            first = mul(emit(n.q, sp), (u32)n.var);
            code.push_back(Q_PUSH | (first ? Q_A : 0) | (at << 8));
            second = emit(n.r, sp + 1);
        } else {
            first = emit(n.r, sp);
            code.push_back(Q_PUSH | (first ? Q_A : 0) | (at << 8));
            second = mul(emit(n.q, sp + 1), (u32)n.var);
        }
        code.push_back(Q_ADDPOP | (first ? Q_S : 0) | (at << 8));
        return first || second;
    }
};

// Compiles every constraint of a call.  kinds[v] != 0: codeword v is a lifted base-field column.  stage: keep the
// variables of a constraint in shared memory when they fit.  Returns 0, or 1 with a message in `why`.
inline int q_compile(u32 width, u32 n_constraints, const u32 *h_mono_off, const u64 *h_coeffs, const u32 *h_factors,
                     u32 max_factors, const std::vector<u32> &kinds, bool stage, std::vector<u64> &consts,
                     std::vector<u32> &code, std::vector<u32> &prog_off, u32 &max_words, char *why, size_t why_len) {
    const u32 n_vars = 2 * width;
    prog_off.assign(n_constraints + 1, 0);
    max_words = 0;
    for (u32 c = 0; c < n_constraints; ++c) {
        prog_off[c] = (u32)code.size();
        std::vector<QMono> ms;
        std::vector<unsigned char> used(n_vars, 0);
        for (u32 m = h_mono_off[c]; m < h_mono_off[c + 1]; ++m) {
            QMono q;
            q.e.assign(n_vars, 0);
            u64 degree = 0;
            for (u32 f = 0; f < max_factors; ++f) {
                const u32 fac = h_factors[m * max_factors + f], e = fac & 0xFF, v = fac >> 8;
                if (e == 0) continue;
                q.e[v] += e;  // a variable listed twice multiplies twice
                used[v] = 1;
                degree += e;
            }
            const u64 r = gl_pow(GL_EPS, degree);  // 2^64 = EPS (mod p)
            for (int j = 0; j < 3; ++j) q.coef[j] = gl_mul(h_coeffs[3 * m + j] % GL_P, r);
            ms.push_back(std::move(q));
        }
        QCompiler qc(n_vars, kinds, width, consts);
        if (ms.empty()) {  // the zero polynomial
            QMono z;
            z.e.assign(n_vars, 0);
            z.coef[0] = z.coef[1] = z.coef[2] = 0;
            ms.push_back(z);
        }
        const int root = qc.build(std::move(ms));
        const u32 need = qc.nodes[root].need;
        if (need > Q_MAX_STACK) {
            snprintf(why, why_len, "constraint %u needs an evaluation stack of %u entries (limit %u)", c, need, Q_MAX_STACK);
            return 1;
        }
        // stage the constraint's variables in shared memory when they fit next to the stack
        qc.slot.assign(n_vars, 0xFFFFFFFFu);
        std::vector<u32> loads;
        u32 words = 0;
        for (u32 v = 0; v < n_vars; ++v)
            if (used[v]) {
                qc.slot[v] = words;
                words += qc.var_base(v) ? 1 : 3;
                loads.push_back(v | (qc.var_base(v) ? 0 : Q_EXT_VAR));
            }
        if (!stage || words + 3 * need > Q_MAX_WORDS) {
            qc.direct = true;
            loads.clear();
            words = 0;
        }
        qc.stack0 = words;
        max_words = std::max(max_words, words + 3 * need);
        code.push_back((u32)loads.size());
        code.insert(code.end(), loads.begin(), loads.end());
        qc.emit(root, 0);
        qc.code.push_back(Q_END);
        code.insert(code.end(), qc.code.begin(), qc.code.end());
    }
    prog_off[n_constraints] = (u32)code.size();
    if (consts.size() / 3 >= (1u << 24) || n_vars >= (1u << 24)) {
        snprintf(why, why_len, "program too large (%zu constants)", consts.size() / 3);
        return 1;
    }
    return 0;
}
