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
#define Q_MAX_STACK 16

// The constraint program (host-compiled, see q_compile): 64-bit instructions for an accumulator + stack machine.
//   Q_START   [push the accumulator to stack[slot]]  then  acc = constant
//   Q_STEP    [acc += stack[slot]]  [acc *= variable]  [acc += constant]      in this order
//   Q_END
// The kind of every operand (base-field value = upper coefficients zero, or extension-field value) is static and
// travels in the instruction, so a multiplication costs 1, 3 or 9 base-field multiplications as the operands need.
enum : u32 {
    Q_END = 0,
    Q_START = 1,
    Q_STEP = 2,
    Q_FMT = 3,
    Q_HAS_STACK = 1u << 2,  // START: push first; STEP: add the popped value first
    Q_HAS_MUL = 1u << 3,
    Q_HAS_ADDC = 1u << 4,
    Q_ACC_X = 1u << 5,    // the accumulator holds an extension-field value when it is pushed / multiplied
    Q_VAR_X = 1u << 6,    // the variable is a genuine extension-field column
    Q_CONST_X = 1u << 7,  // the constant has extension-field coefficients
    Q_POP_X = 1u << 8,    // the popped value is an extension-field value
};
#define Q_SLOT_SHIFT 12   // 4 bits
#define Q_VAR_SHIFT 16    // 24 bits
#define Q_CONST_SHIFT 40  // 24 bits

// ---- interpreter ---------------------------------------------------------------------------------------------
// K points per thread share the decoding of every instruction.  Mem: var(v, j, k) = coefficient j of variable v
// at the thread's k-th point (global memory); get(w, k) / put(w, k, x) = word w of its stack (shared memory).
template <int K, class Mem>
GL_HD void q_run(const u64 *pc, const u64 *consts, Mem &mem, xfe (&acc)[K]) {
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = xfe{{0, 0, 0}};
    for (;;) {
        const u64 op = *pc++;
        const u32 lo = (u32)op;
        const u32 fmt = lo & Q_FMT;
        const u32 hi = (u32)(op >> 32);
        if (fmt == Q_STEP) {
            if (lo & Q_HAS_STACK) {
                const u32 w = 3 * ((lo >> Q_SLOT_SHIFT) & 15);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    acc[k].c[0] = ladd(acc[k].c[0], mem.get(w, k));
                    if (lo & Q_POP_X) {
                        acc[k].c[1] = ladd(acc[k].c[1], mem.get(w + 1, k));
                        acc[k].c[2] = ladd(acc[k].c[2], mem.get(w + 2, k));
                    }
                }
            }
            if (lo & Q_HAS_MUL) {
                const u32 v = (u32)(op >> Q_VAR_SHIFT) & 0xFFFFFF;  // bits 16 .. 39
                if (!(lo & Q_VAR_X)) {
                    u64 x[K];
#pragma unroll
                    for (int k = 0; k < K; ++k) x[k] = mem.var(v, 0, k);
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        acc[k].c[0] = mont_mul(acc[k].c[0], x[k]);
                        if (lo & Q_ACC_X) {
                            acc[k].c[1] = mont_mul(acc[k].c[1], x[k]);
                            acc[k].c[2] = mont_mul(acc[k].c[2], x[k]);
                        }
                    }
                } else {
                    xfe x[K];
#pragma unroll
                    for (int k = 0; k < K; ++k) x[k] = xfe{{mem.var(v, 0, k), mem.var(v, 1, k), mem.var(v, 2, k)}};
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        if (!(lo & Q_ACC_X)) {  // base-field value times an extension-field variable
                            const u64 a = acc[k].c[0];
                            acc[k].c[0] = mont_mul(a, x[k].c[0]);
                            acc[k].c[1] = mont_mul(a, x[k].c[1]);
                            acc[k].c[2] = mont_mul(a, x[k].c[2]);
                        } else {
                            acc[k] = x_mul_mont(acc[k], x[k]);
                        }
                    }
                }
            }
            if (lo & Q_HAS_ADDC) {  // host-scaled constants are canonical
                const u64 *cst = consts + 3 * (hi >> (Q_CONST_SHIFT - 32));
                const u64 c0 = cst[0];
#pragma unroll
                for (int k = 0; k < K; ++k) acc[k].c[0] = ladd(acc[k].c[0], c0);
                if (lo & Q_CONST_X) {
                    const u64 c1 = cst[1], c2 = cst[2];
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        acc[k].c[1] = ladd(acc[k].c[1], c1);
                        acc[k].c[2] = ladd(acc[k].c[2], c2);
                    }
                }
            }
        } else if (fmt == Q_START) {
            const u64 *cst = consts + 3 * (hi >> (Q_CONST_SHIFT - 32));
            if (lo & Q_HAS_STACK) {
                const u32 w = 3 * ((lo >> Q_SLOT_SHIFT) & 15);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    mem.put(w, k, lcanon(acc[k].c[0]));
                    if (lo & Q_ACC_X) {
                        mem.put(w + 1, k, lcanon(acc[k].c[1]));
                        mem.put(w + 2, k, lcanon(acc[k].c[2]));
                    }
                }
            }
            const u64 c0 = cst[0];
            const u64 c1 = (lo & Q_CONST_X) ? cst[1] : 0, c2 = (lo & Q_CONST_X) ? cst[2] : 0;
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = xfe{{c0, c1, c2}};
        } else {
            return;
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
    std::vector<u64> code;

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
    // ---- emission: a peephole over the parts  pop-add -> multiply -> add constant  of one Q_STEP
    u64 pending = 0;  // Q_STEP under construction
    void flush() {
        if (pending) code.push_back(pending);
        pending = 0;
    }
    void start(bool push, bool acc_ext, u32 slot, const QNode &leaf) {
        flush();
        code.push_back((u64)(Q_START | (push ? Q_HAS_STACK : 0u) | (push && acc_ext ? Q_ACC_X : 0u) | (leaf.ext ? Q_CONST_X : 0u)) |
                       ((u64)slot << Q_SLOT_SHIFT) | ((u64)leaf.cidx << Q_CONST_SHIFT));
    }
    void pop_add(u32 slot, bool popped_ext) {
        flush();  // always the first part of a step
        pending = (u64)(Q_STEP | Q_HAS_STACK | (popped_ext ? Q_POP_X : 0u)) | ((u64)slot << Q_SLOT_SHIFT);
    }
    bool mul(bool acc_ext, u32 v) {
        if (pending & (u64)(Q_HAS_MUL | Q_HAS_ADDC)) flush();
        if (!pending) pending = Q_STEP;
        const bool vx = !var_base(v);
        pending |= (u64)(Q_HAS_MUL | (acc_ext ? Q_ACC_X : 0u) | (vx ? Q_VAR_X : 0u)) | ((u64)v << Q_VAR_SHIFT);
        return acc_ext || vx;
    }
    void add_const(const QNode &leaf) {
        if (pending & (u64)Q_HAS_ADDC) flush();
        if (!pending) pending = Q_STEP;
        pending |= (u64)(Q_HAS_ADDC | (leaf.ext ? Q_CONST_X : 0u)) | ((u64)leaf.cidx << Q_CONST_SHIFT);
    }
    // emits code that leaves the node's value in the accumulator.  `sp` = stack slots in use; `push_first`: the
    // accumulator holds a value (of kind push_ext) that must be saved to slot sp - 1 before it is overwritten --
    // the save rides on the Q_START of the first leaf.  Returns "the value is an extension-field value".
    bool emit(int id, u32 sp, bool push_first = false, bool push_ext = false) {
        const QNode &n = nodes[id];
        if (n.var < 0) {
            start(push_first, push_ext, push_first ? sp - 1 : 0, n);
            return n.ext;
        }
        if (n.r < 0) return mul(emit(n.q, sp, push_first, push_ext), (u32)n.var);
        const QNode &r = nodes[n.r];
        if (r.var < 0) {
            const bool a = mul(emit(n.q, sp, push_first, push_ext), (u32)n.var);
            add_const(r);
            return a || r.ext;
        }
        bool first, second;
        if (nodes[n.q].need >= r.need) {
            first = mul(emit(n.q, sp, push_first, push_ext), (u32)n.var);
            second = emit(n.r, sp + 1, true, first);
        } else {
            first = emit(n.r, sp, push_first, push_ext);
            second = mul(emit(n.q, sp + 1, true, first), (u32)n.var);
        }
        pop_add(sp, first);
        return first || second;
    }
};

// Compiles every constraint of a call.  kinds[v] != 0: codeword v is a lifted base-field column.  prog_off[c] = first
// instruction of constraint c in `code`; max_need = stack slots the deepest constraint uses.  Returns 0, or 1 with a
// message in `why`.
inline int q_compile(u32 width, u32 n_constraints, const u32 *h_mono_off, const u64 *h_coeffs, const u32 *h_factors,
                     u32 max_factors, const std::vector<u32> &kinds, std::vector<u64> &consts, std::vector<u64> &code,
                     std::vector<u32> &prog_off, u32 &max_need, char *why, size_t why_len) {
    const u32 n_vars = 2 * width;
    prog_off.assign(n_constraints + 1, 0);
    max_need = 0;
    if (n_vars >= (1u << 24)) {
        snprintf(why, why_len, "too many variables (%u)", n_vars);
        return 1;
    }
    for (u32 c = 0; c < n_constraints; ++c) {
        prog_off[c] = (u32)code.size();
        std::vector<QMono> ms;
        for (u32 m = h_mono_off[c]; m < h_mono_off[c + 1]; ++m) {
            QMono q;
            q.e.assign(n_vars, 0);
            u64 degree = 0;
            for (u32 f = 0; f < max_factors; ++f) {
                const u32 fac = h_factors[m * max_factors + f], e = fac & 0xFF, v = fac >> 8;
                if (e == 0) continue;
                q.e[v] += e;  // a variable listed twice multiplies twice
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
        max_need = std::max(max_need, need);
        qc.emit(root, 0);
        qc.flush();
        qc.code.push_back(Q_END);
        code.insert(code.end(), qc.code.begin(), qc.code.end());
    }
    prog_off[n_constraints] = (u32)code.size();
    if (consts.size() / 3 >= (1u << 24)) {
        snprintf(why, why_len, "program too large (%zu constants)", consts.size() / 3);
        return 1;
    }
    return 0;
}
