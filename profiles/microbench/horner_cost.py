#!/usr/bin/env python3
"""Multiplication count per point of the Brainfuck AIR's 47 constraints (tests/golden/air.json): evaluated monomial
by monomial as the reference hands them over (code/multivariate.py:105-116) against the greedy multivariate Horner
scheme quotient_prog.h compiles them into.  Costs in base-field multiplications: base x base 1, base x extension 3,
extension x extension 9 (lifted base columns count as base-field, as in the kernel).  CPU only.

    python profiles/microbench/horner_cost.py
"""
import json
import os
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def cost_mul(ta, tb):
    return 1 if ta == tb == "b" else 3 if "b" in (ta, tb) else 9


def monomial_cost(cons, is_base):
    """product of the base-field factors, times the coefficient, times the extension-field factors"""
    total = 0
    for exps, coef in cons:
        cext = bool(coef[1] or coef[2])
        nb = sum(e for v, e in enumerate(exps) if e and is_base(v))
        nx = sum(e for v, e in enumerate(exps) if e and not is_base(v))
        total += (nb - 1 + (3 if cext else 1) if nb else 0) + 9 * nx
    return total


def horner(monos, is_base, weighted=True):
    """(multiplications, kind of the value, stack slots, program words) of P = v * Q + R, greedy choice of v"""
    if len(monos) == 1 and not any(monos[0][0]):
        c = monos[0][1]
        return 0, "x" if c[1] or c[2] else "b", 0, 1
    cnt = Counter(v for e, _ in monos for v, x in enumerate(e) if x)
    if weighted:
        v = max(sorted(cnt), key=lambda k: ((cnt[k] - 1) * (1 if is_base(k) else 9), cnt[k]))
    else:
        v = max(sorted(cnt), key=lambda k: cnt[k])
    Q = [(tuple(x - (i == v) for i, x in enumerate(e)), c) for e, c in monos if e[v]]
    R = [(e, c) for e, c in monos if not e[v]]
    cq, tq, nq, oq = horner(Q, is_base, weighted)
    tv = "b" if is_base(v) else "x"
    cost, kind, need, words = cq + cost_mul(tq, tv), "b" if tq == tv == "b" else "x", nq, oq + 1
    if R:
        cr, tr, nr, o_r = horner(R, is_base, weighted)
        cost, words = cost + cr, words + o_r
        kind = "x" if tr == "x" else kind
        if not (len(R) == 1 and not any(R[0][0])):  # a constant is added in place, anything else goes over the stack
            need, words = (max(nq, nr + 1) if nq >= nr else max(nr, nq + 1)), words + 2
    return cost, kind, need, words


def main():
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "air.json")))
    tot = Counter()
    depth = 0
    print("%-18s %-11s %10s %10s %10s" % ("table", "constraints", "monomials", "horner", "(unweighted)"))
    for t in g["tables"]:
        W, bw = t["full_width"], t["base_width"]
        is_base = lambda v: (v % W) < bw  # noqa: E731
        for k in ("boundary", "transition", "terminal"):
            cons = [[(tuple(e), c) for e, c in con] for con in t[k]]
            a = sum(monomial_cost(c, is_base) for c in cons)
            h = [horner(c, is_base) for c in cons]
            u = sum(horner(c, is_base, False)[0] for c in cons)
            depth = max([depth] + [x[2] for x in h])
            print("%-18s %-11s %10d %10d %10d" % (t["name"], k, a, sum(x[0] for x in h), u))
            tot["mono"] += a
            tot["horner"] += sum(x[0] for x in h)
            tot["unweighted"] += u
            tot["words"] += sum(x[3] for x in h)
    print("%-18s %-11s %10d %10d %10d" % ("all", "", tot["mono"], tot["horner"], tot["unweighted"]))
    print("program words %d, deepest stack %d" % (tot["words"], depth))


if __name__ == "__main__":
    main()
