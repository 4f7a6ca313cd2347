"""Shared helpers for the parity tests: seeded inputs (stdlib random, identical to
tests/golden/make_golden.py) and the digest convention of SURVEY.md Appendix C."""
import hashlib
import json
import os
import random

import numpy as np

P = 18446744069414584321
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def have_golden(name):
    return os.path.exists(os.path.join(GOLDEN, name))


def rand_bfe(seed, n):
    R = random.Random(seed)
    return np.array([R.randrange(P) for _ in range(n)], dtype=np.uint64)


def rand_xfe(seed, n):
    """(3, n) planes; three draws per element in coefficient order (c0, c1, c2)"""
    R = random.Random(seed)
    a = np.array([R.randrange(P) for _ in range(3 * n)], dtype=np.uint64).reshape(n, 3)
    return np.ascontiguousarray(a.T)


def bfe_digest(v):
    return hashlib.sha256(np.ascontiguousarray(v, dtype="<u8").tobytes()).hexdigest()


def xfe_digest(planes):
    a = np.ascontiguousarray(np.asarray(planes, dtype="<u8").T)  # (n,3) AoS
    return hashlib.sha256(a.tobytes()).hexdigest()


def root_of_unity(log_n):
    """code/algebra.py:122-136 restated with python ints"""
    r = 1753635133440165772
    for _ in range(32 - log_n):
        r = r * r % P
    return r


def quotient_program(program):
    """tests/golden/quotients.json `program` ([[exponent vector, coefficient triple], ...] per constraint)
    -> (mono_off, coeffs, factors) arrays of b2s_quotients / orc_quotients"""
    mono_off, coeffs, facs = [0], [], []
    for constraint in program:
        for k, c in constraint:
            coeffs.append(c)
            facs.append([(i << 8) | e for i, e in enumerate(k) if e])
        mono_off.append(len(coeffs))
    width = max([len(f) for f in facs] + [1])
    factors = np.zeros((len(facs), width), dtype=np.uint32)
    for m, f in enumerate(facs):
        factors[m, :len(f)] = f
    return (np.asarray(mono_off, dtype=np.uint32), np.asarray(coeffs, dtype=np.uint64).reshape(-1, 3), factors)


def quotient_cases(g):
    """(name, cw (W,3,N), shift, program arrays, kind, height, omicron_inv, expected (C,3,N)) per golden case"""
    N, W = g["N"], g["width"]
    for ti, t in enumerate(g["tables"]):
        cw = np.array(t["codewords"], dtype=np.uint64).transpose(0, 2, 1).copy()
        for kind, name in ((1, "boundary"), (2, "transition"), (3, "terminal")):
            e = t[name]
            want = np.array(e["out"], dtype=np.uint64).reshape(-1, N, 3).transpose(0, 2, 1)
            yield ("table%d_%s" % (ti, name), cw, t["unit_distance"] if kind == 2 else 0, quotient_program(e["program"]),
                   kind, t["height"], t["omicron_inv"], want)
    pm = g["permutation"]
    lhs = np.array(g["tables"][pm["lhs"][0]]["codewords"][pm["lhs"][1]], dtype=np.uint64).T
    rhs = np.array(g["tables"][pm["rhs"][0]]["codewords"][pm["rhs"][1]], dtype=np.uint64).T
    prog = quotient_program([[[[1, 0], [1, 0, 0]], [[0, 1], [P - 1, 0, 0]]]])
    yield ("permutation", np.stack([lhs, rhs]).copy(), 0, prog, 1, 0, 1,
           np.array(pm["out"], dtype=np.uint64).T.reshape(1, 3, N))


def combination_cases(g):
    """tests/golden/combination.json -> (columns, wa, wb, shifts, N, offset, omega, want (3, N)) per case, in the
    argument convention of b2s_combination / orc_combination"""
    for c in g["cases"]:
        N = c["N"]
        cols = [np.array(c["randomizer"], dtype=np.uint64).T.copy()]
        cols += [np.array(col, dtype=np.uint64).reshape(1, N) for col in c["base"]]
        cols += [np.array(col, dtype=np.uint64).T.copy() for col in c["extension"] + c["quotient"]]
        bounds = c["base_degree_bounds"] + c["extension_degree_bounds"] + c["quotient_degree_bounds"]
        w = np.array(c["weights"], dtype=np.uint64)
        wa = np.concatenate([w[:1], w[1::2]])
        wb = np.concatenate([np.zeros((1, 3), dtype=np.uint64), w[2::2]])
        shifts = [0] + [c["max_degree"] - b for b in bounds]
        yield cols, wa, wb, shifts, N, c["offset"], c["omega"], np.array(c["out"], dtype=np.uint64).T.copy()
