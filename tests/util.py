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
