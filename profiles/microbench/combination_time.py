"""Device time of b2s_combination on the column mix of the Brainfuck AIR (code/brainfuck_stark.py:241-298:
1 randomizer + 31 base + 15 extension + 29 quotient codewords -> 151 weights) at several domain sizes.
Run on a GPU box: python profiles/microbench/combination_time.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from stark_brainfuck_b200 import Engine  # noqa: E402

P = 18446744069414584321


def main():
    eng = Engine(0)
    rng = np.random.default_rng(5)
    rep = {}
    for logn in (16, 18, 20, 22):
        N = 1 << logn
        nb, nx = 31, 1 + 15 + 29
        cols = [eng.upload(rng.integers(0, P, (3, N), dtype=np.uint64))]
        cols += [eng.upload(rng.integers(0, P, (1, N), dtype=np.uint64)) for _ in range(nb)]
        cols += [eng.upload(rng.integers(0, P, (3, N), dtype=np.uint64)) for _ in range(nx - 1)]
        n = len(cols)
        wa = rng.integers(0, P, (n, 3), dtype=np.uint64)
        wb = rng.integers(0, P, (n, 3), dtype=np.uint64)
        wb[0] = 0
        shifts = rng.choice([0, 3, N // 8 - 5, N // 8 + 1, N // 4 - 7, N // 4 - 2, N // 4, N // 16], n)
        w = pow(7, (P - 1) >> logn, P)
        for _ in range(2):
            eng.combination(cols, wa, wb, shifts, N, 7, w)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 5
        e0.record()
        for _ in range(iters):
            eng.combination(cols, wa, wb, shifts, N, 7, w)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        bytes_ = N * (nb * 8 + nx * 24 + 24)
        muls = N * (nb * 6 + nx * 12)
        rep[str(logn)] = {"ms": round(ms, 4), "GB/s": round(bytes_ / ms / 1e6, 1), "Gmul/s": round(muls / ms / 1e6, 1),
                          "columns": n}
        del cols
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
