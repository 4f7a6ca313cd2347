"""BASELINE config 3: cubic extension-field NTT of 2^18 points (three planes, one call) + evaluate_domain shapes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from util import root_of_unity  # noqa: E402
from stark_brainfuck_b200 import Engine  # noqa: E402

eng = Engine(0)
P = 18446744069414584321
rng = np.random.default_rng(3)
for logn in (16, 18, 20):
    n = 1 << logn
    w = root_of_unity(logn)
    x = eng.upload(rng.integers(0, P, (3, n), dtype=np.uint64))
    y = eng.empty(3, n)
    eng.ntt(x, logn, w, out=y)
    ms, _ = eng.ntt_timed(x, logn, w, out=y, iters=50)
    msc, _ = eng.ntt_timed(x, logn, w, offset=7, out=y, iters=50)
    print("XFE ntt 2^%d: %.2f us plain, %.2f us on the coset 7*w^k (%.1f / %.1f GB/s algorithmic)" %
          (logn, ms * 1e3, msc * 1e3, 48 * n / ms / 1e6, 48 * n / msc / 1e6))
