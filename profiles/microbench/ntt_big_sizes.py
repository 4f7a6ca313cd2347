"""Large transforms (3-pass plans, GB-sized planes) on one GPU: coset round trip bit-exact, oracle comparison
up to 2^25, device time.  Run on a GPU box: python profiles/microbench/ntt_big_sizes.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as orc  # noqa: E402
from stark_brainfuck_b200 import Engine  # noqa: E402
from util import root_of_unity  # noqa: E402

P = 18446744069414584321
eng = Engine(0)
for logn in (24, 25, 27, 28):
    n = 1 << logn
    x = np.random.default_rng(logn).integers(0, P, n, dtype=np.uint64)
    w = root_of_unity(logn)
    d = eng.upload(x)
    y = eng.ntt(d, logn, w, offset=7)
    back = eng.ntt(y, logn, w, offset=7, inverse=True)
    torch.cuda.synchronize()
    ok_rt = bool(torch.equal(back, d))
    ms, _ = eng.ntt_timed(d, logn, w, offset=7, out=y, iters=3)
    msg = "2^%d: coset round trip exact %s, forward %.3f ms (%.0f GB/s algorithmic)" % (logn, ok_rt, ms, 16.0 * n / ms / 1e6)
    if logn <= 25:
        t0 = time.time()
        ref = orc.coset_evaluate(7, w, x, n)
        msg += ", equals the CPU oracle %s (oracle %.1f s)" % (bool(np.array_equal(eng.download(y)[0], ref)), time.time() - t0)
    print(msg, flush=True)
    del d, y, back
    torch.cuda.empty_cache()
