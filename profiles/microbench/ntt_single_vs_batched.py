"""One transform vs a batch of the same total size; host time per call.  Run on the GPU box."""
import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from util import rand_bfe, root_of_unity
from stark_brainfuck_b200 import Engine
eng = Engine(0)
for lg in (18, 20):
    n = 1 << lg; w = root_of_unity(lg)
    x = eng.upload(rand_bfe(1, n)); y = eng.empty(1, n)
    for _ in range(3): eng.ntt(x, lg, w, out=y)
    ms, _ = eng.ntt_timed(x, lg, w, out=y, iters=20)
    xb = torch.randint(0, 2 ** 62, ((32 << 20) >> lg, n), dtype=torch.int64, device=eng.device); yb = torch.empty_like(xb)
    eng.ntt(xb, lg, w, out=yb)
    msb, _ = eng.ntt_timed(xb, lg, w, out=yb, iters=5)
    print("2^%d: one vector (warm L2, 20 back-to-back calls) %.2f us; 2^25 elements in one call %.1f us" % (lg, ms * 1e3, msb * 1e3))
import time
lg = 20; n = 1 << lg; w = root_of_unity(lg)
x = eng.upload(rand_bfe(1, n)); y = eng.empty(1, n)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): eng.ntt(x, lg, w, out=y)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host time per eng.ntt call (no sync): %.2f us; incl. drain %.2f us" % ((t1 - t0) / 200 * 1e6, (t2 - t0) / 200 * 1e6))
