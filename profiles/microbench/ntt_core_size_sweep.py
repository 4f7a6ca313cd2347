"""Single-vector transform time vs core step size (B2S_NTT_LOG_E=3: 8-point, 4: 16-point), warm L2."""
import os
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from util import rand_bfe, root_of_unity
from stark_brainfuck_b200 import Engine
eng = Engine(0)
for lg in range(10, 21):
    n = 1 << lg; w = root_of_unity(lg)
    x = eng.upload(rand_bfe(1, n)); y = eng.empty(1, n)
    for _ in range(3): eng.ntt(x, lg, w, out=y)
    ms, _ = eng.ntt_timed(x, lg, w, out=y, iters=50)
    x3 = eng.upload(rand_bfe(3, 3 * n).reshape(3, n)); y3 = eng.empty(3, n)
    for _ in range(3): eng.ntt(x3, lg, w, out=y3)
    ms3, _ = eng.ntt_timed(x3, lg, w, out=y3, iters=50)
    print("LOG_E=%s 2^%d: 1 plane %.2f us, 3 planes %.2f us" % (os.environ.get("B2S_NTT_LOG_E"), lg, ms * 1e3, ms3 * 1e3))
