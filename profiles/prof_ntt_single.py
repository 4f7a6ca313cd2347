"""ncu driver: plain 2^20 transforms, single vector and 32 planes (the ntt5 schedule unless B2S_NTT5=0)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from util import rand_bfe, root_of_unity  # noqa: E402
from stark_brainfuck_b200 import Engine  # noqa: E402

eng = Engine(0)
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << logn
w = root_of_unity(logn)
x = eng.upload(rand_bfe(1, n))
y = eng.empty(1, n)
for _ in range(3):
    eng.ntt(x, logn, w, out=y)
    eng.ntt(y, logn, w, inverse=True, out=y)
xb = torch.randint(0, 2 ** 62, (32, n), dtype=torch.int64, device=eng.device)
yb = eng.empty(32, n)
for _ in range(2):
    eng.ntt(xb, logn, w, out=yb)
torch.cuda.synchronize()
ms1, _ = eng.ntt_timed(x, logn, w, out=y, iters=50)
msb, _ = eng.ntt_timed(xb, logn, w, out=yb, iters=5)
print("single 2^%d: %.2f us   32 planes: %.1f us (%.2f us/plane)" % (logn, ms1 * 1e3, msb * 1e3, msb * 1e3 / 32))
