"""FRI commit at 2^20 (fixed challenges): the device side only, for launch lists (ncu --metrics gpu__time_duration.sum)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from stark_brainfuck_b200 import Engine, mirror  # noqa: E402
from util import root_of_unity  # noqa: E402

P = 18446744069414584321
eng = Engine(0)
mirror.register()
tpl = mirror.binding.xfe_templates(mirror.xfield)
logn = 20
n = 1 << logn
cw0 = eng.upload(np.random.default_rng(5).integers(0, P, (3, n), dtype=np.uint64))
best = None
for _ in range(4):
    e0, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.merkle_field(cw0, tpl)
    cw, N, w, off = cw0, n, root_of_unity(logn), 7
    while N // 2 > 4:
        cw, _nodes = eng.fri_fold(cw, [3, 5, 7], off, w, tpl)
        N //= 2
        w, off = w * w % P, off * off % P
    e2.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e2)
    best = t if best is None else min(best, t)
print("FRI commit 2^20, device side: %.3f ms" % best)
# per round, warm, back to back (CUDA events between the rounds)
ev = [torch.cuda.Event(enable_timing=True)]
ev[0].record()
eng.merkle_field(cw0, tpl)
ev.append(torch.cuda.Event(enable_timing=True))
ev[-1].record()
cw, N, w, off = cw0, n, root_of_unity(logn), 7
sizes = [n]
while N // 2 > 4:
    cw, _nodes = eng.fri_fold(cw, [3, 5, 7], off, w, tpl)
    N //= 2
    w, off = w * w % P, off * off % P
    ev.append(torch.cuda.Event(enable_timing=True))
    ev[-1].record()
    sizes.append(N)
torch.cuda.synchronize()
print("per round (leaves: us):", ", ".join("%d: %.1f" % (sz, 1e3 * ev[i].elapsed_time(ev[i + 1])) for i, sz in enumerate(sizes)))
