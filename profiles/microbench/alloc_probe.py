import sys, time, torch, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from stark_brainfuck_b200 import Engine
from util import root_of_unity
eng = Engine(0)
P = 18446744069414584321
def fresh(mb, reps=6, label=""):
    keep = []; ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); keep.append(torch.empty(mb << 20, dtype=torch.uint8, device="cuda")); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{label}: fresh {mb} MB x{reps}: " + " ".join(f"{t:.2f}" for t in ts) + " ms")
    return keep
k = fresh(240, label="idle")
x = eng.upload(np.random.default_rng(1).integers(0, P, (46, 1 << 20), dtype=np.uint64))
torch.cuda.synchronize()
for rep in range(3):
    y = eng.ntt(x, 20, root_of_unity(20), offset=7)      # queues ~0.8 ms of GPU work, uses the library's async pool
    k += fresh(240, reps=3, label="after a queued batched ntt")
torch.cuda.synchronize()
k += fresh(240, reps=3, label="after sync")
print("torch reserved GB", torch.cuda.memory_reserved() / 2**30)
