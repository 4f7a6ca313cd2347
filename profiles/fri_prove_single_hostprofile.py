"""cProfile of the host side of Fri.prove (code/fri.py:178-199) on one GPU, codeword already on the device: where the
milliseconds outside the kernels go.  Run on a GPU box: python profiles/fri_prove_single_hostprofile.py [log_n]"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from stark_brainfuck_b200 import Engine, mirror  # noqa: E402
from stark_brainfuck_b200.glue import DeviceCodeword, Glue  # noqa: E402
from util import root_of_unity  # noqa: E402

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << logn
mirror.register()
eng = Engine(0)
glue = Glue(mirror.binding, eng)
mirror.set_glue(glue)
m = mirror
coeffs = np.random.default_rng(logn).integers(0, 18446744069414584321, (3, n // 4), dtype=np.uint64)
planes = eng.ntt(eng.upload(coeffs), logn, root_of_unity(logn), offset=7)
fri = m.fri.Fri(m.field.generator(), m.field.primitive_nth_root(n), n, 4, 8, m.xfield)
best = 1e9
for _ in range(5):
    ps = m.ip.ProofStream()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fri.prove(DeviceCodeword(glue, planes, m.xfield), ps)
    torch.cuda.synchronize()
    best = min(best, time.perf_counter() - t0)
print("Fri.prove 2^%d wall: %.2f ms, transcript %d bytes" % (logn, best * 1e3, len(ps.serialize())))
pr = cProfile.Profile()
ps = m.ip.ProofStream()
pr.enable()
fri.prove(DeviceCodeword(glue, planes, m.xfield), ps)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
