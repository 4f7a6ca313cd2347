"""Stress / leak check on the GPU box: thousands of calls of every entry point with varying shapes;
free device memory must come back to where it started (the stream-ordered pool keeps scratch)."""
import random
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import torch
from util import rand_bfe, rand_xfe, root_of_unity, P
from stark_brainfuck_b200 import Engine, mirror
eng = Engine(0)
mirror.register()
tpl = mirror.binding.xfe_templates(mirror.xfield)
R = random.Random(1)
torch.cuda.synchronize()
free0, total = torch.cuda.mem_get_info()
for it in range(1500):
    lg = R.randrange(1, 19)
    q = R.choice([1, 1, 3, 5])
    n = 1 << lg
    x = torch.randint(0, 2 ** 62, (q, n), dtype=torch.int64, device=eng.device)
    off = R.choice([1, 7])
    y = eng.ntt(x, lg, root_of_unity(lg), offset=off)
    z = eng.ntt(y, lg, root_of_unity(lg), offset=off, inverse=True)
    if it % 50 == 0:
        assert torch.equal(z, x), (lg, q, off)
    if q == 3 and lg >= 1:
        nodes = eng.merkle_field(x % (2 ** 62), tpl)
        nxt, nn = eng.fri_fold(x, [3, 5, 7], 7, root_of_unity(lg), tpl)
    del x, y, z
torch.cuda.synchronize()
torch.cuda.empty_cache()
free1, _ = torch.cuda.mem_get_info()
print("free before %.1f MB, after %.1f MB, launches %d" % (free0 / 1e6, free1 / 1e6, eng.launch_count()))
