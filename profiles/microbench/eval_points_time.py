"""Polynomial.evaluate_domain on arbitrary points (BASELINE config 3's generic path): 2^18 extension-field
coefficients at 2^10 points, against the oracle on a slice."""
import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from util import rand_xfe, rand_bfe
from oracle import oracle as orc
from stark_brainfuck_b200 import Engine
eng = Engine(0)
m, k = 1 << 18, 1 << 10
c = rand_xfe(3, m); pts = rand_xfe(4, k); bp = rand_bfe(5, k)
dc, dp, dbp = eng.upload(c), eng.upload(pts), eng.upload(bp)
for name, d in (("extension points", dp), ("base points", dbp)):
    out = eng.eval_points(dc, d); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = eng.eval_points(dc, d); e1.record(); torch.cuda.synchronize()
    print("2^18 XFE coefficients x 2^10 %s: %.3f ms" % (name, e0.elapsed_time(e1)))
ref = orc.eval_points(c[:, :4096], pts[:, :8])
got = eng.download(eng.eval_points(eng.upload(c[:, :4096]), eng.upload(pts[:, :8])))
print("matches oracle on a slice:", bool(np.array_equal(ref, got)))
