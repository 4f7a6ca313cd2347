"""Small driver for ncu captures: a few launches of each hot kernel (not a benchmark)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from util import rand_bfe, rand_xfe, root_of_unity  # noqa: E402
from stark_brainfuck_b200 import Engine, mirror  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
eng = Engine(0)
logn = 20
n = 1 << logn
w = root_of_unity(logn)
x = eng.upload(rand_bfe(1, n))
y = eng.empty(1, n)
if what in ("all", "ntt"):
    for _ in range(3):
        eng.ntt(x, logn, w, out=y)
        eng.ntt(y, logn, w, inverse=True, out=y)
    xb = torch.randint(0, 2 ** 62, (32, n), dtype=torch.int64, device=eng.device)
    yb = eng.empty(32, n)
    for _ in range(2):
        eng.ntt(xb, logn, w, offset=7, out=yb)
if what.startswith("nttb"):  # batched transform of another size: nttb16 -> 2^16 x 512 planes (same total work as 32 x 2^20)
    lg = int(what[4:])
    q = (32 << 20) >> lg
    wl = root_of_unity(lg)
    xb = torch.randint(0, 2 ** 62, (q, 1 << lg), dtype=torch.int64, device=eng.device)
    yb = eng.empty(q, 1 << lg)
    for _ in range(3):
        eng.ntt(xb, lg, wl, offset=7, out=yb)
    ms, _ = eng.ntt_timed(xb, lg, wl, offset=7, out=yb, iters=5)
    print("batched 2^%d x %d planes: %.3f ms" % (lg, q, ms))
if what in ("all", "fri"):
    mirror.register()
    tpl = mirror.binding.xfe_templates(mirror.xfield)
    cw = eng.upload(rand_xfe(2, 1 << 18))
    for _ in range(2):
        nodes = eng.merkle_field(cw, tpl)
        nxt, nn = eng.fri_fold(cw, [3, 5, 7], 7, root_of_unity(18), tpl)
if what in ("all", "next"):  # next-row kernels: quotient codewords and the nonlinear combination at a 2^20 domain
    import numpy as np
    rng = np.random.default_rng(3)
    PM = 18446744069414584321
    cols = [eng.upload(rng.integers(0, PM, (3, n), dtype=np.uint64))]
    cols += [eng.upload(rng.integers(0, PM, (1, n), dtype=np.uint64)) for _ in range(31)]
    cols += [eng.upload(rng.integers(0, PM, (3, n), dtype=np.uint64)) for _ in range(44)]
    nc = len(cols)
    wa = rng.integers(0, PM, (nc, 3), dtype=np.uint64)
    wb = rng.integers(0, PM, (nc, 3), dtype=np.uint64)
    wb[0] = 0
    shifts = rng.choice([0, 3, n // 8 - 5, n // 8 + 1, n // 4 - 7, n // 4 - 2, n // 4, n // 16], nc)
    for _ in range(2):
        eng.combination(cols, wa, wb, shifts, n, 7, w)
    del cols
    # a transition-constraint-like program: 8 constraints x 12 monomials x up to 4 factors over 2 x 12 variables
    from util import quotient_program
    import random
    R = random.Random(5)
    W = 12
    program = []
    for _ in range(8):
        cons = []
        for _ in range(12):
            k = [0] * (2 * W)
            for _ in range(R.randrange(1, 5)):
                k[R.randrange(2 * W)] += R.randrange(1, 3)
            cons.append([k, [R.randrange(PM) for _ in range(3)]])
        program.append(cons)
    prog = quotient_program(program)
    cw = eng.upload(rng.integers(0, PM, (3 * W, n), dtype=np.uint64)).reshape(W, 3, n)
    for _ in range(2):
        eng.quotients(cw, n // 1024, *prog, 2, 1024, pow(root_of_unity(10), PM - 2, PM), 7, w)
if what == "frismall":  # the latency-bound tail rounds of a FRI commit: 2^13 .. 2^4 leaves in
    mirror.register()
    tpl = mirror.binding.xfe_templates(mirror.xfield)
    for lg in (13, 10, 8, 7, 5):
        cw = eng.upload(rand_xfe(2, 1 << lg))
        for _ in range(2):
            nxt, nn = eng.fri_fold(cw, [3, 5, 7], 7, root_of_unity(lg), tpl)
if what in ("all", "misc"):  # the small kernels: scale, point evaluation, gathers, openings, the multi-GPU exchange step
    import ctypes as C
    import numpy as np
    xs = eng.upload(rand_xfe(3, n))
    for _ in range(2):
        eng.scale(x, 7)
        eng.scale(xs, [3, 5, 7])
        eng.eval_points(eng.upload(rand_bfe(4, 1 << 16)), eng.upload(rand_bfe(5, 1 << 12)))
        eng.eval_points(eng.upload(rand_xfe(6, 1 << 16)), eng.upload(rand_bfe(7, 1 << 12)))
    mirror.register()
    tpl = mirror.binding.xfe_templates(mirror.xfield)
    nodes = eng.merkle_field(xs, tpl)
    idx = list(range(0, n, n // 64))
    for _ in range(2):
        eng.gather(xs, idx)
        eng.merkle_open(nodes, idx)
    # four-step exchange kernels as rank 0 of 4 sees them (2^20 = 1024 x 1024, 256 local rows), into a local buffer
    G, n1, n2 = 4, 1024, 1024
    q, cpp = n1 // G, n2 // G
    a = torch.randint(0, 2 ** 62, (q, n2), dtype=torch.int64, device=eng.device)
    send = torch.empty(G * cpp * q, dtype=torch.int64, device=eng.device)
    ptrs = (C.c_void_p * G)(*[send.data_ptr() + 8 * r * cpp * q for r in range(G)])
    c = torch.empty((cpp, n1), dtype=torch.int64, device=eng.device)
    for _ in range(2):
        eng.check(eng.lib.b2s_dist_twiddle_transpose(C.c_void_p(a.data_ptr()), a.stride(0), q, n2, 0, w, 1, ptrs, G, q, 0,
                                                     eng.stream_ptr()))
        eng.check(eng.lib.b2s_block_permute(C.c_void_p(send.data_ptr()), C.c_void_p(c.data_ptr()), G, cpp, q,
                                            eng.stream_ptr()))
torch.cuda.synchronize()
print("launches", eng.launch_count())
