"""FRI timing on one GPU (not a benchmark contract; feeds DESIGN.md and bench.py's `extra`):
  * device-side commit phase at N = 2^log_n: Merkle tree of round 0, then fold + next tree per round
    with fixed challenges (no host Fiat-Shamir in the loop), CUDA events on the launching stream;
  * Fri.prove through the reference-named front end (mirror): wall clock incl. object marshalling."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from util import rand_xfe, root_of_unity  # noqa: E402
from stark_brainfuck_b200 import Engine, mirror  # noqa: E402


def device_commit(eng, tpl, log_n, expansion=4, reps=5):
    n = 1 << log_n
    cw0 = eng.upload(rand_xfe(7, n))
    alpha = [3, 5, 7]
    best = None
    for _ in range(reps):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        nodes = eng.merkle_field(cw0, tpl)
        e1.record()
        cw, N, w, off = cw0, n, root_of_unity(log_n), 7
        while N // 2 > expansion:
            cw, nodes = eng.fri_fold(cw, alpha, off, w, tpl)
            N //= 2
            w, off = w * w % eng_P, off * off % eng_P
        e2.record()
        torch.cuda.synchronize()
        t = (e0.elapsed_time(e1), e0.elapsed_time(e2))
        best = t if best is None or t[1] < best[1] else best
    return best


if __name__ == "__main__":
    from stark_brainfuck_b200.engine import P as eng_P
    eng = Engine(0)
    mirror.register()
    tpl = mirror.binding.xfe_templates(mirror.xfield)
    for log_n in (16, 18, 20):
        tree_ms, commit_ms = device_commit(eng, tpl, log_n)
        n = 1 << log_n
        print("device commit 2^%d: round-0 tree %.3f ms (%.1f GB/s of 152 B/leaf), all rounds %.3f ms (%.1f GB/s of 328 B/leaf)"
              % (log_n, tree_ms, 152.0 * n / tree_ms / 1e6, commit_ms, 328.0 * n / commit_ms / 1e6))
    # object-level prove through the front end
    import random
    m = mirror
    g = m.glue()
    for log_n in (12, 16, 18):
        n = 1 << log_n
        xf, bf = m.xfield, m.xfield.modulus.coefficients[0].field
        R = random.Random(200 + log_n)
        coeffs = [m.extension_field.ExtensionFieldElement(
            m.univariate.Polynomial([m.algebra.BaseFieldElement(R.randrange(eng_P), bf) for _ in range(3)]), xf)
            for _ in range(n // 4)]
        fri = m.fri.Fri(m.field.generator(), m.field.primitive_nth_root(n), n, 4, 8, xf)
        t0 = time.perf_counter()
        cw = fri.domain.xevaluate(m.univariate.Polynomial(coeffs))
        t1 = time.perf_counter()
        ps = m.ip.ProofStream()
        fri.prove(cw, ps)
        t2 = time.perf_counter()
        print("front end 2^%d: xevaluate %.3f s, Fri.prove %.3f s (%d objects in the transcript)"
              % (log_n, t1 - t0, t2 - t1, len(ps.objects)))
