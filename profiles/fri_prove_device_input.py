"""Fri.prove (code/fri.py:178-199) from a codeword that is already on the device (what the nonlinear combination
hands over): wall clock of the whole proof, commit phase, and a cProfile of the host side.
  python profiles/fri_prove_device_input.py [log_n] [num_colinearity_tests]"""
import cProfile
import json
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    s = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    from stark_brainfuck_b200 import Engine, mirror
    from stark_brainfuck_b200.glue import DeviceCodeword, Glue
    from util import root_of_unity
    mirror.register()
    eng = Engine(0)
    glue = Glue(mirror.binding, eng)
    mirror.set_glue(glue)
    m = mirror
    n = 1 << logn
    coeffs = np.random.default_rng(logn).integers(0, 18446744069414584321, (3, n // 4), dtype=np.uint64)
    planes = eng.ntt(eng.upload(coeffs), logn, root_of_unity(logn), offset=7)
    fri = m.fri.Fri(m.field.generator(), m.field.primitive_nth_root(n), n, 4, s, m.xfield)
    rep = {"log_n": logn, "num_colinearity_tests": s, "rounds": fri.num_rounds()}
    best, best_c = 1e9, 1e9
    for _ in range(4):
        ps = m.ip.ProofStream()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fri.prove(DeviceCodeword(glue, planes, m.xfield), ps)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
        ps = m.ip.ProofStream()
        t0 = time.perf_counter()
        fri.commit(DeviceCodeword(glue, planes, m.xfield), ps)
        torch.cuda.synchronize()
        best_c = min(best_c, time.perf_counter() - t0)
    rep["prove_ms"], rep["commit_ms"] = round(best * 1e3, 3), round(best_c * 1e3, 3)
    rep["transcript_bytes"] = len(ps.serialize())
    print(json.dumps(rep))
    pr = cProfile.Profile()
    pr.enable()
    fri.prove(DeviceCodeword(glue, planes, m.xfield), m.ip.ProofStream())
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(25)


if __name__ == "__main__":
    main()
