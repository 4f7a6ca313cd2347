"""cProfile of the host side of one FRI prove (one-rank group) -- where the milliseconds outside the
kernels go.  Run on a GPU box: python profiles/fri_prove_hostprofile.py [log_n]"""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29541")

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    import frontend_cases as fc
    from stark_brainfuck_b200 import Engine, mirror
    from stark_brainfuck_b200.dist_fri import DistFri
    from stark_brainfuck_b200.glue import Glue
    from util import root_of_unity
    mirror.register()
    eng = Engine(0)
    glue = Glue(mirror.binding, eng)
    mirror.set_glue(glue)
    m = mirror
    env = fc.make_env(m.algebra, m.univariate, m.extension_field, m.ntt, m.merkle, m.ip, m.fri)
    n = 1 << logn
    coeffs = np.random.default_rng(logn).integers(0, 18446744069414584321, (3, n // 4), dtype=np.uint64)
    full = eng.ntt(eng.upload(coeffs), logn, root_of_unity(logn), offset=7)
    fri = env.Fri(env.field.generator(), env.field.primitive_nth_root(n), n, 4, 8, env.xfield)
    df = DistFri(glue)
    a, b = full[:, :n // 2].contiguous(), full[:, n // 2:].contiguous()
    for _ in range(2):
        df.prove(fri, a, b, env.ProofStream(), env.Merkle)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    df.prove(fri, a, b, env.ProofStream(), env.Merkle)
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
