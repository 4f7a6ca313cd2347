"""N > 1 host logic on CPU: world_size-2 (and 4) gloo process groups over the host-memory test
backend.  Checks the four-step sharded NTT (one all-to-all) against the CPU oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, log_n, log_n1, inverse, ret):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fake_backend import fake_engine
        from stark_brainfuck_b200.dist import DistNTT, scatter_columns
        from util import rand_bfe, root_of_unity
        eng = fake_engine()
        x = rand_bfe(7000 + log_n, 1 << log_n)
        w = root_of_unity(log_n)
        d = DistNTT(eng, exchange="nccl")
        local = eng.upload(scatter_columns(x, log_n, log_n1, rank, world))
        out = d.transform(local, log_n, w, inverse=inverse, log_n1=log_n1)
        ret[rank] = eng.download(out).copy()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,log_n,log_n1,inverse", [(2, 10, 5, False), (2, 11, 5, False), (2, 10, 5, True),
                                                        (4, 8, 4, False), (2, 6, 1, False)])
def test_sharded_ntt_matches_oracle(world, log_n, log_n1, inverse):
    from oracle import oracle as orc
    from stark_brainfuck_b200.dist import assemble_output
    from util import rand_bfe, root_of_unity
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() + log_n * 7 + world) % 2000
    mp.spawn(_worker, args=(world, port, log_n, log_n1, inverse, ret), nprocs=world, join=True)
    x = rand_bfe(7000 + log_n, 1 << log_n)
    w = root_of_unity(log_n)
    got = assemble_output([ret[r] for r in range(world)], log_n, log_n1)
    ref = orc.intt(w, x) if inverse else orc.ntt(w, x)
    assert np.array_equal(got, ref)


def test_shard_units():
    from stark_brainfuck_b200.dist import shard_units
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 7, 46, 64):
            got = [i for r in range(world) for i in shard_units(n, r, world)]
            assert got == list(range(n))
            sizes = [len(shard_units(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
