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


def _fri_worker(rank, world, port, logn, expansion, s, seed, below, ret):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import frontend_cases as fc
        from fake_backend import fake_engine
        from stark_brainfuck_b200 import mirror
        from stark_brainfuck_b200.dist_fri import DistFri, scatter_pair_blocks
        from stark_brainfuck_b200.glue import Glue
        mirror.register()
        glue = Glue(mirror.binding, fake_engine())
        mirror.set_glue(glue)
        m = mirror
        env = fc.make_env(m.algebra, m.univariate, m.extension_field, m.ntt, m.merkle, m.ip, m.fri)
        n = 1 << logn
        fri = env.Fri(env.field.generator(), env.field.primitive_nth_root(n), n, expansion, s, env.xfield)
        _, planes = fc.fri_input(env, logn, expansion, seed)
        a, b = scatter_pair_blocks(planes, rank, world)
        df = DistFri(glue)
        ps = env.ProofStream()
        top = df.prove(fri, glue.engine.upload(a), glue.engine.upload(b), ps, env.Merkle, replicate_below=below)
        ser = ps.serialize()
        ok = None
        if rank == 0 and logn <= 8:  # the (mirror of the) reference's verifier accepts the sharded proof
            cw = [fc.X(env, *[int(planes[j][i]) for j in range(3)]) for i in range(n)]
            ok = fri.verify(env.ProofStream().deserialize(ser), env.Merkle(cw).root())
        ret[rank] = (top, len(ps.objects), [o.hex() for o in ps.objects if isinstance(o, bytes)], ser, ok,
                     df.exchanged_bytes)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,logn,below", [(2, 6, 1), (2, 8, 1), (4, 8, 1), (2, 10, 16), (4, 10, 1), (8, 10, 1),
                                              (4, 10, 1 << 12), (8, 8, 4)])
def test_sharded_fri_transcript_matches_golden(world, logn, below):
    """one codeword over `world` ranks: every rank's transcript is the reference's, byte for byte"""
    import hashlib
    from util import golden
    e = golden("fri_small.json")["gv6"][str(logn)]
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + (os.getpid() + logn * 11 + world) % 2000
    mp.spawn(_fri_worker, args=(world, port, logn, 4, 8, 200 + logn, below, ret), nprocs=world, join=True)
    for r in range(world):
        top, nobj, roots, ser, ok, _ = ret[r]
        assert top == e["top_level_indices"] and nobj == e["num_objects"] and roots == e["round_roots"]
        assert len(ser) == e["transcript_len"] and hashlib.sha256(ser).hexdigest() == e["transcript_sha256"]
    if logn <= 8:
        assert ret[0][4] is True
    # point-to-point traffic per rank: half a block per butterfly round, blocks halving (none in round 0)
    assert ret[0][5] < 3 * 8 * (1 << logn) // (2 * world)
    if below == 1:
        assert ret[0][5] > 0


def _lde_worker(rank, world, port, log_n, expansion, xfe, sharded_coeffs, ret):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fake_backend import fake_engine
        from stark_brainfuck_b200.dist import shard_coset_evaluate
        from util import rand_bfe, rand_xfe, root_of_unity
        eng = fake_engine()
        m = (1 << log_n) // expansion
        c = rand_xfe(9100 + log_n, m) if xfe else rand_bfe(9100 + log_n, m).reshape(1, m)
        if sharded_coeffs:
            c = c[:, rank * (m // world):(rank + 1) * (m // world)]
        out = shard_coset_evaluate(eng, eng.upload(c), log_n, root_of_unity(log_n), 7, rank, world,
                                   gather_coefficients=sharded_coeffs)
        ret[rank] = eng.download(out).copy()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,log_n,expansion,xfe,sharded_coeffs", [
    (2, 10, 4, False, False), (4, 10, 4, True, False), (8, 10, 4, False, False), (8, 9, 2, True, False),
    (4, 8, 1, False, False), (2, 10, 4, True, True), (4, 6, 4, False, True)])
def test_sharded_coset_evaluate_matches_oracle(world, log_n, expansion, xfe, sharded_coeffs):
    """every rank evaluates one residue class of the LDE; with more coefficients than local
    points (expansion < world) they are scaled and folded first"""
    from oracle import oracle as orc
    from stark_brainfuck_b200.dist import assemble_residues
    from util import rand_bfe, rand_xfe, root_of_unity
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 33500 + (os.getpid() + log_n * 13 + world + 3 * expansion) % 2000
    mp.spawn(_lde_worker, args=(world, port, log_n, expansion, xfe, sharded_coeffs, ret), nprocs=world, join=True)
    n = 1 << log_n
    m = n // expansion
    c = rand_xfe(9100 + log_n, m) if xfe else rand_bfe(9100 + log_n, m)
    ref = orc.coset_evaluate(7, root_of_unity(log_n), c, n)
    got = assemble_residues([ret[r] for r in range(world)])
    assert np.array_equal(got if xfe else got[0], ref)


def _rows_worker(rank, world, port, ret):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fake_backend import fake_engine
        from stark_brainfuck_b200.dist import columns_to_rows, shard_eval_points
        from util import rand_bfe, rand_xfe
        eng = fake_engine()
        n, c = 64, 3
        mine = np.stack([rand_bfe(500 + rank * c + j, n) for j in range(c)])  # this rank's whole planes
        rows = columns_to_rows(eng.upload(mine))
        coeffs, pts = rand_xfe(77, 40), rand_bfe(78, 21).reshape(1, 21)
        vals = shard_eval_points(eng, eng.upload(coeffs), eng.upload(pts), rank, world)
        ret[rank] = (eng.download(rows).copy(), eng.download(vals).copy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_columns_to_rows_and_sharded_point_evaluation(world):
    from oracle import oracle as orc
    from util import rand_bfe, rand_xfe
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_rows_worker, args=(world, 37100 + (os.getpid() * 7 + world) % 1500, ret), nprocs=world, join=True)
    n, c = 64, 3
    full = np.stack([rand_bfe(500 + j, n) for j in range(world * c)])  # plane order = rank order
    for r in range(world):
        assert np.array_equal(ret[r][0], full[:, r * (n // world):(r + 1) * (n // world)])
    got = np.concatenate([ret[r][1] for r in range(world)], axis=1)
    lifted = np.zeros((3, 21), dtype=np.uint64)  # the oracle takes extension-field points for extension coefficients
    lifted[0] = rand_bfe(78, 21)
    assert np.array_equal(got, orc.eval_points(rand_xfe(77, 40), lifted))


def _lde_fri_worker(rank, world, port, logn, ret):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import frontend_cases as fc
        from fake_backend import fake_engine
        from stark_brainfuck_b200 import mirror
        from stark_brainfuck_b200.dist import shard_coset_evaluate
        from stark_brainfuck_b200.dist_fri import DistFri, residues_to_pair_blocks
        from stark_brainfuck_b200.glue import Glue
        from util import rand_xfe, root_of_unity
        mirror.register()
        glue = Glue(mirror.binding, fake_engine())
        mirror.set_glue(glue)
        m = mirror
        env = fc.make_env(m.algebra, m.univariate, m.extension_field, m.ntt, m.merkle, m.ip, m.fri)
        n = 1 << logn
        fri = env.Fri(env.field.generator(), env.field.primitive_nth_root(n), n, 4, 8, env.xfield)
        coeffs = rand_xfe(200 + logn, n // 4)  # the polynomial behind tests/golden/fri_small.json
        shard = shard_coset_evaluate(glue.engine, glue.engine.upload(coeffs), logn, root_of_unity(logn), 7, rank, world)
        a, b = residues_to_pair_blocks(shard)
        ps = env.ProofStream()
        top = DistFri(glue).prove(fri, a, b, ps, env.Merkle, replicate_below=4)
        ret[rank] = (top, ps.serialize())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,logn", [(2, 8), (4, 10)])
def test_sharded_lde_feeds_sharded_fri(world, logn):
    """polynomial -> sharded coset evaluation (no exchange) -> one all-to-all into pair blocks -> sharded FRI proof:
    the transcript is the golden one the reference produced from the same polynomial"""
    import hashlib
    from util import golden
    e = golden("fri_small.json")["gv6"][str(logn)]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_lde_fri_worker, args=(world, 38300 + (os.getpid() * 3 + world + logn) % 1500, logn, ret), nprocs=world,
             join=True)
    for r in range(world):
        top, ser = ret[r]
        assert top == e["top_level_indices"] and hashlib.sha256(ser).hexdigest() == e["transcript_sha256"]
