"""Multi-GPU parity + timing of the sharded four-step NTT.  Launch (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_gpu_check.py [--log-n 20] [--exchange nccl,p2p]
Rank 0 gathers the shards, compares with the CPU oracle bit for bit and prints one JSON line."""
import argparse
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--exchange", default="nccl,p2p")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle import oracle as orc
    from stark_brainfuck_b200 import Engine
    from stark_brainfuck_b200.dist import DistNTT, assemble_output, scatter_columns
    from util import rand_bfe, root_of_unity
    eng = Engine(local)
    log_n = args.log_n
    log_n1 = log_n // 2
    x = rand_bfe(7000 + log_n, 1 << log_n)
    w = root_of_unity(log_n)
    ref = orc.ntt(w, x) if rank == 0 else None
    report = {"n_gpus": world, "log_n": log_n}
    for ex in args.exchange.split(","):
        try:
            d = DistNTT(eng, exchange=ex)
            loc = eng.upload(scatter_columns(x, log_n, log_n1, rank, world))
            out = d.transform(loc, log_n, w, log_n1=log_n1)
            back = d.transform(out, log_n, w, inverse=True, log_n1=log_n - log_n1)
            torch.cuda.synchronize()
            ok_rt = bool(torch.equal(back, loc))
            parts = [torch.empty_like(out) for _ in range(world)]
            dist.all_gather(parts, out)
            ok = None
            if rank == 0:
                got = assemble_output([eng.download(p) for p in parts], log_n, log_n1)
                ok = bool(np.array_equal(got, ref))
            # timing: device events, max over ranks
            for _ in range(3):
                d.transform(loc, log_n, w, log_n1=log_n1)
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                d.transform(loc, log_n, w, log_n1=log_n1)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / args.iters], device=eng.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            report[ex] = {"matches_oracle": ok, "roundtrip_exact": ok_rt, "ms_per_transform": float(t[0])}
        except Exception as e:  # report, do not hang the other ranks
            report[ex] = {"error": repr(e)[:300]}
    # sharded coset evaluation (LDE): every rank one residue class, no exchange
    try:
        from stark_brainfuck_b200.dist import assemble_residues, shard_coset_evaluate
        from util import rand_xfe
        for name, expansion in (("lde_expansion4", 4), ("lde_expansion1_folded", 1)):
            c = rand_xfe(9100 + log_n, (1 << log_n) // expansion)
            cd = eng.upload(c)
            out = shard_coset_evaluate(eng, cd, log_n, w, 7, rank, world)
            parts = [torch.empty_like(out) for _ in range(world)]
            dist.all_gather(parts, out)
            ok = None
            if rank == 0:
                got = assemble_residues([eng.download(p) for p in parts])
                ok = bool(np.array_equal(got, orc.coset_evaluate(7, w, c, 1 << log_n)))
            for _ in range(3):
                shard_coset_evaluate(eng, cd, log_n, w, 7, rank, world)
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                shard_coset_evaluate(eng, cd, log_n, w, 7, rank, world)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / args.iters], device=eng.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            report[name] = {"matches_oracle": ok, "ms_per_evaluation": float(t[0])}
    except Exception as e:
        report["lde"] = {"error": repr(e)[:300]}
    # column-sharded planes -> row-sharded planes (one all-to-all), and point-range sharded evaluation
    try:
        from stark_brainfuck_b200.dist import columns_to_rows, shard_eval_points
        from util import rand_bfe as rb
        nn, c = 1 << 12, 3
        mine = np.stack([rb(500 + rank * c + j, nn) for j in range(c)])
        rows = eng.download(columns_to_rows(eng.upload(mine)))
        full = np.stack([rb(500 + j, nn) for j in range(world * c)])
        ok_rows = bool(np.array_equal(rows, full[:, rank * (nn // world):(rank + 1) * (nn // world)]))
        pts = rb(78, 64).reshape(1, 64)
        vals = eng.download(shard_eval_points(eng, eng.upload(rb(77, 300)), eng.upload(pts), rank, world))
        lo = rank * (64 // world)
        ok_pts = bool(np.array_equal(vals[0], orc.eval_points(rb(77, 300), pts[0])[lo:lo + 64 // world]))
        flag = torch.tensor([1 if ok_rows and ok_pts else 0], device=eng.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        report["columns_to_rows_and_point_shards_ok"] = bool(flag.item())
    except Exception as e:
        report["columns_to_rows_and_point_shards_ok"] = repr(e)[:300]
    if rank == 0:
        print(json.dumps(report))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
