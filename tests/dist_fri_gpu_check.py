"""Multi-GPU parity + timing of the sharded FRI prover (stark_brainfuck_b200/dist_fri.py).  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29512 tests/dist_fri_gpu_check.py [--sizes 10,16,20] [--big 24]
Every rank checks its transcript against the golden one (generated from the reference); for --big
sizes (no golden) against the transcript of the same prover on a one-rank group.  Rank 0 prints one
JSON line with wall-clock times (device-synchronised, max over ranks) of the sharded and the
one-GPU run of the same code."""
import argparse
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="10,16")
    ap.add_argument("--big", default="")
    ap.add_argument("--iters", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    solo = None
    for r in range(world):  # a one-rank group per rank: the single-GPU run of the same code
        g = dist.new_group([r])
        if r == rank:
            solo = g
    import frontend_cases as fc
    from stark_brainfuck_b200 import Engine, mirror
    from stark_brainfuck_b200.dist_fri import DistFri
    from stark_brainfuck_b200.glue import Glue
    from util import golden, have_golden, rand_xfe, root_of_unity
    mirror.register()
    eng = Engine(local)
    glue = Glue(mirror.binding, eng)
    mirror.set_glue(glue)
    m = mirror
    env = fc.make_env(m.algebra, m.univariate, m.extension_field, m.ntt, m.merkle, m.ip, m.fri)
    expansion, s = 4, 8
    report = {"n_gpus": world, "cases": {}}

    def run(df, fri, full, G, r):
        n = full.shape[1]
        blk = n // (2 * G)
        a = full[:, r * blk:(r + 1) * blk].contiguous()
        b = full[:, n // 2 + r * blk:n // 2 + (r + 1) * blk].contiguous()
        ps = env.ProofStream()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        top = df.prove(fri, a, b, ps, env.Merkle)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return top, ps, float(dt.item()) * 1e3

    for kind, logs in (("golden", args.sizes), ("big", args.big)):
        for logn in [int(v) for v in logs.split(",") if v]:
            n = 1 << logn
            coeffs = rand_xfe(200 + logn, n // expansion) if kind == "golden" else \
                np.random.default_rng(logn).integers(0, 18446744069414584321, (3, n // expansion), dtype=np.uint64)
            full = eng.ntt(eng.upload(coeffs), logn, root_of_unity(logn), offset=7)  # code/fri.py:33-39 on the device
            fri = env.Fri(env.field.generator(), env.field.primitive_nth_root(n), n, expansion, s, env.xfield)
            # the same blocks through the sharded chain: residue-class LDE (no exchange) + one all-to-all
            from stark_brainfuck_b200.dist import shard_coset_evaluate
            from stark_brainfuck_b200.dist_fri import residues_to_pair_blocks
            if n >= 2 * world * world:
                res = shard_coset_evaluate(eng, eng.upload(coeffs), logn, root_of_unity(logn), 7, rank, world)
                ca, cb = residues_to_pair_blocks(res)
                blk = n // (2 * world)
                same = torch.equal(ca, full[:, rank * blk:(rank + 1) * blk]) and \
                    torch.equal(cb, full[:, n // 2 + rank * blk:n // 2 + (rank + 1) * blk])
                chain_ok = torch.tensor([1 if same else 0], device="cuda")
                dist.all_reduce(chain_ok, op=dist.ReduceOp.MIN)
                report.setdefault("chain_blocks_identical", {})[str(logn)] = bool(chain_ok.item())
            shard, one = DistFri(glue), DistFri(glue, group=solo)
            best = {"sharded_ms": 1e30, "one_gpu_ms": 1e30}
            for _ in range(args.iters):
                top, ps, ms = run(shard, fri, full, world, rank)
                best["sharded_ms"] = min(best["sharded_ms"], ms)
                top1, ps1, ms1 = run(one, fri, full, 1, 0)
                best["one_gpu_ms"] = min(best["one_gpu_ms"], ms1)
            ser, ser1 = ps.serialize(), ps1.serialize()
            ok = top == top1 and ser == ser1
            name = "fri_small.json" if logn <= 12 else "fri_%d.json" % logn
            if kind == "golden" and have_golden(name):
                e = golden(name)
                e = e["gv6"][str(logn)] if logn <= 12 else e
                ok = ok and top == e["top_level_indices"] and len(ps.objects) == e["num_objects"] \
                    and len(ser) == e["transcript_len"] and hashlib.sha256(ser).hexdigest() == e["transcript_sha256"]
                best["golden"] = True
            flag = torch.tensor([1 if ok else 0], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            best["transcripts_identical"] = bool(flag.item())
            best["transcript_len"] = len(ser)
            best["p2p_bytes_rank0"] = shard.exchanged_bytes // args.iters
            report["cases"][str(logn)] = best
            del full
    if rank == 0:
        print(json.dumps(report))
    dist.destroy_process_group()
    bad = [k for k, v in report["cases"].items() if not v["transcripts_identical"]]
    bad += [k for k, v in report.get("chain_blocks_identical", {}).items() if not v]
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
