#!/usr/bin/env python3
"""bench.py -- headline measurement of the hot path (contract: see the task statement / DESIGN.md).

Workload (BASELINE.json configs[1]): one step = forward + inverse radix-2 NTT of one
2^20-element BaseField vector (intt(ntt(x)) == x bit-exact), synthetic uniform-random
coefficients.  With N GPUs every rank transforms its own vector (independent units, no
data-path collective): weak scaling.

metric  = algorithmic field multiplications per second, (n/2)*log2(n) per NTT plus n for the
          n^-1 scaling of the inverse, whole job over all ranks.
value   = inputs resident in HBM; every timed step works on its own (input, output, round-trip)
          buffer triple out of a ring of 20 (480 MB > the 126 MB L2), so inputs are never L2-resident;
          CUDA events on the launching stream.
e2e     = the same step through the C-ABI host-buffer entry point (b2s_ntt_host): pinned host
          input -> H2D -> kernels -> D2H, per transform.
roofline= forward transform (2 launches of ntt4_pass_kernel): 16 B/element algorithmic HBM bytes
          over its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs.
--impl reference times the CPU oracle port (oracle/b2s_oracle.c, single thread -- the reference is
single-threaded Python) on the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

LOG_N = 20
P_MOD = 18446744069414584321
METRIC = "ntt_field_mul_per_s"
UNIT = "field-mul/s"


def muls_per_step(log_n):
    n = 1 << log_n
    return 2 * (n // 2) * log_n + n


def root_of_unity(log_n):
    p = 18446744069414584321
    r = 1753635133440165772
    for _ in range(32 - log_n):
        r = r * r % p
    return r


def synth(seed, n):
    import numpy as np
    rng = np.random.default_rng(seed)
    p = 18446744069414584321
    x = rng.integers(0, p, size=n, dtype=np.uint64, endpoint=False)
    return x


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [t.strip() for t in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args):
    """CPU arm: the oracle port of the reference's ntt/intt (code/ntt.py:4-42) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import oracle as orc
    n = 1 << LOG_N
    w = root_of_unity(LOG_N)
    x = synth(1, n)
    for _ in range(max(args.warmup, 1)):
        y = orc.ntt(w, x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        y = orc.ntt(w, x)
        z = orc.intt(w, y)
    dt = time.perf_counter() - t0
    assert np.array_equal(z, x)
    ms = dt / args.steps * 1e3
    val = muls_per_step(LOG_N) / (ms / 1e3)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": "ntt+intt round trip, 2^20 BaseField (BASELINE configs[1])", "log_n": LOG_N},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": "%d full round trips of the 2^20 workload, oracle/b2s_oracle.c single thread "
                                   "(the Python reference needs 338 s per forward transform)" % args.steps},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from stark_brainfuck_b200 import Engine
    eng = Engine(local)
    dev = eng.device
    n = 1 << LOG_N
    w = root_of_unity(LOG_N)
    NBUF = 20  # ring of (x, y, z) triples: 20 * 24 MB = 480 MB, nothing survives in the 126 MB L2
    x_np = synth(1 + rank, n)
    xs = [eng.upload(x_np if i == 0 else synth(1000 + 97 * rank + i, n)) for i in range(NBUF)]
    ys = [eng.empty(1, n) for _ in range(NBUF)]
    zs = [eng.empty(1, n) for _ in range(NBUF)]

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    for i in range(max(args.warmup, 3)):
        eng.ntt(xs[i % NBUF], LOG_N, w, out=ys[i % NBUF])
        eng.ntt(ys[i % NBUF], LOG_N, w, inverse=True, out=zs[i % NBUF])
    torch.cuda.synchronize(dev)
    assert torch.equal(zs[0], xs[0]), "intt(ntt(x)) != x"
    ref_check = None
    if rank == 0:
        from oracle import oracle as orc  # the checker, outside the timed region
        ref_check = bool(np.array_equal(eng.download(ys[0])[0], orc.ntt(w, x_np)))
        assert ref_check, "GPU ntt differs from the CPU oracle"

    # ---- device-resident timing -----------------------------------------------------
    K = args.steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    launches0 = eng.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    for k in range(K):
        i = (k + 3) % NBUF
        ev[k][0].record()
        eng.ntt(xs[i], LOG_N, w, out=ys[i])
        ev[k][1].record()
        eng.ntt(ys[i], LOG_N, w, inverse=True, out=zs[i])
        ev[k][2].record()
    barrier()
    launches = eng.launch_count() - launches0
    for k in range(min(K, NBUF)):
        i = (k + 3) % NBUF
        assert torch.equal(zs[i], xs[i]), "intt(ntt(x)) != x in timed step %d" % k
    fwd_ms = [e[0].elapsed_time(e[1]) for e in ev]
    tot_ms = [e[0].elapsed_time(e[2]) for e in ev]
    ms_step = sum(tot_ms) / K
    fwd = sum(fwd_ms) / K
    t = torch.tensor([ms_step, fwd], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step, fwd = float(t[0]), float(t[1])

    # ---- end to end through host buffers -------------------------------------------
    h_x = torch.from_numpy(x_np.view(np.int64).reshape(1, n).copy()).pin_memory()
    h_y = torch.empty((1, n), dtype=torch.int64).pin_memory()
    h_z = torch.empty((1, n), dtype=torch.int64).pin_memory()
    for _ in range(3):
        eng.ntt_host(h_x, LOG_N, w, h_out=h_y)
        eng.ntt_host(h_y, LOG_N, w, inverse=True, h_out=h_z)
    assert torch.equal(h_z, h_x)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        eng.ntt_host(h_x, LOG_N, w, h_out=h_y)
        eng.ntt_host(h_y, LOG_N, w, inverse=True, h_out=h_z)
    torch.cuda.synchronize(dev)
    e2e_ms = (time.perf_counter() - t0) / K * 1e3
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t[0])
    clocks = sampler.stop() if sampler else None

    # ---- extra: batched planes (launch overheads amortised), informational ----------
    extra = {}
    try:
        q = 32
        xb = torch.randint(0, 2 ** 62, (q, n), dtype=torch.int64, device=dev)
        yb = eng.empty(q, n)
        eng.ntt(xb, LOG_N, w, out=yb)
        ms_b, _ = eng.ntt_timed(xb, LOG_N, w, out=yb, iters=5)
        extra["batched_32_planes_fwd_ms"] = ms_b
        extra["batched_32_planes_hbm_gbs"] = 16.0 * n * q / (ms_b / 1e3) / 1e9
    except Exception as e:  # informational only
        extra["batched_error"] = str(e)
    # ---- extra: independent round trips issued round-robin on four streams (what a caller with many single
    # vectors and no batch to hand over gets: the kernels of different transforms fill each other's idle SMs)
    try:
        streams = [torch.cuda.Stream(device=dev) for _ in range(4)]
        cur = torch.cuda.current_stream(dev)
        for rep in range(2):  # first repetition warms the per-stream pools
            start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            start.record()
            for s_ in streams:
                s_.wait_event(start)
            KP = 40
            for k in range(KP):
                i = k % NBUF
                with torch.cuda.stream(streams[k % 4]):
                    eng.ntt(xs[i], LOG_N, w, out=ys[i])
                    eng.ntt(ys[i], LOG_N, w, inverse=True, out=zs[i])
            for s_ in streams:
                cur.wait_stream(s_)
            end.record()
            torch.cuda.synchronize(dev)
            extra["four_streams_ms_per_round_trip"] = start.elapsed_time(end) / KP
        assert torch.equal(zs[5], xs[5])
    except Exception as e:  # informational only
        extra["four_streams_error"] = str(e)
    # ---- extra: the Merkle / FRI half of the path at a 2^20 domain (BASELINE configs[3] with SURVEY D8's
    # fix), device side: round-0 tree, then fold + next tree per round with fixed challenges ------------
    try:
        if rank == 0:
            from stark_brainfuck_b200 import mirror
            mirror.register()
            tpl = mirror.binding.xfe_templates(mirror.xfield)
            rng = np.random.default_rng(5)
            cw0 = eng.upload(rng.integers(0, P_MOD, size=(3, n), dtype=np.uint64, endpoint=False))
            best = None
            for _ in range(4):
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record()
                eng.merkle_field(cw0, tpl)
                e1.record()
                cw, N, ww, off = cw0, n, w, 7
                while N // 2 > 4:
                    cw, _nodes = eng.fri_fold(cw, [3, 5, 7], off, ww, tpl)
                    N //= 2
                    ww, off = ww * ww % P_MOD, off * off % P_MOD
                e2.record()
                torch.cuda.synchronize(dev)
                t = (e0.elapsed_time(e1), e0.elapsed_time(e2))
                best = t if best is None or t[1] < best[1] else best
            extra["merkle_2p20_xfe_leaves_ms"] = best[0]
            extra["merkle_2p20_hbm_gbs"] = 152.0 * n / (best[0] / 1e3) / 1e9
            extra["fri_commit_2p20_expansion4_ms"] = best[1]
            extra["fri_commit_2p20_hbm_gbs"] = 328.0 * n / (best[1] / 1e3) / 1e9
            # the whole proof through the reference-named front end (code/fri.py:178-199), codeword already on
            # the device, host Fiat-Shamir loop and transcript objects included: wall clock
            from stark_brainfuck_b200.glue import DeviceCodeword, Glue
            glue = Glue(mirror.binding, eng)
            old_glue = mirror._glue
            mirror.set_glue(glue)
            fri = mirror.fri.Fri(mirror.field.generator(), mirror.field.primitive_nth_root(n), n, 4, 8, mirror.xfield)
            cwx = eng.ntt(eng.upload(rng.integers(0, P_MOD, size=(3, n // 4), dtype=np.uint64, endpoint=False)), LOG_N, w,
                          offset=7)
            best_p = None
            for _ in range(3):
                ps = mirror.ip.ProofStream()
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                fri.prove(DeviceCodeword(glue, cwx, mirror.xfield), ps)
                torch.cuda.synchronize(dev)
                dt = (time.perf_counter() - t0) * 1e3
                best_p = dt if best_p is None else min(best_p, dt)
            extra["fri_prove_2p20_expansion4_s8_wall_ms"] = best_p
            extra["fri_prove_transcript_bytes"] = len(ps.serialize())
            mirror.set_glue(old_glue)
            mirror.unregister()
    except Exception as e:  # informational only
        extra["fri_error"] = str(e)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    muls = muls_per_step(LOG_N)
    value = world * muls / (ms_step / 1e3)
    e2e_val = world * muls / (e2e_ms / 1e3)
    peak, which = peaks()
    achieved = 16.0 * n / (fwd / 1e3) / 1e9
    traffic = None
    try:  # DRAM bytes of the two pass launches from the committed `ncu --set full` capture
        with open(os.path.join(ROOT, "profiles", "ntt_traffic.json")) as f:
            traffic = int(json.load(f)["dram_bytes_per_transform"])
    except Exception:
        pass

    # ---- CPU baseline: the oracle port on this box's host cores (bounded sample) ------
    from oracle import oracle as orc
    cpu_steps = 8
    t0 = time.perf_counter()
    for _ in range(cpu_steps):
        yy = orc.ntt(w, x_np)
        zz = orc.intt(w, yy)
    cpu_dt = (time.perf_counter() - t0) / cpu_steps
    assert np.array_equal(zz, x_np)

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": "ntt+intt round trip, 2^20 BaseField vector per GPU (BASELINE configs[1])",
                   "log_n": LOG_N, "l2": "inputs larger than L2: ring of 20 (in, out, back) buffer triples = 480 MB",
                   "parity": "intt(ntt(x)) == x bit-exact; ntt == CPU oracle: %s" % ref_check},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": 2 * 8 * n,
                "d2h_bytes_per_step": 2 * 8 * n,
                "path": "b2s_ntt_host (C ABI, pinned host buffers) forward then inverse"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": which,
                     "kernel": "ntt4_pass_kernel x2 (forward 2^20 transform, input not L2-resident)",
                     "algorithmic_bytes": 16 * n, "duration_ms": fwd},
        "cpu_baseline": {"value": muls / cpu_dt, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": "%d round trips of the same 2^20 workload, oracle/b2s_oracle.c, 1 thread of %d"
                                   % (cpu_steps, os.cpu_count() or 0)},
        "extra": extra,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None,
                    help="timed steps (default: 500 round trips on the GPU arm -- long enough for a few clock samples; 8 on the CPU reference arm)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 8 if args.impl == "reference" else 500
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
