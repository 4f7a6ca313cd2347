"""Device-time budget of one BrainfuckStark.prove at BASELINE config 5's size (trace padded to 2^16, FRI domain
2^20) with SYNTHETIC traces: the device ops the drop-in issues for a proof, in order, with the real constraint
programs of the reference's AIR (tests/golden/air.json).  Not a benchmark contract; it answers "how much of a
proof is device time".  The two salted trees hash the real row leaves (17- and 10-tuples, templates derived from
the mirror's classes as the glue derives them from a proof's sample row).
Run on a GPU box: python profiles/microbench/prove_device_pipeline.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from stark_brainfuck_b200 import Engine, mirror  # noqa: E402
from util import golden, quotient_program, root_of_unity  # noqa: E402

P = 18446744069414584321


def pipeline_ms(eng, log_h=16, log_n=20, reps=3):
    """best-of-`reps` CUDA-event times (ms) of the device ops of one proof; see the module docstring"""
    h, N = 1 << log_h, 1 << log_n
    mirror.register()
    tpl = mirror.binding.xfe_templates(mirror.xfield)
    air = golden("air.json")
    rng = np.random.default_rng(1)
    w, omicron = root_of_unity(log_n), root_of_unity(log_h)
    oinv = pow(omicron, P - 2, P)
    tables = air["tables"]
    traces_b = [eng.upload(rng.integers(0, P, (t["base_width"], h), dtype=np.uint64)) for t in tables]
    traces_x = [eng.upload(rng.integers(0, P, (3 * (t["full_width"] - t["base_width"]), h), dtype=np.uint64))
                for t in tables]
    progs = [[quotient_program(t[k]) for k in ("boundary", "transition", "terminal")] for t in tables]
    rnd = eng.upload(rng.integers(0, P, (3, N // 4), dtype=np.uint64))
    from stark_brainfuck_b200 import marshal
    m = mirror
    fields = [m.algebra.BaseField.main() for _ in tables]
    row_b = tuple([m.binding.make_xfe(5, 6, 7, m.xfield)] +
                  [m.algebra.BaseFieldElement(3 + j, f) for t, f in zip(tables, fields) for j in range(t["base_width"])])
    row_x = tuple(m.binding.make_xfe(5 + j, 6, 7, m.xfield)
                  for j in range(sum(t["full_width"] - t["base_width"] for t in tables)))
    tplb, tplx = marshal.row_template(m.binding, row_b), marshal.row_template(m.binding, row_x)
    frame = marshal.salt_frame(bytes(range(24)))
    salts = eng.upload_bytes(rng.integers(0, 256, (N, 24), dtype=np.uint8))
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def run():
        marks = [("start", ev())]
        marks[0][1].record()

        def mark(name):
            e = ev()
            e.record()
            marks.append((name, e))
        # randomizer codeword + lde / ldex of every table: intt over <omicron>, coset ntt to the FRI domain
        rcw = eng.ntt(rnd, log_n, w, offset=7)
        base_cw, ext_cw = [], []
        for tb, tx in zip(traces_b, traces_x):
            base_cw.append(eng.ntt(eng.ntt(tb, log_h, omicron, inverse=True), log_n, w, offset=7))
            ext_cw.append(eng.ntt(eng.ntt(tx, log_h, omicron, inverse=True), log_n, w, offset=7))
        mark("lde+ldex (46 columns)")
        # the two salted trees over zipped rows (code/brainfuck_stark.py:178-180, :197-199): randomizer + 16 base
        # columns, then the 10 extension columns; row templates derived from the mirror's classes as the glue does
        base_planes = [rcw[0], rcw[1], rcw[2]] + [b[i] for b in base_cw for i in range(b.shape[0])]
        eng.merkle_rows(base_planes, tplb.modes, tplb.tpl, tplb.seg_off, N, salts, frame[0], frame[1])
        ext_planes = [x[i] for x in ext_cw for i in range(x.shape[0])]
        eng.merkle_rows(ext_planes, tplx.modes, tplx.tpl, tplx.seg_off, N, salts, frame[0], frame[1])
        mark("2 salted row trees (17- and 10-tuples)")
        quotients, assembled = [], []
        for t, bc, xc in zip(tables, base_cw, ext_cw):
            W = t["full_width"]
            cw = torch.zeros((W, 3, N), dtype=torch.int64, device=eng.device)
            cw[:t["base_width"], 0] = bc  # lifted base columns (code/processor_table.py:421)
            cw[t["base_width"]:] = xc.reshape(-1, 3, N)
            assembled.append(cw)
        mark("table planes assembled (lifted base columns + extension columns)")
        for t, cw, pr in zip(tables, assembled, progs):
            flags = [j < t["base_width"] for j in range(t["full_width"])]  # what Glue._table_planes knows
            for kind in (1, 2, 3):
                out, _ = eng.quotients(cw, N // h, *pr[kind - 1], kind, h, oinv, 7, w, base_columns=flags,
                                       check_zerofier=False)  # offset^N != 1: decided on the host, as the glue does
                quotients.append(out)
        mark("quotients (47 constraints, 5 tables, 15 calls)")
        cols = [rcw] + [b[i:i + 1] for b in base_cw for i in range(b.shape[0])]
        cols += [x[3 * i:3 * i + 3] for x in ext_cw for i in range(x.shape[0] // 3)]
        cols += [q[i] for q in quotients for i in range(q.shape[0])]
        nc = len(cols)
        wa = rng.integers(0, P, (nc, 3), dtype=np.uint64)
        wb = rng.integers(0, P, (nc, 3), dtype=np.uint64)
        wb[0] = 0
        shifts = rng.choice([0, 1, h - 1, 2 * h, 3 * h - 2, 3 * h], nc)
        comb = eng.combination(cols, wa, wb, shifts, N, 7, w)
        mark("nonlinear combination (%d columns)" % nc)
        eng.merkle_field(comb, tpl)
        cw, n, ww, off = comb, N, w, 7
        while n // 2 > 4:
            cw, _ = eng.fri_fold(cw, [3, 5, 7], off, ww, tpl)
            n //= 2
            ww, off = ww * ww % P, off * off % P
        mark("combination tree + FRI commit")
        torch.cuda.synchronize()
        # a mark's name belongs to the interval that ENDS at its event
        return [(marks[i + 1][0], marks[i][1].elapsed_time(marks[i + 1][1])) for i in range(len(marks) - 1)] + \
               [("total", marks[0][1].elapsed_time(marks[-1][1]))]

    run()
    best = None
    for _ in range(reps):
        r = run()
        if best is None or r[-1][1] < best[-1][1]:
            best = r
    return {"trace_rows": h, "fri_domain": N, "ms": {name: round(ms, 3) for name, ms in best}}


def main():
    print(json.dumps(pipeline_ms(Engine(0)), indent=1))


if __name__ == "__main__":
    main()
