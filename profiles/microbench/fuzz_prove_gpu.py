#!/usr/bin/env python3
"""Random terminating Brainfuck programs through the UNMODIFIED BrainfuckStark.prove() under the drop-in, once over
libb2s.so on cuda:0 and once over the host-memory test backend (CPU oracle), same seeded urandom: the two proofs must be
byte-identical and the reference's verifier must accept.  Needs the staged reference (baseline/_ref/code) and a GPU.

    python profiles/microbench/fuzz_prove_gpu.py [n_programs] [seed] [nested]      -> gpurun_out/fuzz_prove_gpu.json
"""
import json
import os
import random
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
STAGED = os.environ.get("B2S_REFERENCE_DIR", os.path.join(ROOT, "baseline", "_ref", "code"))


def program(R):
    """(source, input string): straight-line pieces and counted loops `k+ [ > body < - ]`, which always terminate"""
    src, n_in = "", 0
    for _ in range(R.randrange(1, 5)):
        kind = R.randrange(6)
        if kind == 0:
            src += "+" * R.randrange(1, 6)
        elif kind == 1:
            src += "+" * R.randrange(0, 3) + "." * R.randrange(1, 3)
        elif kind == 2:
            src += ","
            n_in += 1
            if R.random() < 0.5:
                src += "."
        elif kind == 3:
            src += ">" + "+" * R.randrange(1, 4) + "<"
        else:
            body = ""
            for _ in range(R.randrange(1, 4)):
                b = R.randrange(4)
                if b == 0:
                    body += "+"
                elif b == 1:
                    body += "."
                elif b == 2:
                    body += ","
                    n_in += 8  # upper bound on the iterations below
                else:
                    body += ">+<"
            src += "[-]" + "+" * R.randrange(1, 8) + "[>" + body + "<-]"
    return src, "".join(chr(R.randrange(1, 120)) for _ in range(n_in))


def nested_program(R):
    """two nested counted loops around a small body: hundreds to a few thousand cycles (FRI domains 2^14 ... 2^18)"""
    a, b = R.randrange(3, 12), R.randrange(3, 12)
    body = "".join(R.choice(["+", "+.", ">+<", "-", "++"]) for _ in range(R.randrange(1, 4)))
    tail = R.choice(["", ".", ">.<", ",."])
    return "+" * a + "[>" + "+" * b + "[>" + body + "<-]<-]" + tail, "z" if "," in tail else ""


def run(backend, source, inputs):
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "r.json")
        p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "e2e_prove_dropin.py"), backend, out, source, inputs],
                           env=dict(os.environ, B2S_REFERENCE_DIR=STAGED), stdout=subprocess.DEVNULL,
                           stderr=subprocess.PIPE, text=True, timeout=1200)
        if p.returncode != 0:
            return {"error": p.stderr.strip().splitlines()[-1] if p.stderr.strip() else "exit %d" % p.returncode}
        return json.load(open(out))


def consumed_inputs(src, inp):
    """the input symbols the program actually reads (the prover is constructed with exactly those), or None when the
    reference's own VM rejects the program (e.g. output of a cell that was never written: KeyError in code/vm.py)"""
    sys.path.insert(0, STAGED)
    from vm import VirtualMachine
    try:
        prog = VirtualMachine.compile(src)
        VirtualMachine.run(prog, input_data=list(inp))
        mats = VirtualMachine.simulate(prog, input_data=list(inp))
    except Exception:
        return None
    return inp[:len(mats[3])]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    R = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
    rows, bad = [], 0
    k = 0
    make = nested_program if len(sys.argv) > 3 and sys.argv[3] == "nested" else program
    while k < n:
        src, inp = make(R)
        inp = consumed_inputs(src, inp)
        if inp is None:
            continue
        k += 1
        g, c = run("gpu", src, inp), run("fake", src, inp)
        ok = ("error" not in g and "error" not in c and g["proof_sha256"] == c["proof_sha256"]
              and g["reference_verifier_accepts"] is True)
        # the same exception over both backends comes from the reference's own code or the host glue, not from the GPU
        # (e.g. code/io_table.py:54-56 refuses a non-empty output table whose symbols are all zero)
        same_error = "error" in g and g.get("error") == c.get("error")
        bad += not (ok or same_error)
        rows.append({"program": src, "inputs": inp, "ok": ok, "same_error_on_both_backends": same_error,
                     "domain": g.get("fri_domain_length"),
                     "gpu_s": g.get("prove_seconds"), "cpu_backend_s": c.get("prove_seconds"),
                     "gpu_error": g.get("error"), "cpu_error": c.get("error")})
        print(k - 1, ok, src, g.get("fri_domain_length"), g.get("prove_seconds"), g.get("error"), c.get("error"), flush=True)
    res = {"programs": n, "failed": bad, "cases": rows}
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        json.dump(res, open(os.path.join(out, "fuzz_prove_gpu.json"), "w"), indent=1)
    print("failed", bad, "of", n)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
