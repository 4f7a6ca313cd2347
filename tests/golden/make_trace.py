#!/usr/bin/env python3
"""Record the device command stream of BrainfuckStark.prove() (the UNMODIFIED reference under the drop-in, seeded
urandom, host-memory test backend = CPU oracle) into tests/golden/trace_<name>.bin for the GPU replay test
(tests/test_gpu_prove_replay.py).  Authoring container only (needs /root/reference/code).

    python tests/golden/make_trace.py pppp          # "++++", FRI domain 1024
    python tests/golden/make_trace.py hello         # Hello World, 907 cycles, FRI domain 2^17

The recorded proof is checked here: the reference's own verifier accepts it and its hash is the golden one
(tests/golden/bfs.json for "++++"; 540a9a28... for Hello World, the hash of tests/test_e2e_prove.py)."""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REFERENCE_DIR = os.environ.get("B2S_REFERENCE_DIR", "/root/reference/code")

PROGRAMS = {
    "pppp": ("++++", ""),
    "io": ("++[>,.<-]", "ab"),
    "hello": ("++++++++[>++++[>++>+++>+++>+<<<<-]>+>+>->>+[<]<-]>>.>---.+++++++..+++.>>.<-.<.+++.------.--------.>>+.>++.", ""),
    # 8 780 cycles, trace padded to 2^14 rows, FRI domain 2^20 = BASELINE config 5's size
    "big": ("+++++++++++[>+++++++++++[>+++++++++++[>+<-]<-]<-]", ""),
}
# hello: the hash of tests/test_e2e_prove.py; big: the proof the GPU produced with the reference staged next to it
# (profiles/artifacts/r02s_prove_2p20_domain_b200.json), reproduced here by the CPU oracle
EXPECTED = {"hello": "540a9a28053b3195231dc7736163b760d8015a7159b973b85307e45ace4a6f3e",
            "big": "a9ce6dd2406c1439d34c285d4396f8a07aa3c43400c019c2bb2a28e2f4a5ecc9"}


def main(name):
    source, inputs = PROGRAMS[name]
    sys.dont_write_bytecode = True
    sys.path.insert(0, REFERENCE_DIR)
    import trace_backend as tb
    from fake_backend import FakeLib
    from stark_brainfuck_b200 import dropin
    urandom = tb.SeededUrandom(1234)
    rec = tb.Recorder(urandom)
    engine = tb.TracingEngine(FakeLib(), rec)
    dropin.install(REFERENCE_DIR, engine=engine)
    os.urandom = urandom
    import salted_merkle
    salted_merkle.urandom = urandom
    from vm import VirtualMachine
    from brainfuck_stark import BrainfuckStark
    program = VirtualMachine.compile(source)
    running_time, input_symbols, output_symbols = VirtualMachine.run(program, input_data=list(inputs))
    matrices = VirtualMachine.simulate(program, input_data=input_symbols)
    bfs = BrainfuckStark(running_time, len(matrices[1]), program, input_symbols, output_symbols)
    t0 = time.time()
    proof = bfs.prove(program, *matrices)
    dt = time.time() - t0
    dropin.uninstall()
    assert bfs.verify(proof), "the reference verifier rejects the recorded proof"
    sha = hashlib.sha256(proof).hexdigest()
    all_reference = {"pppp": "bfs.json", "io": "bfs_io.json"}  # proofs of the all-Python reference (make_golden.py)
    if name in all_reference:
        assert sha == json.load(open(os.path.join(HERE, all_reference[name])))["proof_sha256"]
    if name in EXPECTED:
        assert sha == EXPECTED[name], sha
    meta = {"program": source, "inputs": inputs, "urandom_seed": 1234, "running_time": running_time,
            "fri_domain_length": bfs.fri.domain.length, "proof_sha256": sha, "proof_len": len(proof),
            "reference_verifier_accepts": True, "recorded_over": "tests/fake_backend.py (oracle/b2s_oracle.c)",
            "prove_seconds_cpu_backend": round(dt, 1)}
    path = os.path.join(HERE, "trace_%s.bin" % name)
    rec.save(path, meta)
    n_calls = sum(1 for e in rec.events if e["op"] == "call")
    print(json.dumps(dict(meta, events=len(rec.events), calls=n_calls, literal_bytes=rec.size,
                          file_bytes=os.path.getsize(path)), indent=1))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "pppp")
