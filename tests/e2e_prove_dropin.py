"""BASELINE config 5 at the scale the reference itself can finish: BrainfuckStark.prove() of the
UNMODIFIED reference with the drop-in installed (hot path through the C-ABI surface), seeded
urandom as in SURVEY App. C GV7.  Acceptance = the reference's own verifier accepts, and the
proof is byte-identical to the all-reference proof recorded in tests/golden/bfs.json.

  python tests/e2e_prove_dropin.py [fake|gpu] [out.json] [program | "hello"] [input symbols] [golden file]

`fake` runs the engine calls on the host-memory test backend (authoring container, no GPU);
`gpu` uses libb2s.so on cuda:0 (needs a box that has both a GPU and the reference checkout).
~10 minutes: 93 % of prove() is the reference's pure-Python quotient evaluation,
which an unmodified prove() keeps."""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REFERENCE_DIR = os.environ.get("B2S_REFERENCE_DIR", "/root/reference/code")


HELLO = ("++++++++[>++++[>++>+++>+++>+<<<<-]>+>+>->>+[<]<-]>>.>---.+++++++..+++.>>.<-.<.+++.------.--------.>>+.>++.")


def main(backend="fake", out=None, source="++++", inputs="", golden_name="bfs.json"):
    if source == "hello":
        source = HELLO
    sys.dont_write_bytecode = True
    sys.path.insert(0, REFERENCE_DIR)
    from stark_brainfuck_b200 import dropin
    if backend == "fake":
        from fake_backend import fake_engine
        engine = fake_engine()
    else:
        from stark_brainfuck_b200 import Engine
        engine = Engine(0)
    glue = dropin.install(REFERENCE_DIR, engine=engine)
    from trace_backend import SeededUrandom
    import salted_merkle
    from vm import VirtualMachine
    from brainfuck_stark import BrainfuckStark
    if os.environ.get("B2S_E2E_WARMUP"):
        # an untimed proof of the four-cycle program first: the first use of every kernel in a process loads its
        # module (0.6 s on a fresh box), which is not the proof's time.  The timed run below starts from the same seed.
        os.urandom = salted_merkle.urandom = SeededUrandom(99)
        wp = VirtualMachine.compile("++++")
        wt, wi, wo = VirtualMachine.run(wp, input_data=[])
        wm = VirtualMachine.simulate(wp, input_data=wi)
        BrainfuckStark(wt, len(wm[1]), wp, wi, wo).prove(wp, *wm)
    fake = SeededUrandom(1234)  # the byte stream of bytes(random.Random(1234).getrandbits(8) for ...), generated in bulk
    os.urandom = fake
    salted_merkle.urandom = fake
    program = VirtualMachine.compile(source)
    running_time, input_symbols, output_symbols = VirtualMachine.run(program, input_data=list(inputs))
    processor_matrix, memory_matrix, instruction_matrix, input_matrix, output_matrix = VirtualMachine.simulate(
        program, input_data=input_symbols)
    bfs = BrainfuckStark(running_time, len(memory_matrix), program, input_symbols, output_symbols)
    launches0 = engine.launch_count()
    t0 = time.time()
    proof = bfs.prove(program, processor_matrix, memory_matrix, instruction_matrix, input_matrix, output_matrix)
    dt = time.time() - t0
    launches = engine.launch_count() - launches0
    dropin.uninstall()  # the verifier below is the reference's own, untouched
    t0 = time.time()
    ok = bool(bfs.verify(proof))
    vt = time.time() - t0
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", golden_name)))
    res = {"backend": backend, "program": source, "urandom_seed": 1234, "running_time": running_time,
           "fri_domain_length": bfs.fri.domain.length,
           "proof_len": len(proof), "proof_sha256": hashlib.sha256(proof).hexdigest(),
           "reference_verifier_accepts": ok, "prove_seconds": round(dt, 1), "reference_verify_seconds": round(vt, 1),
           "engine_calls_launching_kernels": int(launches)}
    if source == golden["program"]:  # programs the all-reference proof was recorded for
        res["byte_identical_to_reference_proof"] = hashlib.sha256(proof).hexdigest() == golden["proof_sha256"]
        res["reference_prove_seconds"] = golden["prove_seconds"]
    if os.environ.get("B2S_DUMP_PROOF"):
        with open(os.environ["B2S_DUMP_PROOF"], "wb") as f:
            f.write(proof)
    print(json.dumps(res, indent=1))
    if out:
        with open(out, "w") as f:
            json.dump(res, f, indent=1)
    return res


if __name__ == "__main__":
    main(*(sys.argv[1:6]))
