"""BASELINE config 5 at the scale the reference can finish: the UNMODIFIED
brainfuck_stark.BrainfuckStark.prove("++++") with the drop-in installed -- hot path and quotient
codewords through the C-ABI surface (host-memory test backend here; the GPU box has no reference
checkout) -- must produce a proof that the reference's own verifier accepts and that is byte-identical
to the all-reference proof recorded in tests/golden/bfs.json (seeded urandom)."""
import os
import subprocess
import sys
import json

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_DIR = os.environ.get("B2S_REFERENCE_DIR", "/root/reference/code")


@pytest.mark.skipif(not os.path.isdir(REFERENCE_DIR), reason="reference checkout not available")
def test_prove_under_dropin_is_byte_identical_and_accepted(tmp_path):
    out = str(tmp_path / "res.json")
    # own process: prove() patches os.urandom and imports the reference's modules by bare name
    subprocess.check_call([sys.executable, os.path.join(HERE, "e2e_prove_dropin.py"), "fake", out],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=900)
    res = json.load(open(out))
    assert res["reference_verifier_accepts"] is True
    assert res["byte_identical_to_reference_proof"] is True
    assert res["fri_domain_length"] == 1024 and res["engine_calls_launching_kernels"] > 0


@pytest.mark.skipif(not os.path.isdir(REFERENCE_DIR), reason="reference checkout not available")
def test_hello_world_prove_is_accepted_by_the_reference_verifier(tmp_path):
    """BASELINE north star: brainfuck_stark.prove() on the Hello-World trace (907 cycles, FRI domain 2^17) under the
    drop-in -- about a minute on the host-memory test backend; the all-Python reference needs many hours, so
    acceptance is the reference's own verifier plus a proof hash that stays fixed under the seeded urandom."""
    out = str(tmp_path / "res.json")
    subprocess.check_call([sys.executable, os.path.join(HERE, "e2e_prove_dropin.py"), "fake", out, "hello"],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=1800)
    res = json.load(open(out))
    assert res["reference_verifier_accepts"] is True
    assert res["running_time"] == 907 and res["fri_domain_length"] == 1 << 17
    assert res["proof_sha256"] == "540a9a28053b3195231dc7736163b760d8015a7159b973b85307e45ace4a6f3e"


@pytest.mark.skipif(not os.path.isdir(REFERENCE_DIR), reason="reference checkout not available")
@pytest.mark.parametrize("source,inputs,name,domain", [(",.", "x", "bfs_cat.json", 512),
                                                       ("++..,", "q", "bfs_two.json", 1024),
                                                       ("++[>,.<-]", "ab", "bfs_io.json", 2048),
                                                       ("+++++[>,.<-]", "hello", "bfs_echo.json", 4096),
                                                       ("+++[>+++[>+<-]<-]>>.", "", "bfs_nested.json", 8192),
                                                       ("++++++++[>++++++++<-]>+.", "", "bfs_A.json", 16384),
                                                       ("++++++++++[>+++++++>++++++++++<<-]>++.>+.", "", "bfs_He.json",
                                                        32768)])
def test_prove_with_loop_and_io_is_byte_identical(tmp_path, source, inputs, name, domain):
    """programs with a loop, input and output symbols (FRI domains 2048 ... 32768):
    byte-identical to the all-reference proofs in tests/golden/bfs_io.json / bfs_echo.json / bfs_nested.json / bfs_A.json / bfs_He.json
    (reference: 350 s / 755 s / 1569 s / 3045 s / 6109 s)"""
    out = str(tmp_path / "res.json")
    subprocess.check_call([sys.executable, os.path.join(HERE, "e2e_prove_dropin.py"), "fake", out, source, inputs, name],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=900)
    res = json.load(open(out))
    assert res["reference_verifier_accepts"] is True and res["byte_identical_to_reference_proof"] is True
    assert res["fri_domain_length"] == domain
