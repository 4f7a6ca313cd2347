"""BASELINE config 5 end to end on the GPU: the UNMODIFIED BrainfuckStark.prove() of a reference checkout with the
drop-in installed over libb2s.so on cuda:0.  Needs a box that has both a GPU and the reference: the authoring
container stages a copy under the git-ignored baseline/_ref/code for its gpurun calls (the reference is 25 pure
Python files; nothing of it is in the repository's history).  Skipped where that copy is absent --
tests/test_gpu_prove_replay.py then covers the same call sequence from its recorded command stream."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STAGED = os.environ.get("B2S_REFERENCE_DIR", os.path.join(ROOT, "baseline", "_ref", "code"))


def _prove(tmp_path, *args):
    if not os.path.isfile(os.path.join(STAGED, "brainfuck_stark.py")):
        pytest.skip("no reference checkout staged next to the GPU")
    out = str(tmp_path / "res.json")
    env = dict(os.environ, B2S_REFERENCE_DIR=STAGED)
    subprocess.check_call([sys.executable, os.path.join(HERE, "e2e_prove_dropin.py"), "gpu", out] + list(args), env=env,
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=1800)
    res = json.load(open(out))
    keep = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(keep):
        tag = (args[2][:-5] if len(args) > 2 else args[0]) if args else "pppp"
        with open(os.path.join(keep, "e2e_prove_gpu_%s.json" % tag), "w") as f:
            json.dump(res, f, indent=1)
    return res


def test_prove_pppp_on_the_gpu_is_byte_identical_to_the_reference_proof(tmp_path):
    res = _prove(tmp_path)
    assert res["backend"] == "gpu" and res["reference_verifier_accepts"] is True
    assert res["byte_identical_to_reference_proof"] is True and res["engine_calls_launching_kernels"] > 0


def test_hello_world_prove_on_the_gpu_is_accepted_by_the_reference_verifier(tmp_path):
    res = _prove(tmp_path, "hello")
    assert res["reference_verifier_accepts"] is True and res["fri_domain_length"] == 1 << 17
    assert res["proof_sha256"] == "540a9a28053b3195231dc7736163b760d8015a7159b973b85307e45ace4a6f3e"


@pytest.mark.parametrize("source,inputs,name,domain", [(",.", "x", "bfs_cat.json", 512),
                                                       ("++..,", "q", "bfs_two.json", 1024),
                                                       ("++[>,.<-]", "ab", "bfs_io.json", 2048),
                                                       ("+++++[>,.<-]", "hello", "bfs_echo.json", 4096),
                                                       ("+++[>+++[>+<-]<-]>>.", "", "bfs_nested.json", 8192),
                                                       ("++++++++[>++++++++<-]>+.", "", "bfs_A.json", 16384),
                                                       ("++++++++++[>+++++++>++++++++++<<-]>++.>+.", "", "bfs_He.json",
                                                        32768)])
def test_programs_with_loops_and_io_on_the_gpu_are_byte_identical(tmp_path, source, inputs, name, domain):
    """the five other programs whose all-reference proofs are recorded (tests/golden/bfs_*.json; the reference needs
    350 ... 6 109 s each): input and output tables of height 0, 1 (unit distance = the whole domain), 2 and more"""
    res = _prove(tmp_path, source, inputs, name)
    assert res["reference_verifier_accepts"] is True and res["byte_identical_to_reference_proof"] is True
    assert res["fri_domain_length"] == domain

