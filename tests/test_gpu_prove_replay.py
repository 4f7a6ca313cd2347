"""BASELINE config 5 on the B200: the complete device command stream of BrainfuckStark.prove() -- recorded in the
authoring container from the UNMODIFIED reference under the drop-in (tests/golden/make_trace.py; the recorded
proof is byte-identical to the all-reference proof for "++++" and is accepted by the reference verifier for
Hello World, hash 540a9a28...) -- replayed through libb2s.so on cuda:0.  Every region a kernel writes and every
byte that goes back to the host (the three Merkle roots, all openings, the codewords the host materialises) must
equal the recording, so the proof the host assembles from them is the recorded proof."""
import json
import os
import time

import pytest

pytestmark = pytest.mark.gpu

import trace_backend as tb  # noqa: E402
from util import GOLDEN  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# "big": an 8 780-cycle program on a 2^20 FRI domain, BASELINE config 5's size (proof a9ce6dd2..., the same bytes the GPU
# produced with the reference staged next to it and the CPU oracle produced when the trace was recorded)
@pytest.mark.parametrize("name,domain,min_calls", [("pppp", 1024, 70), ("io", 2048, 70), ("hello", 1 << 17, 80),
                                                   ("big", 1 << 20, 80)])
def test_prove_command_stream_replays_bit_exact(name, domain, min_calls):
    from stark_brainfuck_b200 import Engine
    path = os.path.join(GOLDEN, "trace_%s.bin" % name)
    if not os.path.exists(path):
        pytest.skip("trace not recorded")
    eng = Engine(0)
    launches0 = eng.launch_count()
    t0 = time.perf_counter()
    res = tb.replay(path, eng)
    dt = time.perf_counter() - t0
    launches = eng.launch_count() - launches0
    assert res["meta"]["fri_domain_length"] == domain and res["meta"]["reference_verifier_accepts"] is True
    assert res["calls"] >= min_calls and res["kernel_outputs_checked"] >= min_calls and launches >= min_calls
    # the same stream without the per-kernel comparisons: device time of a whole proof
    t0 = time.perf_counter()
    tb.replay(path, eng, check_kernels=False, check_reads=False)
    dt_fast = time.perf_counter() - t0
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "prove_replay_%s.json" % name), "w") as f:
            json.dump(dict(res, kernel_launches=int(launches), replay_seconds_with_checks=round(dt, 3),
                           replay_seconds_no_checks=round(dt_fast, 4)), f, indent=1)
