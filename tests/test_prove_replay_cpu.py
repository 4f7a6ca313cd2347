"""The recorded device command stream of BrainfuckStark.prove() (tests/golden/trace_*.bin, written by
tests/golden/make_trace.py from the UNMODIFIED reference under the drop-in) replayed WITHOUT the reference and
without the glue, here over the host-memory test backend: checks the recorder / replayer pair that
tests/test_gpu_prove_replay.py uses on the B200."""
import os

import pytest

import trace_backend as tb
from util import GOLDEN


@pytest.mark.parametrize("name", ["pppp", "io"])
def test_replay_reproduces_every_recorded_output(name):
    from fake_backend import fake_engine
    path = os.path.join(GOLDEN, "trace_%s.bin" % name)
    res = tb.replay(path, fake_engine())
    assert res["calls"] >= 70 and res["kernel_outputs_checked"] >= 70 and res["host_reads_checked"] >= 20
    assert res["meta"]["reference_verifier_accepts"] is True


def test_replay_detects_a_wrong_kernel():
    """a backend whose fold is off by one must fail the comparison"""
    from fake_backend import FakeLib
    from stark_brainfuck_b200 import Engine

    class Broken(FakeLib):
        def b2s_fri_fold(self, *a):
            rc = super().b2s_fri_fold(*a)
            import ctypes as C
            C.c_uint64.from_address(a[6].value if hasattr(a[6], "value") else a[6]).value ^= 1
            return rc
    with pytest.raises(AssertionError):
        tb.replay(os.path.join(GOLDEN, "trace_pppp.bin"), Engine(lib=Broken(), device="cpu"))


def test_seeded_urandom_is_the_stream_of_the_golden_proofs():
    import random
    R = random.Random(1234)
    u = tb.SeededUrandom(1234, chunk=64)
    for n in (1, 24, 27, 5, 300, 24):
        assert u(n) == bytes(R.getrandbits(8) for _ in range(n))
