import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


# ---- switching between the real reference modules and the standalone mirror ----------------
# Both use the bare module names (pickle records them), so a test module picks ONE of the two
# fixtures below; each restores sys.modules on teardown.
REFERENCE_DIR = os.environ.get("B2S_REFERENCE_DIR", "/root/reference/code")
REF_MODULES = ("algebra", "univariate", "extension_field", "ntt", "merkle", "salted_merkle", "ip", "fri",
               "multivariate", "table", "processor_table", "instruction_table", "memory_table", "io_table",
               "permutation_argument", "evaluation_argument", "brainfuck_stark", "vm",
               "test_ntt", "test_fri", "test_merkle", "test_brainfuck_stark", "test_extension_field")


def _purge():
    saved = {}
    for name in REF_MODULES:
        if name in sys.modules:
            saved[name] = sys.modules.pop(name)
    return saved


@pytest.fixture(scope="module")
def reference_dropin():
    """the UNMODIFIED reference (only in the authoring container) with the drop-in installed over a
    host-memory test backend"""
    if not os.path.isdir(REFERENCE_DIR):
        pytest.skip("reference checkout not available")
    from fake_backend import fake_engine
    from stark_brainfuck_b200 import dropin
    saved = _purge()
    sys.path.insert(0, REFERENCE_DIR)
    old_dont = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    glue = dropin.install(REFERENCE_DIR, engine=fake_engine())
    yield glue
    dropin.uninstall()
    sys.dont_write_bytecode = old_dont
    sys.path.remove(REFERENCE_DIR)
    _purge()
    sys.modules.update(saved)


def _mirror_env(engine):
    from stark_brainfuck_b200 import mirror
    from stark_brainfuck_b200.glue import Glue
    saved = _purge()
    mirror.register()
    old = mirror._glue
    mirror.set_glue(Glue(mirror.binding, engine))
    return mirror, saved, old


@pytest.fixture(scope="module")
def mirror_cpu():
    """the standalone mirror over the host-memory test backend (host logic without a GPU)"""
    from fake_backend import fake_engine
    mirror, saved, old = _mirror_env(fake_engine())
    yield mirror
    mirror.set_glue(old)
    mirror.unregister()
    sys.modules.update(saved)


@pytest.fixture(scope="module")
def mirror_gpu():
    """the standalone mirror over libb2s.so on cuda:0"""
    from stark_brainfuck_b200 import Engine
    mirror, saved, old = _mirror_env(Engine(0))
    yield mirror
    mirror.set_glue(old)
    mirror.unregister()
    sys.modules.update(saved)
