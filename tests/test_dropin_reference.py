"""Host logic against the UNMODIFIED reference (authoring container only): the drop-in is installed
over a host-memory test backend, then (a) every front-end case must reproduce the golden values,
(b) the reference's own test files run unchanged, (c) uninstall() restores the originals."""
import importlib
import sys

import pytest

import frontend_cases as fc


@pytest.fixture(scope="module")
def env(reference_dropin):
    mods = [importlib.import_module(m) for m in
            ("algebra", "univariate", "extension_field", "ntt", "merkle", "ip", "fri")]
    e = fc.make_env(*mods)
    e.glue = reference_dropin
    return e


def test_patched_everywhere(env):
    import fri
    import ntt
    assert ntt.ntt.__self__ is env.glue and fri.ntt.__self__ is env.glue  # star-imported copy rebound too
    assert ntt.fast_multiply.__module__ == "ntt"  # untouched reference code riding on the patched ntt


def test_ntt_golden(env):
    fc.case_ntt_golden(env)


def test_ntt_quirks(env):
    fc.case_ntt_quirks(env)


def test_coset_and_poly(env):
    fc.case_coset_and_poly(env)


def test_merkle(env):
    fc.case_merkle(env)


def test_fri_transcripts(env):
    fc.case_fri_transcripts(env)


def test_test_fri_config(env):
    fc.case_test_fri_config(env)


def test_gv3(env):
    fc.case_gv3(env)


def test_fri_errors(env):
    fc.case_fri_errors(env)


@pytest.mark.parametrize("modname,funcs", [
    ("test_ntt", ["test_ntt", "test_intt", "test_multiply", "test_divide", "test_interpolate",
                  "test_coset_evaluate", "test_batch_inverse"]),
    ("test_merkle", ["test_merkle"]),
    ("test_fri", ["test_fri"]),
])
def test_reference_own_tests_run_unmodified(env, modname, funcs, capsys):
    """code/test_ntt.py, code/test_merkle.py, code/test_fri.py imported from the read-only checkout"""
    launches0 = env.glue.engine.launch_count()
    mod = importlib.import_module(modname)
    for fn in funcs:
        getattr(mod, fn)()
    assert env.glue.engine.launch_count() > launches0  # they really went through the engine


def test_uninstall_restores(reference_dropin):
    from stark_brainfuck_b200 import dropin
    import fri
    import ntt
    import univariate
    glue = reference_dropin
    eng = glue.engine
    dropin.uninstall()
    try:
        assert ntt.ntt.__module__ == "ntt" and fri.ntt is ntt.ntt
        assert univariate.Polynomial.scale.__qualname__ == "Polynomial.scale"
        assert fri.Fri.prove.__qualname__ == "Fri.prove"
    finally:
        from conftest import REFERENCE_DIR
        dropin.install(REFERENCE_DIR, engine=eng)


def test_identity_graph_of_transform_outputs(reference_dropin):
    """pickle memoises by identity: constant / zero / generic extension-field polynomials must give lists
    whose pickles equal the reference's own (shared coefficient objects of constant columns)."""
    import pickle
    import random
    from conftest import REFERENCE_DIR
    from stark_brainfuck_b200 import dropin
    import algebra
    import extension_field
    import fri as fri_mod
    import ntt as ntt_mod
    import univariate
    xf = extension_field.ExtensionField.main()
    bf = xf.modulus.coefficients[0].field
    f = algebra.BaseField.main()
    R = random.Random(99)

    def X(*c):
        return extension_field.ExtensionFieldElement(
            univariate.Polynomial([algebra.BaseFieldElement(v, bf) for v in c]), xf)

    n = 16
    dom = fri_mod.Fri.Domain(f.generator(), f.primitive_nth_root(n), n)
    P = 18446744069414584321
    polys = {
        "constant": univariate.Polynomial([X(5, 6, 7)]),
        "constant_with_explicit_zeros": univariate.Polynomial([X(R.randrange(P), 1)] + [xf.zero()] * 3),
        "zero": univariate.Polynomial([xf.zero()] * 2),
        "generic": univariate.Polynomial([X(R.randrange(P), R.randrange(P), R.randrange(P)) for _ in range(4)]),
        "leading_zero": univariate.Polynomial([xf.zero(), X(3)]),
        "c0_plus_c8_x8": univariate.Polynomial([X(4, 5)] + [xf.zero()] * 7 + [X(R.randrange(P), 2, 3)]),
        "even_powers_only": univariate.Polynomial([X(1), xf.zero(), X(2, 2), xf.zero(), xf.zero(), xf.zero(), X(7)]),
        "x4_only": univariate.Polynomial([xf.zero()] * 4 + [X(R.randrange(P))]),
    }
    lone = [X(9, 8, 7)] + [xf.zero()] * 7

    def run():
        out = {k: pickle.dumps(dom.xevaluate(p, xf)) for k, p in polys.items()}
        out["ntt_lone"] = pickle.dumps(ntt_mod.ntt(xf.lift(f.primitive_nth_root(8)), lone))
        return out

    eng = reference_dropin.engine
    got = run()
    dropin.uninstall()
    try:
        want = run()
    finally:
        dropin.install(REFERENCE_DIR, engine=eng)
    for k in want:
        assert got[k] == want[k], k
