"""Host logic of the standalone mirror (what runs on the GPU box) over the host-memory test
backend: same cases, same golden values as the drop-in run against the real reference."""
import pytest

import frontend_cases as fc


@pytest.fixture(scope="module")
def env(mirror_cpu):
    m = mirror_cpu
    return fc.make_env(m.algebra, m.univariate, m.extension_field, m.ntt, m.merkle, m.ip, m.fri, m.salted_merkle)


def test_ntt_golden(env):
    fc.case_ntt_golden(env, max_log=10)


def test_ntt_quirks(env):
    fc.case_ntt_quirks(env)


def test_coset_and_poly(env, mirror_cpu):
    env.glue = mirror_cpu.glue()
    fc.case_coset_and_poly(env)


def test_merkle(env):
    fc.case_merkle(env)


def test_fri_transcripts(env):
    fc.case_fri_transcripts(env, logs=(4, 6, 8))


def test_test_fri_config(env):
    fc.case_test_fri_config(env)


def test_gv3(env):
    fc.case_gv3(env)


def test_fri_errors(env):
    fc.case_fri_errors(env)


def test_mirror_field_arithmetic(env):
    """the mirror's own per-element arithmetic against the reference's known answers"""
    from util import golden
    S = golden("small.json")
    for c in S["xfe_arith"]:
        a, b = fc.X(env, *c["a"]), fc.X(env, *c["b"])
        assert fc.triples([a * b, a + b, a - b, a.inverse(), a / b]) == [c["mul"], c["add"], c["sub"], c["inv"], c["div"]]
    f = env.field
    for a, b, s, d, m, i, n in S["bfe_arith"]:
        A, B = f(a), f(b)
        assert [(A + B).value, (A - B).value, (A * B).value, A.inverse().value, (-A).value] == [s, d, m, i, n]
    for c in S["sample"]:
        seed = bytes.fromhex(c["seed"])
        assert fc.triples([env.xfield.sample(seed)]) == [c["xfe"]] and f.sample(seed).value == c["bfe"]
    for k, v in S["roots"].items():
        assert f.primitive_nth_root(1 << int(k)).value == v


def test_nonlinear_combination(env, mirror_cpu):
    fc.case_combination(env, mirror_cpu.glue())


def test_table_lde(env, mirror_cpu):
    fc.case_lde(env, mirror_cpu.glue())


def test_lazy_codewords(env, mirror_cpu):
    fc.case_lazy_codewords(env, mirror_cpu.glue())


def test_quotients_through_the_glue(env, mirror_cpu):
    fc.case_quotients_glue(env, mirror_cpu.glue())


def test_zerofier_decision_on_the_host(env, mirror_cpu):
    fc.case_zerofier_decision(env, mirror_cpu.glue())


def test_salted_row_trees(env, mirror_cpu):
    fc.case_salted(env, mirror_cpu.glue())
