"""Host logic against the UNMODIFIED reference (authoring container only): the drop-in is installed
over a host-memory test backend, then (a) every front-end case must reproduce the golden values,
(b) the reference's own test files run unchanged, (c) uninstall() restores the originals."""
import importlib

import numpy as np
import pickle

import pytest

import frontend_cases as fc


@pytest.fixture(scope="module")
def env(reference_dropin):
    mods = [importlib.import_module(m) for m in
            ("algebra", "univariate", "extension_field", "ntt", "merkle", "ip", "fri", "salted_merkle")]
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
    ("test_merkle", ["test_merkle", "test_salted_merkle"]),
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
        # the next rows too: table methods, salted trees, and the recompiled prove() with its hook
        import brainfuck_stark
        import salted_merkle
        import table
        assert table.Table.lde.__qualname__ == "Table.lde" and table.Table.ldex.__qualname__ == "Table.ldex"
        assert table.Table.interpolate_columns.__qualname__ == "Table.interpolate_columns"
        assert table.Table.transition_quotients.__qualname__ == "Table.transition_quotients"
        assert salted_merkle.SaltedMerkle.__init__.__qualname__ == "SaltedMerkle.__init__"
        prove = brainfuck_stark.BrainfuckStark.prove
        assert prove.__qualname__ == "BrainfuckStark.prove" and not hasattr(prove, "__wrapped__")
        assert dropin._COMB_HOOK not in brainfuck_stark.__dict__ and dropin._COMB_HOOK not in prove.__code__.co_names
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


def test_quotient_codewords_match_reference(reference_dropin):
    """Table.boundary/transition/terminal_quotients and PermutationArgument.quotient on
    the device path against the reference's own per-point loops (values and pickles), on a toy table."""
    import pickle
    import random
    from conftest import REFERENCE_DIR
    from stark_brainfuck_b200 import dropin
    import algebra
    import extension_field
    import fri as fri_mod
    import multivariate
    import permutation_argument
    import table
    import univariate
    P = 18446744069414584321
    R = random.Random(4242)
    f = algebra.BaseField.main()
    xf = extension_field.ExtensionField.main()
    bf = xf.modulus.coefficients[0].field

    def X(*c):
        return extension_field.ExtensionFieldElement(
            univariate.Polynomial([algebra.BaseFieldElement(v, bf) for v in c]), xf)

    def rx():
        return X(R.randrange(P), R.randrange(P), R.randrange(P))

    N, W = 64, 3
    dom = fri_mod.Fri.Domain(f.generator(), f.primitive_nth_root(N), N)

    def mpoly(n_vars, n_mono):
        d = {}
        for _ in range(n_mono):
            k = [0] * n_vars
            for _ in range(R.randrange(0, 4)):
                k[R.randrange(n_vars)] += R.randrange(1, 4)
            d[tuple(k)] = rx() if R.random() < 0.8 else X(R.randrange(5))
        return multivariate.MPolynomial(d)

    class Toy(table.Table):
        def __init__(self, length):
            super().__init__(xf, 2, W, length, 1, f.primitive_nth_root(N), N)
            self.b = [mpoly(W, 3), mpoly(W, 1), multivariate.MPolynomial(dict())]
            self.t = [mpoly(2 * W, 6), mpoly(2 * W, 2)]
            self.e = [mpoly(W, 4)]

        def boundary_constraints_ext(self, challenges):
            return self.b

        def transition_constraints_ext(self, challenges):
            return self.t

        def terminal_constraints_ext(self, challenges, terminals):
            return self.e

    tables = [Toy(5), Toy(0)]
    for t in tables:
        t.codewords = [[rx() for _ in range(N)] for _ in range(W)]
        t.codewords[1][3] = xf.zero()
        t.codewords[2][7] = X(11)
    pa = permutation_argument.PermutationArgument(tables, (0, 1), (1, 2))

    def run():
        out = []
        for t in tables:
            out.append(t.boundary_quotients(dom, t.codewords, None))
            out.append(t.transition_quotients(dom, t.codewords, None))
            out.append(t.terminal_quotients(dom, t.codewords, None, None))
        out.append([pa.quotient(dom)])
        return out

    eng = reference_dropin.engine
    launches0 = eng.launch_count()
    got = run()
    assert eng.launch_count() > launches0
    dropin.uninstall()
    try:
        want = run()
    finally:
        dropin.install(REFERENCE_DIR, engine=eng)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert len(g) == len(w)
        for gc, wc in zip(g, w):
            assert [[c.value for c in x.polynomial.coefficients] for x in gc] == \
                   [[c.value for c in x.polynomial.coefficients] for x in wc]
            assert pickle.dumps(gc) == pickle.dumps(wc)


def test_salted_merkle_matches_reference(reference_dropin):
    """SaltedMerkle over tuple leaves, seeded salts: same roots, openings and pickles"""
    import pickle
    import random
    from conftest import REFERENCE_DIR
    from stark_brainfuck_b200 import dropin
    import algebra
    import extension_field
    import salted_merkle
    import univariate
    xf = extension_field.ExtensionField.main()
    bf = xf.modulus.coefficients[0].field
    f = algebra.BaseField.main()
    R = random.Random(7)
    P = 18446744069414584321
    leaves = {n: [tuple([algebra.BaseFieldElement(R.randrange(P), f) for _ in range(3)] +
                        [extension_field.ExtensionFieldElement(univariate.Polynomial(
                            [algebra.BaseFieldElement(R.randrange(P), bf) for _ in range(R.randrange(0, 4))]), xf)])
                  for _ in range(n)] for n in (1, 5, 64)}
    old = salted_merkle.urandom

    def run():
        out = []
        for n, data in leaves.items():
            S = random.Random(1000 + n)
            salted_merkle.urandom = lambda k: bytes(S.getrandbits(8) for _ in range(k))
            t = salted_merkle.SaltedMerkle(data)
            opened = [t.open(i) for i in sorted({0, n - 1, n // 2})]
            for i, (salt, path) in zip(sorted({0, n - 1, n // 2}), opened):
                assert salted_merkle.SaltedMerkle.verify(t.root(), i, salt, path, data[i])
            out.append(pickle.dumps((t.root(), opened, t.depth, t.num_leafs, [t.leafs[i] for i in range(n)])))
        return out

    eng = reference_dropin.engine
    try:
        got = run()
        dropin.uninstall()
        try:
            want = run()
        finally:
            dropin.install(REFERENCE_DIR, engine=eng)
    finally:
        salted_merkle.urandom = old
    assert got == want
    with pytest.raises(AssertionError):
        salted_merkle.SaltedMerkle([])


def test_salted_row_trees_golden(env):
    fc.case_salted(env, env.glue)


def test_nonlinear_combination_matches_reference_block(env):
    """SURVEY 8(f) row 3 against the outputs of the reference's own statements"""
    fc.case_combination(env, env.glue)


def test_prove_is_recompiled_with_guarded_combination_block(env, monkeypatch):
    """install() recompiles BrainfuckStark.prove from the reference's own source with the combination
    block guarded by the hook; the block itself stays as the fallback (DEBUG, or a hook returning None)"""
    import inspect
    import brainfuck_stark
    from stark_brainfuck_b200 import dropin
    assert dropin.installed()
    prove = brainfuck_stark.BrainfuckStark.prove.__wrapped__  # under the keep_planes() scope
    assert inspect.getsourcefile(prove).endswith("brainfuck_stark.py")
    assert dropin._COMB_HOOK in prove.__code__.co_names and dropin._COMB_HOOK in brainfuck_stark.__dict__
    assert "terms" in prove.__code__.co_varnames  # the reference's block is still compiled in
    hook = brainfuck_stark.__dict__[dropin._COMB_HOOK]
    monkeypatch.setenv("DEBUG", "1")
    assert hook(None) is None  # DEBUG: reference block runs, nothing else is touched
    monkeypatch.delenv("DEBUG")

    class Moved:
        def prove(self):
            pass
    with pytest.raises(LookupError):
        dropin.guarded_prove_source(Moved.prove)


def test_sample_and_lift_build_the_reference_object_graph(env):
    """the constructor-free ExtensionField.sample / lift and BaseField.sample of the drop-in against the reference's
    own methods: same pickles (values, trimming, field objects) for bytes of every length the prover uses"""
    import random
    import algebra
    import extension_field
    from stark_brainfuck_b200 import dropin
    R = random.Random(12)
    bsample = algebra.BaseField.sample.__wrapped__ if hasattr(algebra.BaseField.sample, "__wrapped__") else None
    saved = dropin._state["saved"]["attrs"]
    orig = {(obj.__name__, name): fn for obj, name, fn in saved if name in ("sample", "lift")}
    ob, ox, ol = orig[("BaseField", "sample")], orig[("ExtensionField", "sample")], orig[("ExtensionField", "lift")]
    assert bsample is None and ob is not algebra.BaseField.sample and ox is not extension_field.ExtensionField.sample
    f, xf = env.field, env.xfield
    cases = [b"", b"\x00", b"\x01\x02", bytes(24), bytes(27), b"\xff" * 24, b"\xff" * 27, bytes(16) + b"\x07" * 8,
             bytes(8) + b"\x01" + bytes(15), b"\x05" + bytes(23)]
    cases += [bytes(R.getrandbits(8) for _ in range(n)) for n in (1, 2, 3, 7, 24, 24, 27, 27, 32, 32, 64, 100)]
    P = f.p
    cases += [(P).to_bytes(8, "big") * 3, (P - 1).to_bytes(8, "big") * 3, (P + 5).to_bytes(9, "big") * 3]
    for b in cases:
        assert pickle.dumps(f.sample(b)) == pickle.dumps(ob(f, b)), b
        got, want = xf.sample(b), ox(xf, b)
        assert pickle.dumps(got) == pickle.dumps(want), b
        assert got.field is xf and all(c.field is xf.modulus.coefficients[0].field for c in got.polynomial.coefficients)
        assert pickle.dumps([got, got]) == pickle.dumps([want, want])
    for b in (list(b"\x01\x02\x03"), bytearray(b"\x09" * 24)):  # other byte containers: the reference's own code path
        assert pickle.dumps(xf.sample(b)) == pickle.dumps(ox(xf, b)) and f.sample(b).value == ob(f, b).value
    for v in (0, 1, P - 1):
        e = env.BaseFieldElement(v, f)
        assert pickle.dumps(xf.lift(e)) == pickle.dumps(ol(xf, e)) and xf.lift(xf.lift(e)) is not None
    x = xf.sample(b"\x01" * 27)
    assert xf.lift(x) is x
    # Polynomial.degree: the same integer as the reference's scan, for base-field and extension-field coefficients
    import univariate
    od = [fn for obj, name, fn in saved if name == "degree"][0]
    assert od is not univariate.Polynomial.degree
    for _ in range(300):
        n = R.randrange(0, 9)
        vals = [R.choice([0, 0, 1, P - 1, R.randrange(P)]) for _ in range(n)]
        pb = univariate.Polynomial([env.BaseFieldElement(v, f) for v in vals])
        assert pb.degree() == od(pb), vals
        px = univariate.Polynomial([fc.X(env, v, 0, R.choice([0, v])) for v in vals])
        assert px.degree() == od(px), vals
    assert univariate.Polynomial([]).degree() == -1


def test_extension_field_arithmetic_builds_the_reference_object_graph(env):
    """the direct ExtensionField.multiply / add / subtract / negate of the drop-in against the reference's own methods
    (Polynomial product + long division): pickles of (operands, result) -- values, trimming, the field object of every
    coefficient, coefficient objects shared with an operand or with each other -- over sparse, empty and full operands
    whose coefficients carry different BaseField objects"""
    import random
    import algebra
    import extension_field
    import univariate
    from stark_brainfuck_b200 import dropin
    saved = dropin._state["saved"]["attrs"]
    orig = {name: fn for obj, name, fn in saved if obj is extension_field.ExtensionField}
    XF = extension_field.ExtensionField
    assert all(orig[n] is not XF.__dict__[n] for n in ("multiply", "add", "subtract", "negate"))
    R = random.Random(31)
    xf = env.xfield
    P = env.field.p
    fields = [env.field, xf.modulus.coefficients[0].field, algebra.BaseField.main(), algebra.BaseField.main()]

    def elem(pattern):
        co = []
        for k in pattern:
            v = 0 if k == "0" else (R.choice([1, P - 1, 2]) if k == "s" else R.randrange(1, P))
            co.append(algebra.BaseFieldElement(v, R.choice(fields)))
        return extension_field.ExtensionFieldElement(univariate.Polynomial(co), xf)
    patterns = ["", "r", "s", "0r", "rr", "00r", "0rr", "r0r", "rrr", "sss", "s0s", "00s"]
    n = 0
    for pa in patterns:
        for pb in patterns:
            for _ in range(3):
                a, b = elem(pa), elem(pb)
                for name in ("multiply", "add", "subtract"):
                    got, want = XF.__dict__[name](xf, a, b), orig[name](xf, a, b)
                    assert pickle.dumps((a, b, got)) == pickle.dumps((a, b, want)), (name, pa, pb)
                    assert got.field is xf
                    n += 1
                assert pickle.dumps((a, xf.negate(a))) == pickle.dumps((a, orig["negate"](xf, a)))
                assert pickle.dumps((a, a * a, a - a, a + a)) == pickle.dumps(
                    (a, orig["multiply"](xf, a, a), orig["subtract"](xf, a, a), orig["add"](xf, a, a)))
    assert n > 1200
    # products that cancel to a lower degree, and operands the fast path must hand to the reference: untrimmed
    # coefficient lists, four coefficients, a foreign modulus
    one, x = xf.one(), elem("0s")
    inv = x.inverse()
    assert pickle.dumps((x, inv, x * inv)) == pickle.dumps((x, inv, orig["multiply"](xf, x, inv)))
    long = extension_field.ExtensionFieldElement.__new__(extension_field.ExtensionFieldElement)
    long.__dict__ = {"polynomial": univariate.Polynomial([env.field(1), env.field(2), env.field(3), env.field(4)]), "field": xf}
    untrimmed = extension_field.ExtensionFieldElement.__new__(extension_field.ExtensionFieldElement)
    untrimmed.__dict__ = {"polynomial": univariate.Polynomial([env.field(5), env.field(0)]), "field": xf}
    for odd in (long, untrimmed):
        for name in ("multiply", "add", "subtract"):
            assert pickle.dumps(XF.__dict__[name](xf, odd, one)) == pickle.dumps(orig[name](xf, odd, one))
            assert pickle.dumps(XF.__dict__[name](xf, one, odd)) == pickle.dumps(orig[name](xf, one, odd))
    other = extension_field.ExtensionField(univariate.Polynomial([env.field(2), env.field(P - 1), env.field(0), env.field(1)]))
    y, z = other.sample(bytes(range(27))), other.sample(bytes(range(3, 30)))
    assert pickle.dumps(other.multiply(y, z)) == pickle.dumps(orig["multiply"](other, y, z))


def test_lazy_codewords_inside_prove(env):
    """the device views prove() works on under the drop-in: rows hook compiled into prove(), every Table.extend wrapped
    so that its lifting statement runs over the views, and the views themselves against the reference's codewords"""
    import brainfuck_stark
    import io_table
    import instruction_table
    import memory_table
    import processor_table
    from stark_brainfuck_b200 import dropin
    from stark_brainfuck_b200.glue import DeviceCodeword
    prove = brainfuck_stark.BrainfuckStark.prove.__wrapped__
    assert dropin._ROWS_HOOK in prove.__code__.co_names
    assert all(dropin.lazy_rows_source(stmt) == stmt.split(" = ")[0] + " = %s(%s)" % (dropin._ROWS_HOOK, stmt[stmt.index("*") + 1:-2])
               for stmt in dropin._ROWS_STATEMENTS)
    assert brainfuck_stark.__dict__[dropin._ROWS_HOOK]([[1, 2], [3, 4]]) == [(1, 3), (2, 4)]  # host lists: zip as written
    for cls, meth in ((processor_table.ProcessorTable, "extend"), (instruction_table.InstructionTable, "extend"),
                      (memory_table.MemoryTable, "extend"), (io_table.IOTable, "extend_iotable")):
        assert hasattr(cls.__dict__[meth], "__wrapped__"), (cls, meth)
    fc.case_lazy_codewords(env, env.glue)
    # Table.extend over device views: the reference's statement sees an empty list, the views are lifted by the glue;
    # over host lists (outside prove) the method is the reference's own
    glue = dropin.current_glue()  # the instance the wrappers consult (another test may have re-installed)
    t = io_table.InputTable(env.field, 2, env.field(7), 8)
    t.matrix = [[env.field(3)], [env.field(4)]]
    t.pad()
    plain = [[env.field(5), env.field(0)]]
    t.codewords = plain
    t.extend([env.xfield.one()] * 11, [])
    assert t.codewords is not plain and t.codewords[0][0].polynomial.coefficients[0] is plain[0][0]
    assert t.codewords[0][1].polynomial.coefficients == []
    t2 = io_table.InputTable(env.field, 2, env.field(7), 8)
    t2.matrix = [[env.field(3)], [env.field(4)]]
    t2.pad()
    with glue.keep_planes(lazy=True):
        view = DeviceCodeword(glue, glue.engine.upload(np.array([[5, 0]], dtype=np.uint64)), env.field, "b")
        t2.codewords = [view]
        t2.extend([env.xfield.one()] * 11, [])
        assert type(t2.codewords[0]) is DeviceCodeword and t2.codewords[0].kind == "l"
        assert pickle.dumps(list(t2.codewords[0])) == pickle.dumps(t.codewords[0])
        assert t2.codewords[0][0].polynomial.coefficients[0] is view[0]
    assert fc.triples([row[-1] for row in t2.matrix]) == fc.triples([row[-1] for row in t.matrix])


def test_table_lde_matches_reference(env):
    """SURVEY 8(f) row 2 through the reference's own Table class (lde/ldex/interpolate_columns rebound)"""
    import table
    fc.case_lde(env, env.glue, make_table=table.Table)
    # and through the rebound methods themselves, randomizers from the module's os.urandom
    from util import golden
    g = golden("lde.json")
    c = g["cases"][4]
    dom = env.Fri.Domain(env.field(g["offset"]), env.field(g["omega"]), g["N"])
    t = table.Table(env.field, c["base_width"], c["full_width"], c["length"], c["num_randomizers"],
                    env.field(g["omega"]), g["N"])
    t.matrix = [[env.BaseFieldElement(v, env.field) for v in row] for row in c["base"]]
    saved = table.os.urandom
    table.os.urandom = fc.seeded_urandom(c["urandom_seed"])
    try:
        cws = t.lde(dom)
        assert cws is t.codewords and [fc.vals(cw) for cw in cws] == c["base_codewords"]
        t.field = env.xfield
        t.matrix = [[env.xfield.lift(v) for v in t.matrix[r]] + [fc.X(env, *v) for v in c["ext"][r]]
                    for r in range(c["height"])]
        ext = t.ldex(dom, env.xfield)
        assert [fc.triples(cw) for cw in ext] == c["ext_codewords"] and len(t.codewords) == c["full_width"]
    finally:
        table.os.urandom = saved
