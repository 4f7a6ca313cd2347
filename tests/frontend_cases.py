"""Front-end parity cases shared by the three environments:
   reference + drop-in + test backend (CPU), mirror + test backend (CPU), mirror + libb2s.so (GPU).
Each case takes `env`, a namespace with the reference-named classes/functions."""
import hashlib
import pickle
import random
import types

import numpy as np
import pytest

from util import P, bfe_digest, golden, have_golden, xfe_digest

S = golden("small.json")
M = golden("merkle.json")
FS = golden("fri_small.json")


def make_env(algebra, univariate, extension_field, ntt, merkle, ip, fri, salted_merkle=None):
    env = types.SimpleNamespace()
    env.salted_merkle = salted_merkle
    env.BaseField, env.BaseFieldElement = algebra.BaseField, algebra.BaseFieldElement
    env.Polynomial = univariate.Polynomial
    env.ExtensionField, env.ExtensionFieldElement = extension_field.ExtensionField, extension_field.ExtensionFieldElement
    env.ntt, env.intt = ntt.ntt, ntt.intt
    env.fast_coset_evaluate, env.fast_coset_interpolate = ntt.fast_coset_evaluate, ntt.fast_coset_interpolate
    env.fast_multiply = getattr(ntt, "fast_multiply", None)  # the reference's own helper; not mirrored
    env.Merkle, env.ProofStream, env.Fri = merkle.Merkle, ip.ProofStream, fri.Fri
    env.field = env.BaseField.main()
    env.xfield = env.ExtensionField.main()
    env.xbf = env.xfield.modulus.coefficients[0].field
    return env


def vals(xs):
    return [x.value for x in xs]


def triples(xs):
    out = []
    for x in xs:
        c = [co.value for co in x.polynomial.coefficients]
        out.append(c + [0] * (3 - len(c)))
    return out


def bdig(xs):
    return bfe_digest(np.array(vals(xs), dtype=np.uint64))


def xdig(xs):
    return xfe_digest(np.array(triples(xs), dtype=np.uint64).T.reshape(3, -1))


def X(env, *c):
    return env.ExtensionFieldElement(env.Polynomial([env.BaseFieldElement(v, env.xbf) for v in c]), env.xfield)


def rand_bfe_list(env, seed, n):
    R = random.Random(seed)
    return [env.BaseFieldElement(R.randrange(P), env.field) for _ in range(n)]


def rand_xfe_list(env, seed, n):
    R = random.Random(seed)
    return [X(env, R.randrange(P), R.randrange(P), R.randrange(P)) for _ in range(n)]


# ---------------------------------------------------------------------------------------------
def case_ntt_golden(env, max_log=12):
    w8 = env.field.primitive_nth_root(8)
    v = [env.BaseFieldElement(i, env.field) for i in range(1, 9)]
    y = env.ntt(w8, v)
    assert vals(y) == S["gv1_ntt8"] and vals(env.intt(w8, y)) == list(range(1, 9))
    assert all(type(t) is env.BaseFieldElement and t.field is env.field for t in y)
    for logn in range(1, max_log + 1):
        n = 1 << logn
        w = env.field.primitive_nth_root(n)
        x = rand_bfe_list(env, logn, n)
        e = S["ntt_bfe"][str(logn)]
        assert bdig(env.ntt(w, x)) == e["ntt_digest"]
        assert bdig(env.intt(w, x)) == e["intt_digest"]
    for logn in range(1, min(max_log, 12) + 1):
        n = 1 << logn
        w = env.xfield.lift(env.field.primitive_nth_root(n))
        x = rand_xfe_list(env, 100 + logn, n)
        y = env.ntt(w, x)
        assert xdig(y) == S["ntt_xfe"][str(logn)]["ntt_digest"]
        assert all(t.field is env.xfield for t in y)
        assert all(c.field is env.xbf for t in y for c in t.polynomial.coefficients)


def case_ntt_quirks(env):
    f = env.field
    one_elem = [f(5)]
    assert env.ntt(f.primitive_nth_root(2), one_elem) is one_elem  # code/ntt.py:8-9
    empty = []
    assert env.ntt(f.primitive_nth_root(2), empty) is empty
    assert env.intt(f.one(), one_elem) is one_elem  # code/ntt.py:32-33
    with pytest.raises(AssertionError):
        env.ntt(f.primitive_nth_root(8), [f(i) for i in range(6)])
    with pytest.raises(AssertionError):
        env.ntt(f.primitive_nth_root(16), [f(i) for i in range(8)])
    with pytest.raises(AssertionError):
        env.ntt(f.primitive_nth_root(4), [f(i) for i in range(8)])
    with pytest.raises(AssertionError):
        env.intt(f.primitive_nth_root(4), [f(i) for i in range(8)])
    with pytest.raises(AssertionError):
        env.intt(f.primitive_nth_root(16), [f(i) for i in range(8)])


def case_coset_and_poly(env):
    Fri = env.Fri
    for key, e in S["coset"].items():
        logn, m = (int(t) for t in key.split("_"))
        n = 1 << logn
        dom = Fri.Domain(env.field.generator(), env.field.primitive_nth_root(n), n)
        poly = env.Polynomial(rand_bfe_list(env, 300 + logn, m))
        ev = dom.evaluate(poly)
        assert bdig(ev) == e["evaluate_digest"] and len(poly.coefficients) == m
        xev = dom.xevaluate(env.Polynomial(rand_xfe_list(env, 400 + logn, m)), env.xfield)
        assert xdig(xev) == e["xevaluate_digest"]
        if "interpolate_digest" in e:
            ip = dom.interpolate(rand_bfe_list(env, 500 + logn, n))
            assert type(ip) is env.Polynomial and len(ip.coefficients) == e["interpolate_len"]
            assert bdig(ip.coefficients) == e["interpolate_digest"]
            xip = dom.xinterpolate(rand_xfe_list(env, 600 + logn, n))
            assert xdig(xip.coefficients) == e["xinterpolate_digest"]
    two = env.BaseFieldElement(2, env.field)
    poly = env.Polynomial(rand_bfe_list(env, 777, 300))
    assert bdig(env.fast_coset_evaluate(poly, two, env.field.primitive_nth_root(512), 512)) == \
        S["coset_offset2"]["digest"]
    # Polynomial.scale / evaluate_domain
    poly = env.Polynomial(rand_bfe_list(env, 701, 37))
    sc = poly.scale(env.BaseFieldElement(S["scale_bfe"]["factor"], env.field))
    assert type(sc) is env.Polynomial and bdig(sc.coefficients) == S["scale_bfe"]["digest"]
    xpoly = env.Polynomial(rand_xfe_list(env, 702, 29))
    assert xdig(xpoly.scale(X(env, *S["scale_xfe"]["factor"])).coefficients) == S["scale_xfe"]["digest"]
    assert bdig(poly.evaluate_domain(rand_bfe_list(env, 703, 50))) == S["evaluate_domain_bfe"]["digest"]
    assert xdig(xpoly.evaluate_domain(rand_xfe_list(env, 704, 41))) == S["evaluate_domain_xfe"]["digest"]
    assert poly.evaluate_domain([]) == []
    # structured domains (BASELINE config 3): points offset * omega^k are routed to ONE coset transform; the values
    # are the reference's (golden) and those of the generic running-power path on the same points in another order
    w64 = env.xfield.lift(env.field.primitive_nth_root(64))
    cos = [env.xfield.lift(env.field.generator()) * (w64 ^ i) for i in range(64)]
    xp64 = env.Polynomial(rand_xfe_list(env, 705, 64))
    g = getattr(env, "glue", None)
    before = g.coset_routed if g is not None else 0
    on_coset = xp64.evaluate_domain(cos)
    assert xdig(on_coset) == S["evaluate_domain_xfe_coset64"]["digest"]
    assert all(v.field is env.xfield for v in on_coset) and pickle.dumps(on_coset[3]) == pickle.dumps(xp64.evaluate(cos[3]))
    swapped = [cos[1], cos[0]] + cos[2:]  # not a coset in this order: generic path
    assert triples(xp64.evaluate_domain(swapped)) == triples([on_coset[1], on_coset[0]] + on_coset[2:])
    rotated = cos[5:] + cos[:5]  # the coset of offset * omega^5
    assert triples(xp64.evaluate_domain(rotated)) == triples(on_coset[5:] + on_coset[:5])
    bcos = [env.field.generator() * (env.field.primitive_nth_root(64) ^ i) for i in range(64)]
    assert vals(poly.evaluate_domain(bcos)) == [poly.evaluate(b).value for b in bcos]
    short = env.Polynomial(rand_bfe_list(env, 706, 100))  # more coefficients than points: generic path
    assert vals(short.evaluate_domain(bcos[:2])) == [short.evaluate(b).value for b in bcos[:2]]
    if g is not None:
        assert g.coset_routed == before + 3
    # fast_multiply rides on the patched ntt/intt (code/ntt.py:45-79)
    a = env.Polynomial(rand_bfe_list(env, 31, 20))
    b = env.Polynomial(rand_bfe_list(env, 32, 25))
    if env.fast_multiply is not None:
        fm = env.fast_multiply(a, b, env.field.primitive_nth_root(64), 64)
        assert vals(fm.coefficients) == vals((a * b).coefficients)


def case_merkle(env):
    Merkle = env.Merkle
    g = M["gv2"]
    leaves = [X(env, *[v for v in c][:k]) for c, k in zip(g["leaves"], (3, 3, 1, 0))]
    t = Merkle(leaves)
    assert t.root().hex() == g["root"] and [b.hex() for b in t.open(2)] == g["open2"]
    assert t.num_leafs == 4 and t.depth == 2 and len(t.nodes) == 8
    assert [t.nodes[k].hex() for k in range(8)] == g["nodes"]
    assert t.leafs[1] is leaves[1] and t.nodes[3] is t.nodes[3] and t.open(0)[1] is t.open(1)[1]
    for i in range(4):
        assert Merkle.verify(t.root(), i, t.open(i), leaves[i])
    assert not Merkle.verify(t.root(), 1, t.open(1), leaves[2])
    # base-field leaves
    from test_oracle import blob_leaves, tree_values
    for logn in (0, 1, 5, 8):
        planes = tree_values(logn)
        bl = [env.BaseFieldElement(int(v), env.field) for v in planes[0]]
        assert Merkle(bl).root().hex() == M["bfe_trees"][str(logn)]["root"]
        xl = [X(env, *[int(planes[j][i]) for j in range(3)]) for i in range(1 << logn)]
        # X() does not trim; the constructor does (code/extension_field.py:6-9)
        tx = Merkle(xl)
        assert tx.root().hex() == M["xfe_trees"][str(logn)]["root"]
        assert [b.hex() for b in tx.open((1 << logn) - 1)] == M["xfe_trees"][str(logn)]["open_last"]
    # arbitrary picklable leaves, non power of two counts (code/test_merkle.py:57-61)
    for n in (1, 2, 3, 5, 13, 100):
        e = M["blob_trees"][str(n)]
        lv = blob_leaves(n)
        tb = Merkle(lv)
        assert tb.root().hex() == e["root"] and tb.depth == e["depth"]
        assert [b.hex() for b in tb.open(0)] == e["open0"]
        assert [b.hex() for b in tb.open(n - 1)] == e["open_last"]
        assert hashlib.sha256(b"".join(tb.nodes[1:])).hexdigest() == e["nodes_sha256"]
        assert all(Merkle.verify(tb.root(), i, tb.open(i), lv[i]) for i in range(n))
    # non-canonical identity (a lifted element carries a foreign BaseField object): host pickling path
    lifted = [env.xfield.lift(env.BaseFieldElement(i + 1, env.field)) for i in range(4)]
    tl = Merkle(lifted)
    assert tl.nodes[4] == hashlib.blake2b(pickle.dumps(lifted[0])).digest()
    with pytest.raises(IndexError):
        Merkle([]).root()


def fri_input(env, logn, expansion, seed):
    """make_golden.py fri_input: plain-int coset NTT of a random polynomial, canonical objects"""
    from oracle import oracle as orc
    from util import rand_xfe, root_of_unity
    n = 1 << logn
    coeffs = rand_xfe(seed, n // expansion)
    planes = orc.coset_evaluate(7, root_of_unity(logn), coeffs, n)
    return [X(env, *[int(planes[j][i]) for j in range(3)]) for i in range(n)], planes


def run_fri_case(env, logn, expansion, s, seed, e):
    n = 1 << logn
    fri = env.Fri(env.field.generator(), env.field.primitive_nth_root(n), n, expansion, s, env.xfield)
    cw, planes = fri_input(env, logn, expansion, seed)
    assert xfe_digest(planes) == e["codeword_digest"]
    ps = env.ProofStream()
    top = fri.prove(cw, ps)
    ser = ps.serialize()
    assert top == e["top_level_indices"]
    assert len(ps.objects) == e["num_objects"]
    assert [o.hex() for o in ps.objects if isinstance(o, bytes)] == e["round_roots"]
    assert len(ser) == e["transcript_len"]
    assert hashlib.sha256(ser).hexdigest() == e["transcript_sha256"]
    return fri, cw, ser


def case_fri_transcripts(env, logs=(4, 5, 6, 8, 10)):
    for logn in logs:
        e = FS["gv6"][str(logn)]
        fri, cw, ser = run_fri_case(env, logn, 4, 8, 200 + logn, e)
        if logn <= 8:
            assert fri.verify(env.ProofStream().deserialize(ser), env.Merkle(cw).root()) is True
    run_fri_case(env, 12, 32, 40, 1312, FS["exp32_s40_12"])  # BASELINE config 4's "40 checks" needs expansion 32


def case_test_fri_config(env):
    """code/test_fri.py:5-59 with the transcript additionally pinned byte for byte"""
    xfield, field = env.xfield, env.field
    n = 1024
    fri = env.Fri(field.generator(), field.primitive_nth_root(n), n, 16, 17, xfield)
    polynomial = env.Polynomial([xfield(i) for i in range(64)])
    codeword = fri.domain.xevaluate(polynomial)
    e = FS["test_fri"]
    assert xdig(codeword) == e["codeword_digest"]
    root = env.Merkle(codeword).root()
    assert root.hex() == e["root0"]
    ps = env.ProofStream()
    top = fri.prove(codeword, ps)
    ser = ps.serialize()
    assert top == e["top_level_indices"] and len(ps.objects) == e["num_objects"]
    assert (len(ser), hashlib.sha256(ser).hexdigest()) == (e["transcript_len"], e["transcript_sha256"])
    assert fri.verify(ps, root) is True
    ps = env.ProofStream()
    for i in range(0, 63 // 3):
        codeword[i] = xfield.zero()  # in-place mutation of the same list object
    fri.prove(codeword, ps)
    assert not fri.verify(ps, [])


def case_gv3(env):
    R = random.Random(5)
    n = 256
    fri = env.Fri(env.field.generator(), env.field.primitive_nth_root(n), n, 4, 8, env.xfield)
    poly = env.Polynomial([env.xfield.sample(bytes(R.getrandbits(8) for _ in range(30))) for _ in range(64)])
    cw = fri.domain.xevaluate(poly, env.xfield)
    e = FS["gv3"]
    assert xdig(cw) == e["codeword_digest"]
    ps = env.ProofStream()
    assert fri.prove(cw, ps) == e["top_level_indices"]
    ser = ps.serialize()
    assert (len(ser), hashlib.sha256(ser).hexdigest()) == (e["transcript_len"], e["transcript_sha256"])


def case_fri_errors(env):
    f, xf = env.field, env.xfield
    with pytest.raises(AssertionError):  # code/fri.py:69-72: 40 checks with expansion 4
        fri = env.Fri(f.generator(), f.primitive_nth_root(64), 64, 4, 40, xf)
        cw, _ = fri_input(env, 6, 4, 206)
        fri.prove(cw, env.ProofStream())
    with pytest.raises(AssertionError):  # code/fri.py:179-180
        fri = env.Fri(f.generator(), f.primitive_nth_root(64), 64, 4, 8, xf)
        fri.prove([xf.zero()] * 32, env.ProofStream())
    with pytest.raises(AssertionError):  # code/fri.py:52
        env.Fri(f.generator(), f.primitive_nth_root(4), 4, 4, 1, xf)
    with pytest.raises(IndexError):  # single-round FRI: codewords[1] does not exist (code/fri.py:186-187)
        fri = env.Fri(f.generator(), f.primitive_nth_root(8), 8, 4, 8, xf)
        cw, _ = fri_input(env, 3, 4, 203)
        fri.prove(cw, env.ProofStream())


def case_combination(env, glue):
    """the nonlinear combination block of BrainfuckStark.prove (code/brainfuck_stark.py:241-298)
    against the outputs of the reference's own statements (tests/golden/combination.json): values, the
    pickle of the resulting list (object graph and field identities) and the Merkle root over it."""
    from stark_brainfuck_b200.glue import DeviceCodeword
    for c in golden("combination.json")["cases"]:
        N = c["N"]
        dom = env.Fri.Domain(env.field(c["offset"]), env.field(c["omega"]), N)
        rnd = [X(env, *t) for t in c["randomizer"]]
        base = [[env.BaseFieldElement(v, env.field) for v in col] for col in c["base"]]
        ext = [[X(env, *t) for t in col] for col in c["extension"]]
        quo = [[X(env, *t) for t in col] for col in c["quotient"]]
        weights = [X(env, *t) for t in c["weights"]]
        args = (c["base_degree_bounds"], c["extension_degree_bounds"], c["quotient_degree_bounds"])
        for attach in (False, True):
            scope = glue.keep_planes()
            scope.__enter__()
            if attach:  # codewords that still have their device planes (what the drop-in's own ops return)
                up = glue.engine.upload
                for col in base:
                    glue.remember_planes(col, up(np.array(vals(col), dtype=np.uint64)))
                for col in ext + quo:
                    glue.remember_planes(col, up(np.array(triples(col), dtype=np.uint64).T.copy()))
                assert all(glue.planes_of(col) is not None for col in base + ext + quo)
            out = glue.combination_codeword(dom, env.xfield, c["max_degree"], rnd, base, args[0], ext, args[1], quo,
                                            args[2], weights)
            assert isinstance(out, DeviceCodeword) and len(out) == N
            tree = env.Merkle(out)
            assert tree.root().hex() == c["root"]
            assert tree.leafs[3] is out[3]  # the tree serves the codeword's own (lazily built) objects
            assert triples(out[5:7]) == c["out"][5:7]
            full = out.materialize()
            assert triples(full) == c["out"]
            assert hashlib.sha256(pickle.dumps(full)).hexdigest() == c["out_pickle_sha256"]
            if attach and quo:  # a list whose ends were replaced is not read from its stale planes
                quo[0][0] = X(env, 1, 2, 3)
                assert glue.planes_of(quo[0]) is None
                quo[0][0] = X(env, *c["quotient"][0][0])
            scope.__exit__(None, None, None)
            assert all(glue.planes_of(col) is None for col in base + ext + quo)  # nothing outlives the scope
        # weights whose coefficients carry a foreign field object: the reference's result would carry it too
        odd = [env.ExtensionFieldElement(env.Polynomial([env.BaseFieldElement(5, env.field)]), env.xfield)] + weights[1:]
        assert glue.combination_codeword(dom, env.xfield, c["max_degree"], rnd, base, args[0], ext, args[1], quo, args[2],
                                         odd) is None
        with pytest.raises(AssertionError):
            glue.combination_codeword(dom, env.xfield, c["max_degree"], rnd, base, args[0], ext, args[1], quo, args[2],
                                      weights[:-1])


def seeded_urandom(seed):
    U = random.Random(seed)
    return lambda n: bytes(U.getrandbits(8) for _ in range(n))


def case_lde(env, glue, make_table=None):
    """Table.lde / ldex (code/table.py:112-148) against the reference's codewords
    (tests/golden/lde.json): values and pickles (object graph incl. the shared coefficient objects of constant
    columns).  `make_table` builds a real reference Table where one is available; otherwise a stub with the
    attributes the reference's methods read."""
    g = golden("lde.json")
    N = g["N"]
    dom = env.Fri.Domain(env.field(g["offset"]), env.field(g["omega"]), N)
    for c in g["cases"]:
        h, bw, fw = c["height"], c["base_width"], c["full_width"]
        if make_table is not None:
            t = make_table(env.field, bw, fw, c["length"], c["num_randomizers"], env.field(g["omega"]), N)
            assert t.height == h and t.omicron.value == c["omicron"]
        else:
            t = types.SimpleNamespace(field=env.field, base_width=bw, full_width=fw, length=c["length"], height=h,
                                      num_randomizers=c["num_randomizers"], omicron=env.field(c["omicron"]))
        base = [[env.BaseFieldElement(v, env.field) for v in row] for row in c["base"]]
        ext = [[X(env, *v) for v in row] for row in c["ext"]]
        draw = seeded_urandom(c["urandom_seed"])
        t.matrix = [list(row) for row in base]
        with glue.keep_planes():
            base_cw = glue.table_lde(t, dom, draw)
            assert [vals(cw) for cw in base_cw] == c["base_codewords"]
            assert hashlib.sha256(pickle.dumps(base_cw)).hexdigest() == c["base_pickle_sha256"]
            if h:
                assert all(glue.planes_of(cw) is not None for cw in base_cw)
        t.field = env.xfield
        t.matrix = [[env.xfield.lift(v) for v in base[r]] + ext[r] for r in range(h)]
        ext_cw = glue.table_lde(t, dom, draw, xfield=env.xfield)
        assert [triples(cw) for cw in ext_cw] == c["ext_codewords"]
        assert hashlib.sha256(pickle.dumps(ext_cw)).hexdigest() == c["ext_pickle_sha256"]
        # interpolate_columns: the polynomials take the table's values at its points (incl. the drawn randomizers)
        if h:
            draw2 = seeded_urandom(5)
            polys = glue.table_interpolate_columns(t, dom.omega, N, range(bw, fw), draw2)
            draw3 = seeded_urandom(5)
            for k, p in enumerate(polys):
                assert len(p.coefficients) <= h + t.num_randomizers
                want = [ext[r][k] for r in range(h)] + [env.xfield.sample(draw3(24)) for _ in range(t.num_randomizers)]
                pts = [t.omicron ^ i for i in range(h)] + [dom.omega ^ (2 * i + 1) for i in range(t.num_randomizers)]
                assert triples([p.evaluate(env.xfield.lift(x)) for x in pts]) == triples(want)
    with pytest.raises(AssertionError):  # omega does not have claimed order (code/table.py:113-114)
        glue.table_interpolate_columns(t, dom.omega, N // 2, range(1), seeded_urandom(1))


def case_lazy_codewords(env, glue):
    """The device views BrainfuckStark.prove() works on under the drop-in (glue.keep_planes(lazy=True)): lde / ldex /
    Domain.xevaluate return DeviceCodewords whose elements, rows and lifted views are built on first access -- same
    values, same pickles (tests/golden/lde.json), same objects on repeated access -- and the salted tree over their
    rows equals the tree over the materialised rows."""
    from stark_brainfuck_b200.glue import DeviceCodeword, LazyLeafs, LazyRows
    g = golden("lde.json")
    N = g["N"]
    dom = env.Fri.Domain(env.field(g["offset"]), env.field(g["omega"]), N)
    seen = shared = 0
    for c in g["cases"]:
        h, bw, fw = c["height"], c["base_width"], c["full_width"]
        if not h:
            continue
        t = types.SimpleNamespace(field=env.field, base_width=bw, full_width=fw, length=c["length"], height=h,
                                  num_randomizers=c["num_randomizers"], omicron=env.field(c["omicron"]))
        base = [[env.BaseFieldElement(v, env.field) for v in row] for row in c["base"]]
        ext = [[X(env, *v) for v in row] for row in c["ext"]]
        draw = seeded_urandom(c["urandom_seed"])
        t.matrix = [list(row) for row in base]
        with glue.keep_planes(lazy=True):
            base_cw = glue.table_lde(t, dom, draw)
            assert all(type(cw) is DeviceCodeword and cw.kind == "b" and len(cw) == N for cw in base_cw)
            t.field = env.xfield
            t.matrix = [[env.xfield.lift(v) for v in base[r]] + ext[r] for r in range(h)]
            ext_cw = glue.table_lde(t, dom, draw, xfield=env.xfield)
            assert all(type(cw) is DeviceCodeword and cw.kind == "x" for cw in ext_cw)
            seen += len(ext_cw)
            shared += sum(1 for cw in ext_cw if cw._share)  # constant columns: outputs share coefficient objects
            # the transposition of prove() and the lifted views of Table.extend, before anything is materialised
            rows = glue.rows_of(base_cw + ext_cw)
            assert type(rows) is LazyRows and len(rows) == N
            lifted = glue.lift_codewords(env.xfield, base_cw)
            assert all(type(cw) is DeviceCodeword and cw.kind == "l" for cw in lifted)
            r5 = rows[5]
            assert rows[5] is r5 and rows[-N + 5] is r5 and type(r5) is tuple and len(r5) == fw
            assert all(r5[k] is (base_cw + ext_cw)[k][5] for k in range(fw))
            assert vals(r5[:bw]) == [cw[5] for cw in c["base_codewords"]]
            assert triples(r5[bw:]) == [cw[5] for cw in c["ext_codewords"]]
            l5 = lifted[0][5]
            want = env.xfield.lift(base_cw[0][5])
            assert lifted[0][5] is l5 and pickle.dumps(l5) == pickle.dumps(want)
            assert (l5.polynomial.coefficients[0] is base_cw[0][5]) if base_cw[0][5].value else l5.polynomial.coefficients == []
            assert rows[2:4] == [rows[2], rows[3]] and base_cw[0][1:3] == [base_cw[0][1], base_cw[0][2]]
            with pytest.raises(IndexError):
                rows[N]
            with pytest.raises(IndexError):
                base_cw[0][N]
            # salted tree over the lazy rows == tree over the same rows as host tuples (same salts)
            if env.salted_merkle is not None and N & (N - 1) == 0:
                sm = env.salted_merkle
                old = sm.urandom
                try:
                    sm.urandom = seeded_urandom(77)
                    lazy_tree = sm.SaltedMerkle(rows)
                    assert type(lazy_tree.leafs) is LazyLeafs and len(lazy_tree.leafs) == N
                    leaf3 = lazy_tree.leafs[3]
                    assert lazy_tree.leafs[3] is leaf3 and leaf3[0] is rows[3]
                    salt, path = lazy_tree.open(3)
                    assert salt is leaf3[1] and sm.SaltedMerkle.verify(lazy_tree.root(), 3, salt, path, leaf3[0])
                    sm.urandom = seeded_urandom(77)
                    host_tree = sm.SaltedMerkle([tuple(r) for r in rows])
                    assert host_tree.root() == lazy_tree.root()
                    assert [s_ for _, s_ in host_tree.leafs] == lazy_tree.leafs.salts
                    assert host_tree.open(N - 1) == lazy_tree.open(N - 1)
                finally:
                    sm.urandom = old
            # whole codewords through the views: the reference's values and pickles
            assert [vals(cw) for cw in base_cw] == c["base_codewords"]
            assert hashlib.sha256(pickle.dumps([list(cw) for cw in base_cw])).hexdigest() == c["base_pickle_sha256"]
            assert [triples(cw) for cw in ext_cw] == c["ext_codewords"]
            assert hashlib.sha256(pickle.dumps([list(cw) for cw in ext_cw])).hexdigest() == c["ext_pickle_sha256"]
            assert list(rows)[5] is r5 and list(lifted[0])[5] is l5 and len(list(rows)) == N
            # a randomizer-style codeword (code/brainfuck_stark.py:164-167) and its use as transform input
            poly = env.Polynomial(rand_xfe_list(env, 3, N // 4))
            rc = glue.domain_xevaluate(dom, poly)
            assert type(rc) is DeviceCodeword and rc.kind == "x"
            back = glue.domain_xinterpolate(dom, rc)
            assert triples(back.coefficients[:N // 4]) == triples(poly.coefficients)
        # outside the scope everything is a plain list again
        assert type(glue.domain_xevaluate(dom, poly)) is list
    assert seen > 0 and shared > 0


def case_quotients_glue(env, glue):
    """SURVEY 8(f) row 1 through the glue (object lists in, codewords out) against the reference's quotient
    codewords (tests/golden/quotients.json): plain lists outside keep_planes(), lazy device codewords and the
    per-table plane cache inside."""
    from stark_brainfuck_b200.glue import (DeviceCodeword, ZEROFIER_BOUNDARY, ZEROFIER_TERMINAL, ZEROFIER_TRANSITION)
    g = golden("quotients.json")
    N, W = g["N"], g["width"]
    dom = env.Fri.Domain(env.field(g["offset"]), env.field(g["omega"]), N)

    def constraints(program):
        return [types.SimpleNamespace(dictionary={tuple(k): X(env, *c) for k, c in cons}) for cons in program]
    for scoped in (False, True):
        scope = glue.keep_planes() if scoped else None
        if scope:
            scope.__enter__()
        for t in g["tables"]:
            cws = [[X(env, *v) for v in col] for col in t["codewords"]]
            if scoped:  # one column as if it had come out of a device op
                glue.remember_planes(cws[1], glue.engine.upload(np.array(t["codewords"][1], dtype=np.uint64).T.copy()))
            for kind, name in ((ZEROFIER_BOUNDARY, "boundary"), (ZEROFIER_TRANSITION, "transition"),
                               (ZEROFIER_TERMINAL, "terminal")):
                out = glue.quotient_codewords(dom, cws, W, constraints(t[name]["program"]), kind, height=t["height"],
                                              omicron_inv=t["omicron_inv"],
                                              shift=t["unit_distance"] if kind == ZEROFIER_TRANSITION else 0)
                assert all(isinstance(q, DeviceCodeword if scoped else list) for q in out)
                assert [triples(list(q)) for q in out] == t[name]["out"], name
        if scope:
            assert any(isinstance(k, tuple) for k in glue._kept)  # the assembled table planes were kept
            scope.__exit__(None, None, None)


def case_zerofier_decision(env, glue):
    """Glue._zerofier_may_vanish (the host-side replacement of the device flag's read-back) against brute force over
    small domains: whenever it answers False, no zerofier of code/table.py:161-163, :194-201, :256-259 has a root on
    offset * <omega>; and the reference's own configuration (offset = the field's generator) takes the fast path."""
    from stark_brainfuck_b200.glue import ZEROFIER_BOUNDARY, ZEROFIER_TERMINAL, ZEROFIER_TRANSITION
    P = env.field.p
    gen = env.field.generator().value
    skipped = 0
    for log_n in (2, 3, 5):
        N = 1 << log_n
        w = env.field.primitive_nth_root(N).value
        roots = [pow(w, i, P) for i in range(N)]
        for offset in (gen, 1, w, pow(w, 3, P), pow(gen, N, P), pow(env.field.primitive_nth_root(2 * N).value, 1, P), 7):
            dom = env.Fri.Domain(env.field(offset), env.field(w), N)
            pts = [offset * r % P for r in roots]
            for height in (0, 1, 2, N // 2, N, 3, 2 * N):
                omicron = env.field.primitive_nth_root(height).value if height in (1, 2, N // 2, N, 2 * N) and height else 1
                for oinv in (pow(omicron, P - 2, P), offset, 5):
                    truth = {ZEROFIER_BOUNDARY: any(x == 1 for x in pts),
                             ZEROFIER_TRANSITION: height == 0 or any(pow(x, height, P) == 1 for x in pts),
                             ZEROFIER_TERMINAL: any(x == oinv for x in pts)}
                    for kind, vanishes in truth.items():
                        may = glue._zerofier_may_vanish(dom, kind, height, oinv)
                        assert may or not vanishes, (N, offset, kind, height, oinv)
                        skipped += not may
    assert skipped > 100
    big = env.Fri.Domain(env.field(gen), env.field.primitive_nth_root(1 << 20), 1 << 20)
    oinv = env.field.primitive_nth_root(1 << 16).inverse().value
    assert not any(glue._zerofier_may_vanish(big, k, 1 << 16, oinv)
                   for k in (ZEROFIER_BOUNDARY, ZEROFIER_TRANSITION, ZEROFIER_TERMINAL))


def salted_rows(env, case):
    """the rows of a tests/golden/salted.json case, rebuilt with one BaseField object per field_id"""
    fields, cols = {}, []
    for c in case["columns"]:
        if c["kind"] == "b":
            f = fields.setdefault(c["field_id"], env.BaseField.main())
            cols.append([env.BaseFieldElement(v, f) for v in c["values"]])
        else:
            cols.append([X(env, *t) for t in c["values"]])
    return cols, list(zip(*cols))


def case_salted(env, glue=None):
    """SaltedMerkle over zipped rows (code/salted_merkle.py, code/brainfuck_stark.py:178-199)
    against the reference's own trees (tests/golden/salted.json): roots, every node, openings."""
    sm = env.salted_merkle
    old = sm.urandom
    try:
        for case in golden("salted.json")["cases"]:
            cols, rows = salted_rows(env, case)
            n = case["n"]
            for kept in (False, True) if glue is not None else (False,):
                R = random.Random(case["salt_seed"])
                sm.urandom = lambda k: bytes(R.getrandbits(8) for _ in range(k))
                if kept:  # columns that are already on the device, as inside BrainfuckStark.prove
                    with glue.keep_planes():
                        for col in cols:
                            arr = (glue.B.bfe_to_np(col).reshape(1, -1) if glue.B.is_bfe(col[0]) else glue.B.xfe_to_np(col))
                            glue.remember_planes(col, glue.engine.upload(arr))
                        t = sm.SaltedMerkle(rows)
                else:
                    t = sm.SaltedMerkle(rows)
                assert [s_.hex() for _, s_ in t.leafs] == case["salts"]
                assert t.root().hex() == case["root"], case["name"]
                assert [t.nodes[n + i].hex() for i in range(n)] == case["leaf_digests"]
                assert hashlib.sha256(b"".join(t.nodes[1:])).hexdigest() == case["nodes_sha256"]
                salt, path = t.open(case["open_index"])
                assert [d.hex() for d in path] == case["open_path"] and salt.hex() == case["salts"][case["open_index"]]
                assert sm.SaltedMerkle.verify(t.root(), case["open_index"], salt, path, rows[case["open_index"]])
                assert t.leafs[0][0] is rows[0] and t.num_leafs == n and t.depth == n.bit_length() - 1
        # leaves the row templates cannot express (not tuples of field elements; not a power of two): host pickling
        R = random.Random(5)
        sm.urandom = lambda k: bytes(R.getrandbits(8) for _ in range(k))
        for data in ([("a", 1), ("b", 2), ("c", 3)], [env.BaseFieldElement(i, env.field) for i in range(4)],
                     [(env.BaseFieldElement(i, env.field), "x") for i in range(2)]):
            t = sm.SaltedMerkle(data)
            want = [hashlib.blake2b(pickle.dumps(e) + pickle.dumps(s_)).digest() for e, s_ in t.leafs]
            npo2 = 1 << t.depth
            assert [t.nodes[npo2 + i] for i in range(len(data))] == want
            salt, path = t.open(1)
            assert sm.SaltedMerkle.verify(t.root(), 1, salt, path, data[1])
        if glue is not None:
            # leaves that are a codeword living on the device (this used to hit an unbound local)
            from stark_brainfuck_b200.glue import DeviceCodeword
            xs = rand_xfe_list(env, 9, 8)
            dc = DeviceCodeword(glue, glue.engine.upload(glue.B.xfe_to_np(xs)), env.xfield)
            t = sm.SaltedMerkle(dc)
            want = [hashlib.blake2b(pickle.dumps(e) + pickle.dumps(s_)).digest() for e, s_ in t.leafs]
            assert [t.nodes[8 + i] for i in range(8)] == want and triples([e for e, _ in t.leafs]) == triples(xs)
    finally:
        sm.urandom = old
