"""GPU parity, array level: every C-ABI entry point against the CPU oracle on seeded inputs
and against the golden digests produced by the unmodified reference."""
import hashlib
import pickle
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as orc  # noqa: E402
from util import P, bfe_digest, golden, have_golden, rand_bfe, rand_xfe, root_of_unity, xfe_digest  # noqa: E402

S = golden("small.json")
M = golden("merkle.json")
F = golden("fold.json")


@pytest.fixture(scope="module")
def eng():
    from stark_brainfuck_b200 import Engine
    return Engine(0)


def tpl_pair():
    from stark_brainfuck_b200.marshal import templates_from_marker_pickles
    x = templates_from_marker_pickles([bytes.fromhex(M["xfe_marker_pickles"][str(k)]) for k in range(4)], 3, True)
    b = templates_from_marker_pickles([bytes.fromhex(M["bfe_marker_pickle"])], 1, False)
    return x, b


@pytest.mark.parametrize("logn", list(range(0, 23)))
def test_ntt_bfe_vs_oracle(eng, logn):
    n = 1 << logn
    x = rand_bfe(logn, n)
    w = root_of_unity(logn) if logn else 1
    d = eng.upload(x)
    y = eng.download(eng.ntt(d, logn, w))[0]
    assert np.array_equal(y, orc.ntt(w, x))
    z = eng.download(eng.ntt(d, logn, w, inverse=True))[0]
    assert np.array_equal(z, orc.intt(w, x))
    # bit-exact round trip, in place
    t = eng.ntt(d, logn, w)
    eng.ntt(t, logn, w, inverse=True, out=t)
    assert np.array_equal(eng.download(t)[0], x)
    e = S["ntt_bfe"].get(str(logn))
    if e:
        assert bfe_digest(y) == e["ntt_digest"] and bfe_digest(z) == e["intt_digest"]


@pytest.mark.parametrize("logn", (16, 18, 20))
def test_ntt_bfe_big_golden(eng, logn):
    g = golden("ntt_big.json")
    if str(logn) not in g:
        pytest.skip("golden not generated")
    x = rand_bfe(logn, 1 << logn)
    w = root_of_unity(logn)
    d = eng.upload(x)
    assert bfe_digest(eng.download(eng.ntt(d, logn, w))[0]) == g[str(logn)]["ntt_digest"]
    if "intt_digest" in g[str(logn)]:
        assert bfe_digest(eng.download(eng.ntt(d, logn, w, inverse=True))[0]) == g[str(logn)]["intt_digest"]


@pytest.mark.parametrize("logn", list(range(1, 13)) + [14, 16, 18])
def test_ntt_xfe_golden(eng, logn):
    e = S["ntt_xfe"].get(str(logn)) or (golden("xntt_big.json").get(str(logn)) if have_golden("xntt_big.json") else None)
    x = rand_xfe(100 + logn, 1 << logn)
    w = root_of_unity(logn)
    y = eng.download(eng.ntt(eng.upload(x), logn, w))
    assert np.array_equal(y, orc.xntt(w, x))
    if e:
        assert xfe_digest(y) == e["ntt_digest"]


def test_ntt_other_roots_and_asserts(eng):
    # any primitive root must work, not only the canonical one
    for logn in (3, 7, 11, 13):
        n = 1 << logn
        w = pow(root_of_unity(logn), 5, P)
        x = rand_bfe(50 + logn, n)
        assert np.array_equal(eng.download(eng.ntt(eng.upload(x), logn, w))[0], orc.ntt(w, x))
    d = eng.upload(rand_bfe(1, 8))
    with pytest.raises(AssertionError):
        eng.ntt(d, 3, root_of_unity(4))  # order 16
    with pytest.raises(AssertionError):
        eng.ntt(d, 3, root_of_unity(2))  # not primitive


@pytest.mark.parametrize("logn,m", [(3, 2), (3, 8), (6, 16), (6, 17), (9, 128), (9, 512), (11, 512), (11, 513),
                                     (11, 2048), (14, 4096), (17, 1 << 15), (20, 1 << 18)])
def test_coset_evaluate_interpolate(eng, logn, m):
    n = 1 << logn
    w = root_of_unity(logn)
    c = rand_bfe(300 + logn, m)
    y = eng.download(eng.ntt(eng.upload(c), logn, w, offset=7))[0]
    assert np.array_equal(y, orc.coset_evaluate(7, w, c, n))
    key = "%d_%d" % (logn, m)
    if key in S["coset"]:
        assert bfe_digest(y) == S["coset"][key]["evaluate_digest"]
    xc = rand_xfe(400 + logn, m)
    xy = eng.download(eng.ntt(eng.upload(xc), logn, w, offset=7))
    assert np.array_equal(xy, orc.coset_evaluate(7, w, xc, n))
    if key in S["coset"]:
        assert xfe_digest(xy) == S["coset"][key]["xevaluate_digest"]
    if m == n:
        v = rand_bfe(500 + logn, n)
        ip = eng.download(eng.ntt(eng.upload(v), logn, w, offset=7, inverse=True))[0]
        assert np.array_equal(ip, orc.coset_interpolate(7, w, v))
        if "interpolate_digest" in S["coset"].get(key, {}):
            assert bfe_digest(ip) == S["coset"][key]["interpolate_digest"]
    # evaluate then interpolate returns the zero-padded coefficients
    back = eng.download(eng.ntt(eng.upload(y), logn, w, offset=7, inverse=True))[0]
    assert np.array_equal(back[:m], c) and not back[m:].any()


def test_batched_planes(eng):
    logn, q = 12, 7
    n = 1 << logn
    w = root_of_unity(logn)
    x = np.stack([rand_bfe(60 + i, n) for i in range(q)])
    y = eng.download(eng.ntt(eng.upload(x), logn, w, offset=7))
    for i in range(q):
        assert np.array_equal(y[i], orc.coset_evaluate(7, w, x[i], n))


def test_scale_and_eval_points(eng):
    c = rand_bfe(701, 37)
    f = S["scale_bfe"]["factor"]
    assert bfe_digest(eng.download(eng.scale(eng.upload(c), f))[0]) == S["scale_bfe"]["digest"]
    xc = rand_xfe(702, 29)
    assert xfe_digest(eng.download(eng.scale(eng.upload(xc), S["scale_xfe"]["factor"]))) == S["scale_xfe"]["digest"]
    pts = rand_bfe(703, 50)
    assert bfe_digest(eng.download(eng.eval_points(eng.upload(c), eng.upload(pts)))[0]) == \
        S["evaluate_domain_bfe"]["digest"]
    xpts = rand_xfe(704, 41)
    assert xfe_digest(eng.download(eng.eval_points(eng.upload(xc), eng.upload(xpts)))) == \
        S["evaluate_domain_xfe"]["digest"]
    # mixed shapes against the oracle
    got = eng.download(eng.eval_points(eng.upload(xc), eng.upload(pts)))
    lifted = np.zeros((3, 50), dtype=np.uint64)
    lifted[0] = pts
    assert np.array_equal(got, orc.eval_points(xc, lifted))
    got = eng.download(eng.eval_points(eng.upload(c), eng.upload(xpts)))
    cl = np.zeros((3, 37), dtype=np.uint64)
    cl[0] = c
    assert np.array_equal(got, orc.eval_points(cl, xpts))


def mixed_tree_values(logn):
    from test_oracle import tree_values
    return tree_values(logn)


@pytest.mark.parametrize("logn", list(range(0, 11)) + [13, 16])
def test_merkle_field_trees(eng, logn):
    xt, bt = tpl_pair()
    if logn <= 10:
        planes = mixed_tree_values(logn)
    else:
        planes = rand_xfe(100 + logn, 1 << logn)
    n = 1 << logn
    nodes = eng.download_bytes(eng.merkle_field(eng.upload(planes), xt))
    ref = orc.merkle_field(orc.templates_from_marker_pickles(
        [bytes.fromhex(M["xfe_marker_pickles"][str(k)]) for k in range(4)], 3, True), planes)
    assert nodes[64:] == ref[1:].tobytes()
    if logn <= 10:
        assert nodes[64:128].hex() == M["xfe_trees"][str(logn)]["root"]
        assert hashlib.sha256(nodes[64:]).hexdigest() == M["xfe_trees"][str(logn)]["nodes_sha256"]
    bn = eng.download_bytes(eng.merkle_field(eng.upload(planes[0].copy()), bt))
    if logn <= 10:
        assert bn[64:128].hex() == M["bfe_trees"][str(logn)]["root"]
        assert hashlib.sha256(bn[64:]).hexdigest() == M["bfe_trees"][str(logn)]["nodes_sha256"]
    # openings
    d_nodes = eng.merkle_field(eng.upload(planes), xt)
    idx = sorted({0, n - 1, n // 3})
    paths = eng.merkle_open(d_nodes, idx)
    for i, p in zip(idx, paths):
        assert p == orc.merkle_open(ref, i)


def test_merkle_gv2_and_uniform(eng):
    xt, _ = tpl_pair()
    g = M["gv2"]
    planes = np.array(g["leaves"], dtype=np.uint64).T.copy()
    d = eng.merkle_field(eng.upload(planes), xt)
    assert eng.root(d).hex() == g["root"]
    assert [b.hex() for b in eng.merkle_open(d, [2])[0]] == g["open2"]
    for logn, e in M["xfe_trees_uniform"].items():
        d = eng.merkle_field(eng.upload(rand_xfe(100 + int(logn), 1 << int(logn))), xt)
        assert eng.root(d).hex() == e["root"]


@pytest.mark.parametrize("n", (1, 2, 3, 5, 8, 13, 64, 100))
def test_merkle_blobs(eng, n):
    from test_oracle import blob_leaves
    e = M["blob_trees"][str(n)]
    blobs = [pickle.dumps(x) for x in blob_leaves(n)]
    d = eng.merkle_blobs(blobs)
    assert eng.root(d).hex() == e["root"]
    ref = orc.merkle_blobs(blobs)
    got = np.frombuffer(eng.download_bytes(d), dtype=np.uint8).reshape(-1, 64)
    npo2 = got.shape[0] // 2
    for k in range(1, npo2 + n):
        assert bytes(got[k]) == bytes(ref[k]), k


def test_merkle_blob_lengths(eng):
    R = random.Random(7)
    blobs = [bytes(R.getrandbits(8) for _ in range(L)) for L in list(range(0, 140)) + [255, 256, 257, 383, 384, 385, 1000]]
    d = eng.merkle_blobs(blobs)
    got = np.frombuffer(eng.download_bytes(d), dtype=np.uint8).reshape(-1, 64)
    npo2 = got.shape[0] // 2
    for i, b in enumerate(blobs):
        assert bytes(got[npo2 + i]) == hashlib.blake2b(b).digest(), len(b)


@pytest.mark.parametrize("logn", (1, 2, 4, 6, 8, 12, 17))
def test_fri_fold(eng, logn):
    xt, _ = tpl_pair()
    n = 1 << logn
    cw = rand_xfe(1100 + logn, n)
    R = random.Random(1100 + logn)  # make_golden.py draws alpha from a fresh Random(seed)
    alpha = [R.randrange(P) for _ in range(3)]
    w = root_of_unity(logn)
    nxt, nodes = eng.fri_fold(eng.upload(cw), alpha, 7, w, xt)
    ref = orc.fri_fold(cw, alpha, 7, w)
    assert np.array_equal(eng.download(nxt), ref)
    e = F.get("fold_%d" % logn)
    if e:
        assert alpha == e["alpha"] and xfe_digest(ref) == e["digest"]
    otpl = orc.templates_from_marker_pickles(
        [bytes.fromhex(M["xfe_marker_pickles"][str(k)]) for k in range(4)], 3, True)
    assert eng.download_bytes(nodes)[64:] == orc.merkle_field(otpl, ref)[1:].tobytes()
    nxt2, none = eng.fri_fold(eng.upload(cw), alpha, 7, w, None)
    assert none is None and np.array_equal(eng.download(nxt2), ref)
    # structured input: a zero codeword folds to zeros (k = 0 leaf template)
    z = np.zeros((3, n), dtype=np.uint64)
    nz, nn = eng.fri_fold(eng.upload(z), alpha, 7, w, xt)
    assert not eng.download(nz).any()
    assert eng.download_bytes(nn)[64:] == orc.merkle_field(otpl, np.zeros((3, n // 2), dtype=np.uint64))[1:].tobytes()


def test_gather(eng):
    x = rand_xfe(9, 1000)
    idx = [0, 999, 5, 5, 123]
    got = eng.gather(eng.upload(x), idx)
    assert np.array_equal(got, x[:, idx].T)


def test_ntt_host_path(eng):
    import torch
    logn = 14
    x = rand_bfe(logn, 1 << logn)
    h = torch.from_numpy(x.view(np.int64).reshape(1, -1).copy()).pin_memory()
    out = eng.ntt_host(h, logn, root_of_unity(logn))
    assert np.array_equal(out.numpy().view(np.uint64)[0], orc.ntt(root_of_unity(logn), x))


def test_launch_counter_moves(eng):
    a = eng.launch_count()
    eng.ntt(eng.upload(rand_bfe(1, 64)), 6, root_of_unity(6))
    assert eng.launch_count() > a


def test_quotients_golden_and_oracle(eng):
    """SURVEY 8(f) row 1 through the C ABI: the reference's golden quotient codewords, then random programs
    at larger domains against the oracle, and the vanishing-zerofier flag."""
    from util import quotient_cases
    g = golden("quotients.json")
    for name, cw, shift, prog, kind, height, oinv, want in quotient_cases(g):
        W, _, N = cw.shape
        d = eng.upload(cw.reshape(3 * W, N)).reshape(W, 3, N)
        out, vanishes = eng.quotients(d, shift, *prog, kind, height, oinv, g["offset"], g["omega"])
        assert not vanishes, name
        assert np.array_equal(eng.download(out.reshape(-1, N)).reshape(-1, 3, N), want), name
    R = random.Random(77)
    for logn, W, ncons in ((10, 5, 4), (14, 11, 3)):
        N = 1 << logn
        cw = np.stack([rand_xfe(500 + logn + j, N) for j in range(W)])
        program = []
        for _ in range(ncons):
            cons = []
            for _ in range(R.randrange(1, 12)):
                k = [0] * (2 * W)
                for _ in range(R.randrange(0, 5)):
                    k[R.randrange(2 * W)] += R.randrange(1, 9)
                cons.append([k, [R.randrange(P) for _ in range(3)]])
            program.append(cons)
        from util import quotient_program
        prog = quotient_program(program)
        for lifted in (False, True):
            if lifted:  # every other column is a lifted base-field column; some monomials get base-field coefficients
                cw[::2, 1:, :] = 0
                for cons in program:
                    for mono in cons[::2]:
                        mono[1][1] = mono[1][2] = 0
                prog = quotient_program(program)
            d = eng.upload(cw.reshape(3 * W, N)).reshape(W, 3, N)
            for kind, height in ((2, 8), (2, 0)):
                oinv = pow(root_of_unity(3), P - 2, P)
                out, vanishes = eng.quotients(d, N // 8, *prog, kind, height, oinv, 7, root_of_unity(logn))
                ref, rv = orc.quotients(cw, N // 8, *prog, kind, height, oinv, 7, root_of_unity(logn))
                assert vanishes == rv is False
                assert np.array_equal(eng.download(out.reshape(-1, N)).reshape(-1, 3, N), ref), (logn, lifted, kind, height)
    # rows are taken modulo N: a table of height 1 has unit distance N (code/table.py:37-40), and more wraps around
    for shift in (N, N + 3, 5 * N + 1):
        out, vanishes = eng.quotients(d, shift, *prog, 2, 8, oinv, 7, root_of_unity(logn))
        ref, rv = orc.quotients(cw, shift, *prog, 2, 8, oinv, 7, root_of_unity(logn))
        assert not vanishes and not rv
        assert np.array_equal(eng.download(out.reshape(-1, N)).reshape(-1, 3, N), ref), shift
    # offset 1 puts x = 1 on the domain: boundary zerofier vanishes
    name, cw, shift, prog, kind, height, oinv, want = next(iter(quotient_cases(g)))
    W, _, N = cw.shape
    _, vanishes = eng.quotients(eng.upload(cw.reshape(3 * W, N)).reshape(W, 3, N), 0, *prog, 1, 0, 1, 1, g["omega"])
    assert vanishes


@pytest.mark.parametrize("logn", (4, 5, 6, 7, 9, 11, 12, 13, 15))
def test_ntt_many_planes_16_point_core(eng, logn):
    """more than 2^19 elements per call select the 16-point core steps also for short transforms
    (every tail radix and step count of that variant); a few planes are compared with the oracle"""
    n = 1 << logn
    q = min((1 << 20) // n, 65535)  # gridDim.z limit; still more than 2^19 elements in total
    w = root_of_unity(logn)
    rng = np.random.default_rng(logn)
    x = rng.integers(0, P, size=(q, n), dtype=np.uint64, endpoint=False)
    d = eng.upload(x)
    y = eng.ntt(d, logn, w, offset=7)
    back = eng.ntt(y, logn, w, offset=7, inverse=True)
    assert np.array_equal(eng.download(back), x)
    yh = eng.download(y)
    for plane in (0, q // 2, q - 1):
        assert np.array_equal(yh[plane], orc.coset_evaluate(7, w, x[plane], n))


def test_ntt_three_pass_plan(eng):
    """2^23 = 2^8 * 2^8 * 2^7: forward vs oracle, bit-exact round trip"""
    logn = 23
    n = 1 << logn
    x = rand_bfe(23, n)
    w = root_of_unity(logn)
    d = eng.upload(x)
    y = eng.ntt(d, logn, w)
    assert np.array_equal(eng.download(y)[0], orc.ntt(w, x))
    assert np.array_equal(eng.download(eng.ntt(y, logn, w, inverse=True))[0], x)


def test_combination_golden_and_oracle(eng):
    """SURVEY 8(f) row 3 through the C ABI: the reference's golden combination codewords, then random column
    sets (base and extension columns, strided views, repeated and zero shifts, zero weights) against the oracle."""
    from util import combination_cases
    for cols, wa, wb, shifts, N, offset, omega, want in combination_cases(golden("combination.json")):
        out = eng.combination([eng.upload(c) for c in cols], wa, wb, shifts, N, offset, omega)
        assert np.array_equal(eng.download(out), want)
    R = random.Random(78)
    for logn, n_cols in ((4, 3), (9, 7), (10, 12), (13, 40), (16, 9)):
        N = 1 << logn
        w = root_of_unity(logn)
        host, dev = [], []
        big = eng.upload(rand_xfe(900 + logn, 2 * N))  # extension columns as views into a wider buffer (stride 2N)
        for c in range(n_cols):
            if c == 0:
                h = eng.download(big)[:, N // 2:N // 2 + N].copy()
                d = big[:, N // 2:N // 2 + N]
            elif c % 3 == 1:
                h = rand_bfe(910 + logn + c, N).reshape(1, N)
                d = eng.upload(h)
            else:
                h = rand_xfe(920 + logn + c, N)
                d = eng.upload(h)
            host.append(h)
            dev.append(d)
        wa = np.array([[R.randrange(P) for _ in range(3)] for _ in range(n_cols)], dtype=np.uint64)
        wb = np.array([[R.randrange(P) for _ in range(3)] for _ in range(n_cols)], dtype=np.uint64)
        wb[0] = 0
        wa[1] = 0
        shifts = [R.choice([0, 1, 5, N // 4 + 3, N - 1, 3 * N + 1]) for _ in range(n_cols)]
        out = eng.combination(dev, wa, wb, shifts, N, 7, w)
        assert np.array_equal(eng.download(out), orc.combination(host, wa, wb, shifts, 7, w)), logn
    out = eng.combination([], np.zeros((0, 3)), np.zeros((0, 3)), [], 16, 7, root_of_unity(4))
    assert not eng.download(out).any()
    from stark_brainfuck_b200._lib import B2SError
    with pytest.raises(B2SError):
        eng.combination([eng.upload(rand_bfe(1, 12))], np.zeros((1, 3)), np.zeros((1, 3)), [0], 12, 7, 1)


def test_merkle_upper_rebuilds_inner_nodes(eng):
    """b2s_merkle_upper (the replicated top levels of a multi-GPU tree): inner nodes from the leaf-level digests
    alone, against the tree b2s_merkle_field built and against the oracle"""
    import torch
    tpl = tpl_pair()[0]
    for logn in (0, 1, 4, 9, 13):
        n = 1 << logn
        nodes = eng.merkle_field(eng.upload(rand_xfe(40 + logn, n)), tpl)
        fresh = torch.zeros_like(nodes)
        fresh[n:] = nodes[n:]
        eng.merkle_upper(fresh)
        assert torch.equal(fresh[1:], nodes[1:])
        host = np.zeros((2 * n, 64), dtype=np.uint8)
        host[n:] = np.frombuffer(eng.download_bytes(nodes[n:]), dtype=np.uint8).reshape(n, 64)
        assert orc.merkle_upper(host)[1:].tobytes() == eng.download_bytes(nodes[1:])
    from stark_brainfuck_b200._lib import B2SError
    with pytest.raises(B2SError):
        eng.merkle_upper(torch.zeros((6, 64), dtype=torch.uint8, device=eng.device))


def test_quotients_with_the_brainfuck_air_programs(eng):
    """the real constraint polynomials of the reference's five tables (tests/golden/air.json: up to 244 monomials,
    4 factors, exponents to 8 per table) on random codewords: device == oracle, all three zerofier kinds"""
    from util import quotient_program
    air = golden("air.json")
    # 2^10 points for every table; the processor table again at 2^16, where the kernel runs two points per thread
    for ti, t, logn in [(ti, t, 10) for ti, t in enumerate(air["tables"])] + [(0, air["tables"][0], 16)]:
        N = 1 << logn
        w = root_of_unity(logn)
        W = t["full_width"]
        for lifted in (False, True) if logn == 10 else (True,):
            cw = np.stack([rand_xfe(3000 + 17 * ti + j, N) for j in range(W)])
            if lifted:
                # what BrainfuckStark.prove() hands over: base columns lifted into the extension field (zero upper
                # planes) -- the kernel's base-field fast path; one extension column with a few base-field values
                cw[:t["base_width"], 1:, :] = 0
                cw[W - 1, 1:, ::7] = 0
            d = eng.upload(cw.reshape(3 * W, N)).reshape(W, 3, N)
            for kind, name in ((1, "boundary"), (2, "transition"), (3, "terminal")):
                prog = quotient_program(t[name])
                height = 64
                oinv = pow(root_of_unity(6), P - 2, P)
                out, vanishes = eng.quotients(d, N // height, *prog, kind, height, oinv, 7, w)
                ref, rv = orc.quotients(cw, N // height, *prog, kind, height, oinv, 7, w)
                assert not vanishes and not rv
                assert np.array_equal(eng.download(out.reshape(-1, N)).reshape(-1, 3, N), ref), (t["name"], name, lifted)
                if lifted:  # the caller's own flags instead of the library's scan (what the glue passes)
                    flags = [j < t["base_width"] for j in range(W)]
                    out, _ = eng.quotients(d, N // height, *prog, kind, height, oinv, 7, w, base_columns=flags)
                    assert np.array_equal(eng.download(out.reshape(-1, N)).reshape(-1, 3, N), ref), (t["name"], name, "flags")
                    # ... and without the zero flag's read-back (the glue rules a vanishing zerofier out on the host):
                    # the call must not need its own synchronisation to be correct
                    out, _ = eng.quotients(d, N // height, *prog, kind, height, oinv, 7, w, base_columns=flags,
                                           check_zerofier=False)
                    assert np.array_equal(eng.download(out.reshape(-1, N)).reshape(-1, 3, N), ref), (t["name"], name, "async")


def test_open_multi_matches_single_calls(eng):
    """b2s_open_multi (the whole query phase in one call) against b2s_gather / b2s_merkle_open per tree"""
    tpl = tpl_pair()[0]
    R = random.Random(5)
    trees = []
    for logn in (12, 7, 3, 1, 0):
        planes = eng.upload(rand_xfe(70 + logn, 1 << logn))
        trees.append((planes, eng.merkle_field(planes, tpl), [R.randrange(1 << logn) for _ in range(R.randrange(1, 9))]))
    sets = []
    for planes, nodes, idx in trees:
        sets += [(planes, None, idx), (None, nodes, idx), (planes, nodes, idx[:2])]
    sets.append((trees[0][0], None, []))
    got = eng.open_multi(sets)
    assert len(got) == len(sets)
    for (planes, nodes, idx), (vals, paths) in zip(sets, got):
        if planes is not None and idx:
            assert np.array_equal(vals, eng.gather(planes, idx))
        else:
            assert vals is None or len(vals) == 0
        if nodes is not None and idx:
            want = eng.merkle_open(nodes, idx)
            assert [[paths[q, j].tobytes() for j in range(paths.shape[1])] for q in range(len(idx))] == want
    from stark_brainfuck_b200._lib import B2SError
    with pytest.raises(B2SError):
        eng.open_multi([(None, trees[1][1], [1 << 7])])


def _planes(seed, n, ext):
    return rand_xfe(seed, n) if ext else rand_bfe(seed, n).reshape(1, n)


def _lift(a):
    out = np.zeros((3, a.shape[1]), dtype=np.uint64)
    out[:a.shape[0]] = a
    return out


@pytest.mark.parametrize("m", (65, 300, 4096))
@pytest.mark.parametrize("k", (1, 50, 1 << 10))
def test_eval_points_multi_chunk_vs_oracle(eng, m, k):
    """b2s_eval_points beyond one chunk of coefficients (m > 64), all four plane combinations
    (code/univariate.py:145-154; Glue._interpolated_planes depends on this branch for every table of height >= 128)"""
    for cx in (False, True):
        for px in (False, True):
            c, pts = _planes(900 + m, m, cx), _planes(901 + k, k, px)
            got = eng.download(eng.eval_points(eng.upload(c), eng.upload(pts)))
            if not cx and not px:
                want = orc.eval_points(c[0], pts[0]).reshape(1, k)
            else:
                want = orc.eval_points(_lift(c), _lift(pts))
            assert np.array_equal(got, want), (m, k, cx, px)


def test_eval_points_config3_sizes(eng):
    """BASELINE config 3: 2^18 extension-field coefficients.  At 50 arbitrary points against the oracle; at 2^10
    points of the coset 7 * omega^k against the coset transform (itself pinned to the oracle and the goldens)."""
    logm = 18
    m = 1 << logm
    c = rand_xfe(118, m)
    dc = eng.upload(c)
    pts = rand_xfe(119, 50)
    assert np.array_equal(eng.download(eng.eval_points(dc, eng.upload(pts))), orc.eval_points(c, pts))
    w = root_of_unity(logm)
    cos = np.array([7 * pow(w, i, P) % P for i in range(1 << 10)], dtype=np.uint64)
    full = eng.download(eng.ntt(dc, logm, w, offset=7))
    assert np.array_equal(eng.download(eng.eval_points(dc, eng.upload(cos))), full[:, :1 << 10])
    b = rand_bfe(120, m)
    bpts = rand_bfe(121, 1 << 10)
    assert np.array_equal(eng.download(eng.eval_points(eng.upload(b), eng.upload(bpts)))[0], orc.eval_points(b, bpts))


@pytest.mark.parametrize("logn", (12, 20))
def test_scale_big_vs_oracle(eng, logn):
    """b2s_scale at 2^12 and 2^20 coefficients, base-field and extension-field factor (code/univariate.py:168-169)"""
    n = 1 << logn
    c = rand_bfe(930 + logn, n)
    f = 0x123456789ABCDEF % P
    assert np.array_equal(eng.download(eng.scale(eng.upload(c), f))[0], orc.scale(f, c))
    xc = rand_xfe(931 + logn, n)
    xf = [int(v) for v in rand_xfe(932, 1)[:, 0]]
    assert np.array_equal(eng.download(eng.scale(eng.upload(xc), xf)), orc.xscale(xf, xc))


def test_row_leaves_vs_oracle(eng):
    """b2s_merkle_rows (zipped, salted rows; code/salted_merkle.py:25-35) against orc_row_leaves on a synthetic
    template: every integer width, trimmed shapes reported as exceptions, the row-list pass, the inner nodes."""
    import torch
    n = 1 << 11
    R = random.Random(77)
    edge = [0, 1, 255, 256, 65535, 65536, (1 << 31) - 1, 1 << 31, (1 << 32) - 1, 1 << 39, (1 << 63) - 1, 1 << 63, P - 1]
    planes = [np.array([R.choice(edge) if R.random() < 0.4 else R.randrange(P) for _ in range(n)], dtype=np.uint64)
              for _ in range(7)]
    planes[3][planes[3] == 0] = 1
    planes[3][5] = planes[3][900] = 0  # mode 1 violated: exceptions
    planes[6][:] = 0                   # mode 2 plane
    planes[6][77] = 9                  # ... violated
    modes = np.array([0, 0, 0, 1, 0, 0, 2], dtype=np.uint8)
    segs = [bytes(R.getrandbits(8) for _ in range(L)) for L in (140, 3, 0, 17, 250, 1, 33)]
    tpl = b"".join(segs)
    seg_off = np.cumsum([0] + [len(x) for x in segs]).astype(np.uint32)
    salts = np.array([[R.getrandbits(8) for _ in range(24)] for _ in range(n)], dtype=np.uint8)
    pre, suf = b"\x80\x04\x95\x1c\0\0\0\0\0\0\0C\x18", b"\x94."
    want = np.zeros((2 * n, 64), dtype=np.uint8)
    exc_want = orc.row_leaves(planes, modes, tpl, seg_off, n, want, salts, pre, suf)
    assert sorted(exc_want.tolist()) == [5, 77, 900]
    dev = [eng.upload(a)[0] for a in planes]
    d_salts = eng.upload_bytes(salts)
    nodes, exc = eng.merkle_rows(dev, modes, tpl, seg_off, n, d_salts, pre, suf)
    assert sorted(exc.tolist()) == [5, 77, 900]
    got = np.frombuffer(eng.download_bytes(nodes), dtype=np.uint8).reshape(2 * n, 64)
    ok = np.ones(n, dtype=bool)
    ok[[5, 77, 900]] = False
    assert np.array_equal(got[n:][ok], want[n:][ok])
    # second pass over the exception rows with a template of their shape (here: every plane emitted, no checks)
    modes2 = np.zeros(7, dtype=np.uint8)
    segs2 = segs + [b"xyz"]
    tpl2, seg2 = b"".join(segs2), np.cumsum([0] + [len(x) for x in segs2]).astype(np.uint32)
    rows = np.array([5, 77, 900], dtype=np.uint32)
    orc.row_leaves(planes, modes2, tpl2, seg2, n, want, salts, pre, suf, rows=rows)
    d_rows = eng.upload_bytes(rows.view(np.uint8)).view(torch.int32)
    nodes, exc = eng.merkle_rows(dev, modes2, tpl2, seg2, n, d_salts, pre, suf, rows=d_rows, nodes=nodes, build_upper=False)
    assert len(exc) == 0
    eng.merkle_upper(nodes)
    want = orc.merkle_upper(want)
    got = np.frombuffer(eng.download_bytes(nodes), dtype=np.uint8).reshape(2 * n, 64)
    assert np.array_equal(got[1:], want[1:])
    # rows whose preimages differ in length by far more than the kernel's ring allows a thread to run ahead of its
    # warp (40 integers of 2 bytes against 40 of 11): the byte-by-byte path at the end of the kernel
    n2 = 512
    wide = [np.array([(R.randrange(256) if (i // 3) % 2 else R.randrange(1 << 63, P)) for i in range(n2)], dtype=np.uint64)
            for _ in range(40)]
    segs3 = [bytes(R.getrandbits(8) for _ in range(R.randrange(0, 40))) for _ in range(41)]
    tpl3, seg3 = b"".join(segs3), np.cumsum([0] + [len(x) for x in segs3]).astype(np.uint32)
    modes3 = np.zeros(40, dtype=np.uint8)
    want3 = np.zeros((2 * n2, 64), dtype=np.uint8)
    assert len(orc.row_leaves(wide, modes3, tpl3, seg3, n2, want3, salts[:n2], pre, suf)) == 0
    want3 = orc.merkle_upper(want3)
    nodes3, exc3 = eng.merkle_rows([eng.upload(a)[0] for a in wide], modes3, tpl3, seg3, n2, eng.upload_bytes(salts[:n2]),
                                   pre, suf)
    assert len(exc3) == 0
    assert eng.download_bytes(nodes3)[64:] == want3[1:].tobytes()
    # unsalted rows, one leaf
    one = [a[:1].copy() for a in planes[:6]]
    w1 = np.zeros((2, 64), dtype=np.uint8)
    orc.row_leaves(one, modes2[:6], tpl, seg_off, 1, w1)
    n1, _ = eng.merkle_rows([eng.upload(a)[0] for a in one], modes2[:6], tpl, seg_off, 1)
    assert eng.download_bytes(n1)[64:] == w1[1:].tobytes()


def test_blob_leaf_digests_every_length(eng):
    """b2s_merkle_blobs leaf digests for every preimage length 0..600 against hashlib (one leaf per tree and many per
    tree).  Round 2 found a miscompiled conditional byte load in this kernel (profiles/microbench/blob_san_repro.cu)
    that only showed under compute-sanitizer: the tail of the last block is now read under ordinary branches."""
    R = random.Random(11)
    blobs = [bytes(R.getrandbits(8) for _ in range(L)) for L in range(0, 601)]
    n = len(blobs)
    npo2 = 1024
    nodes = np.frombuffer(eng.download_bytes(eng.merkle_blobs(blobs)), dtype=np.uint8).reshape(-1, 64)
    for i, b in enumerate(blobs):
        assert bytes(nodes[npo2 + i]) == hashlib.blake2b(b).digest(), len(b)
    for L in (0, 1, 127, 128, 129, 385, 427, 512, 513):
        one = np.frombuffer(eng.download_bytes(eng.merkle_blobs([blobs[L]])), dtype=np.uint8).reshape(-1, 64)
        assert bytes(one[1]) == hashlib.blake2b(blobs[L]).digest(), L
    assert n == 601
