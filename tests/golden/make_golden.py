#!/usr/bin/env python3
"""Generate golden vectors for the hot path by running the UNMODIFIED reference.

Runs only in the authoring container (needs /root/reference/code, which never
travels to the GPU box).  Every fixture it writes under tests/golden/ is a small
JSON file; inputs are re-derivable anywhere from `random.Random(seed)` (stdlib),
outputs are stored as values (small cases) or sha256 digests (large cases).

Digest convention: sha256 over the concatenation of
little-endian u64 values; ExtensionFieldElement = c0,c1,c2 with trimmed
coefficients zero-filled, elements in list order.

Usage:  python tests/golden/make_golden.py <group> [...]
Groups: small  ntt_big  xntt_big  fri_small  fri_16  fri_18  fri_20  bfs  quotients  combination  lde  air  salted
Heavy groups are meant to run in the background, one process each.
"""
import hashlib
import json
import os
import pickle
import random
import struct
import sys
import time

REF = os.environ.get("B2S_REFERENCE_DIR", "/root/reference/code")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

from algebra import BaseField, BaseFieldElement  # noqa: E402
from extension_field import ExtensionField, ExtensionFieldElement  # noqa: E402
from univariate import Polynomial  # noqa: E402
import ntt as ref_ntt  # noqa: E402
from merkle import Merkle  # noqa: E402
from ip import ProofStream  # noqa: E402
from fri import Fri  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
P = 18446744069414584321
field = BaseField.main()
xfield = ExtensionField.main()
XBF = xfield.modulus.coefficients[0].field  # canonical inner base field (SURVEY B5 rule 1)


def dump(name, obj):
    path = os.path.join(HERE, name)
    with open(path, "w") as f:
        json.dump(obj, f, indent=1, sort_keys=True)
        f.write("\n")
    print("wrote", path, flush=True)


def bfe_digest(values):
    return hashlib.sha256(b"".join(struct.pack("<Q", v.value) for v in values)).hexdigest()


def xfe_triple(x):
    c = [co.value for co in x.polynomial.coefficients]
    return c + [0] * (3 - len(c))


def xfe_digest(values):
    h = hashlib.sha256()
    for x in values:
        h.update(struct.pack("<3Q", *xfe_triple(x)))
    return h.hexdigest()


def X(*coeffs):
    """canonical-identity ExtensionFieldElement (SURVEY B5 rule 5)"""
    return ExtensionFieldElement(Polynomial([BaseFieldElement(c, XBF) for c in coeffs]), xfield)


def rand_bfe_list(seed, n):
    R = random.Random(seed)
    return [BaseFieldElement(R.randrange(P), field) for _ in range(n)]


def rand_xfe_list(seed, n):
    R = random.Random(seed)
    return [X(R.randrange(P), R.randrange(P), R.randrange(P)) for _ in range(n)]


# --------------------------------------------------------------------------
def plain_ntt(vals, omega):
    """iterative plain-int NTT used ONLY to build big FRI inputs (B5 rule 5);
    validated below against the reference at small sizes before use."""
    n = len(vals)
    a = list(vals)
    logn = n.bit_length() - 1
    # bit reversal
    j = 0
    for i in range(1, n):
        bit = n >> 1
        while j & bit:
            j ^= bit
            bit >>= 1
        j |= bit
        if i < j:
            a[i], a[j] = a[j], a[i]
    length = 2
    while length <= n:
        w_len = pow(omega, n // length, P)
        half = length // 2
        tw = [1] * half
        for i in range(1, half):
            tw[i] = tw[i - 1] * w_len % P
        for start in range(0, n, length):
            for i in range(half):
                u = a[start + i]
                v = a[start + i + half] * tw[i] % P
                a[start + i] = (u + v) % P
                a[start + i + half] = (u - v) % P
        length <<= 1
    return a


def fri_input(logn, expansion, seed):
    """random degree < n/expansion XFE polynomial evaluated on the coset 7*omega^k
    with the plain-int NTT, wrapped in canonical-identity objects."""
    n = 1 << logn
    R = random.Random(seed)
    m = n // expansion
    coeffs = [(R.randrange(P), R.randrange(P), R.randrange(P)) for _ in range(m)]
    omega = field.primitive_nth_root(n).value
    planes = []
    for pl in range(3):
        g = 1
        col = []
        for i in range(m):
            col.append(coeffs[i][pl] * g % P)
            g = g * 7 % P
        col += [0] * (n - m)
        planes.append(plain_ntt(col, omega))
    cw = []
    for i in range(n):
        t = [planes[0][i], planes[1][i], planes[2][i]]
        while t and t[-1] == 0:
            t.pop()
        cw.append(X(*t))
    return coeffs, cw


def run_fri(logn, expansion, s, seed, check_xevaluate=False):
    n = 1 << logn
    omega = field.primitive_nth_root(n)
    fri = Fri(field.generator(), omega, n, expansion, s, xfield)
    coeffs, cw = fri_input(logn, expansion, seed)
    if check_xevaluate:
        poly = Polynomial([X(*c) for c in coeffs])
        ref_cw = fri.domain.xevaluate(poly, xfield)
        assert xfe_digest(ref_cw) == xfe_digest(cw), "plain-int NTT disagrees with reference xevaluate"
        assert pickle.dumps(ref_cw) == pickle.dumps(cw), "identity graph differs from reference objects"
    ps = ProofStream()
    t0 = time.time()
    top = fri.prove(cw, ps)
    dt = time.time() - t0
    ser = ps.serialize()
    root0 = Merkle(cw).root().hex() if logn <= 14 else None
    out = {
        "log_n": logn, "expansion": expansion, "num_colinearity_tests": s, "seed": seed,
        "codeword_digest": xfe_digest(cw), "top_level_indices": top,
        "transcript_len": len(ser), "transcript_sha256": hashlib.sha256(ser).hexdigest(),
        "num_objects": len(ps.objects), "reference_seconds": round(dt, 3),
        "round_roots": [o.hex() for o in ps.objects if isinstance(o, bytes)],
        "root0": root0,
    }
    if logn <= 14:
        v = fri.verify(ProofStream().deserialize(ser), Merkle(cw).root())
        out["verify"] = bool(v)
    return out


# --------------------------------------------------------------------------
def group_small():
    g = {}
    # GV1
    w8 = field.primitive_nth_root(8)
    v = [BaseFieldElement(i, field) for i in range(1, 9)]
    g["gv1_ntt8"] = [x.value for x in ref_ntt.ntt(w8, v)]
    g["gv1_intt8"] = [x.value for x in ref_ntt.intt(w8, ref_ntt.ntt(w8, v))]
    # roots
    g["roots"] = {str(k): field.primitive_nth_root(1 << k).value for k in range(1, 25)}
    # BFE ntt/intt digests + full small outputs  (seed = log n)
    g["ntt_bfe"] = {}
    for logn in range(1, 15):
        n = 1 << logn
        w = field.primitive_nth_root(n)
        x = rand_bfe_list(logn, n)
        t0 = time.time()
        y = ref_ntt.ntt(w, x)
        dt = time.time() - t0
        z = ref_ntt.intt(w, x)
        e = {"ntt_digest": bfe_digest(y), "intt_digest": bfe_digest(z), "ntt_seconds": round(dt, 4)}
        if logn <= 4:
            e["ntt_values"] = [t.value for t in y]
            e["intt_values"] = [t.value for t in z]
        g["ntt_bfe"][str(logn)] = e
        print("ntt_bfe", logn, dt, flush=True)
    # XFE ntt digests (seed = 100 + log n)
    g["ntt_xfe"] = {}
    for logn in range(1, 13):
        n = 1 << logn
        w = xfield.lift(field.primitive_nth_root(n))
        x = rand_xfe_list(100 + logn, n)
        t0 = time.time()
        y = ref_ntt.ntt(w, x)
        dt = time.time() - t0
        e = {"ntt_digest": xfe_digest(y), "ntt_seconds": round(dt, 4)}
        if logn <= 10:
            e["intt_digest"] = xfe_digest(ref_ntt.intt(w, x))
        g["ntt_xfe"][str(logn)] = e
        print("ntt_xfe", logn, dt, flush=True)
    # coset evaluate / interpolate via Fri.Domain  (seed 300+, 400+)
    g["coset"] = {}
    for logn in (3, 6, 9, 11):
        n = 1 << logn
        w = field.primitive_nth_root(n)
        dom = Fri.Domain(field.generator(), w, n)
        for m in sorted({n // 4, n // 4 + 1, n}):
            poly = Polynomial(rand_bfe_list(300 + logn, m))
            ev = dom.evaluate(Polynomial(list(poly.coefficients)))
            xpoly = Polynomial(rand_xfe_list(400 + logn, m))
            xev = dom.xevaluate(xpoly, xfield)
            e = {"evaluate_digest": bfe_digest(ev), "xevaluate_digest": xfe_digest(xev)}
            if m == n:
                ip = dom.interpolate(rand_bfe_list(500 + logn, n))
                e["interpolate_digest"] = bfe_digest(ip.coefficients)
                e["interpolate_len"] = len(ip.coefficients)
                xip = dom.xinterpolate(rand_xfe_list(600 + logn, n))
                e["xinterpolate_digest"] = xfe_digest(xip.coefficients)
            g["coset"]["%d_%d" % (logn, m)] = e
        print("coset", logn, flush=True)
    # coset with a non-generator offset (test_ntt.py::test_coset_evaluate uses offset 2)
    n = 512
    w = field.primitive_nth_root(n)
    poly = Polynomial(rand_bfe_list(777, 300))
    two = BaseFieldElement(2, field)
    g["coset_offset2"] = {"digest": bfe_digest(ref_ntt.fast_coset_evaluate(poly, two, w, n))}
    # Polynomial.scale and evaluate_domain  (seed 700+)
    R = random.Random(700)
    poly = Polynomial(rand_bfe_list(701, 37))
    fac = BaseFieldElement(R.randrange(P), field)
    g["scale_bfe"] = {"factor": fac.value, "digest": bfe_digest(poly.scale(fac).coefficients)}
    xpoly = Polynomial(rand_xfe_list(702, 29))
    xfac = X(R.randrange(P), R.randrange(P), R.randrange(P))
    g["scale_xfe"] = {"factor": xfe_triple(xfac), "digest": xfe_digest(xpoly.scale(xfac).coefficients)}
    pts = rand_bfe_list(703, 50)
    g["evaluate_domain_bfe"] = {"digest": bfe_digest(poly.evaluate_domain(pts))}
    xpts = rand_xfe_list(704, 41)
    g["evaluate_domain_xfe"] = {"digest": xfe_digest(xpoly.evaluate_domain(xpts))}
    # structured: XFE poly on a base-field coset lifted into the extension field
    w64 = xfield.lift(field.primitive_nth_root(64))
    off = xfield.lift(field.generator())
    cos = [off * (w64 ^ i) for i in range(64)]
    xp64 = Polynomial(rand_xfe_list(705, 64))
    g["evaluate_domain_xfe_coset64"] = {"digest": xfe_digest(xp64.evaluate_domain(cos))}
    # XFE arithmetic known answers (seed 800)
    R = random.Random(800)
    cases = []
    for _ in range(64):
        a = X(R.randrange(P), R.randrange(P), R.randrange(P))
        b = X(R.randrange(P), R.randrange(P), R.randrange(P))
        cases.append({"a": xfe_triple(a), "b": xfe_triple(b), "mul": xfe_triple(a * b),
                      "inv": xfe_triple(a.inverse()), "div": xfe_triple(a / b),
                      "add": xfe_triple(a + b), "sub": xfe_triple(a - b)})
    # degenerate shapes
    for a, b in [(X(5), X(0, 0, 7)), (X(0, 3), X(0, 0, P - 1)), (X(P - 1, P - 1, P - 1), X(P - 1, P - 1, P - 1)),
                 (X(1), X(2)), (X(0, 1), X(0, 1))]:
        cases.append({"a": xfe_triple(a), "b": xfe_triple(b), "mul": xfe_triple(a * b),
                      "inv": xfe_triple(a.inverse()), "div": xfe_triple(a / b),
                      "add": xfe_triple(a + b), "sub": xfe_triple(a - b)})
    g["xfe_arith"] = cases
    # BFE arithmetic known answers (seed 801)
    R = random.Random(801)
    bc = []
    specials = [0, 1, 2, P - 1, P - 2, (1 << 32) - 1, 1 << 32, (1 << 32) + 1, (1 << 63), P >> 1, 0xFFFFFFFF00000000]
    pairs = [(a, b) for a in specials for b in specials] + [(R.randrange(P), R.randrange(P)) for _ in range(64)]
    for a, b in pairs:
        A, B = BaseFieldElement(a, field), BaseFieldElement(b, field)
        bc.append([a, b, (A + B).value, (A - B).value, (A * B).value, A.inverse().value, (-A).value])
    g["bfe_arith"] = bc
    # xfield.sample known answers
    R = random.Random(802)
    sm = []
    for _ in range(8):
        seed = bytes(R.getrandbits(8) for _ in range(32))
        sm.append({"seed": seed.hex(), "xfe": xfe_triple(xfield.sample(seed)), "bfe": field.sample(seed).value})
    g["sample"] = sm
    dump("small.json", g)

    # ---- pickle leaf templates and Merkle ---------------------------------
    m = {}
    marks = [0xA1, 0xA2, 0xA3]
    tpl = {}
    for k in range(4):
        tpl[str(k)] = pickle.dumps(X(*marks[:k])).hex()
    m["xfe_marker_pickles"] = tpl
    m["bfe_marker_pickle"] = pickle.dumps(BaseFieldElement(0xA1, field)).hex()
    m["xfe_inner_bfe_marker_pickle"] = pickle.dumps(BaseFieldElement(0xA1, XBF)).hex()
    # int encodings at boundaries
    ints = [0, 1, 255, 256, 65535, 65536, (1 << 31) - 1, 1 << 31, (1 << 32) - 1, 1 << 32, (1 << 39) - 1, 1 << 39,
            (1 << 40), (1 << 47) - 1, 1 << 47, (1 << 55) - 1, 1 << 55, (1 << 56), (1 << 63) - 1, 1 << 63, P - 1]
    m["int_pickles"] = {str(v): pickle.dumps(v)[2:-1].hex() for v in ints}
    # GV2
    leaves = [X(1, 2, 3), X(P - 1, 0, 5), X(7), X()]
    t = Merkle(leaves)
    m["gv2"] = {"leaves": [xfe_triple(x) for x in leaves],
                "preimage_lens": [len(pickle.dumps(x)) for x in leaves],
                "root": t.root().hex(), "open2": [b.hex() for b in t.open(2)],
                "nodes": [b.hex() for b in t.nodes]}
    # random XFE / BFE codeword trees (seed 900+log n); boundary-sized ints mixed in
    m["xfe_trees"] = {}
    m["bfe_trees"] = {}
    for logn in range(0, 11):
        n = 1 << logn
        R = random.Random(900 + logn)
        vals = []
        for i in range(n):
            c = []
            for _ in range(3):
                r = R.random()
                if r < 0.15:
                    c.append(0)
                elif r < 0.5:
                    c.append(R.choice(ints))
                else:
                    c.append(R.randrange(P))
            vals.append(c)
        lv = []
        for c in vals:
            c = list(c)
            while c and c[-1] == 0:
                c.pop()
            lv.append(X(*c))
        t = Merkle(lv)
        m["xfe_trees"][str(logn)] = {"root": t.root().hex(), "values": vals if logn <= 5 else None,
                                     "nodes_sha256": hashlib.sha256(b"".join(t.nodes[1:])).hexdigest(),
                                     "open_last": [b.hex() for b in t.open(n - 1)]}
        bl = [BaseFieldElement(c[0], field) for c in vals]
        tb = Merkle(bl)
        m["bfe_trees"][str(logn)] = {"root": tb.root().hex(),
                                     "nodes_sha256": hashlib.sha256(b"".join(tb.nodes[1:])).hexdigest()}
    # uniform random XFE trees at the sizes GV-style (seed = 100+log n inputs)
    m["xfe_trees_uniform"] = {}
    for logn in (4, 8, 12, 14):
        lv = rand_xfe_list(100 + logn, 1 << logn)
        t = Merkle(lv)
        m["xfe_trees_uniform"][str(logn)] = {"root": t.root().hex()}
    # generic blobs (test_merkle.py style: lists of byte strings), non power of two counts
    m["blob_trees"] = {}
    for n in (1, 2, 3, 5, 8, 13, 64, 100):
        R = random.Random(1000 + n)
        lv = [[bytes(R.getrandbits(8) for _ in range(R.randrange(0, 256))),
               bytes(R.getrandbits(8) for _ in range(R.randrange(0, 256)))] for _ in range(n)]
        t = Merkle(lv)
        m["blob_trees"][str(n)] = {"root": t.root().hex(), "depth": t.depth,
                                   "open0": [b.hex() for b in t.open(0)],
                                   "open_last": [b.hex() for b in t.open(n - 1)],
                                   "nodes_sha256": hashlib.sha256(b"".join(t.nodes[1:])).hexdigest()}
    dump("merkle.json", m)

    # ---- FRI fold, sample_indices -----------------------------------------
    f = {}
    for logn in (4, 6, 8):
        n = 1 << logn
        R = random.Random(1100 + logn)
        cw = rand_xfe_list(1100 + logn, n)
        alpha = X(R.randrange(P), R.randrange(P), R.randrange(P))
        omega = xfield.lift(field.primitive_nth_root(n))
        offset = xfield.lift(field.generator())
        one = xfield.one()
        two = one + one
        nxt = [two.inverse() * ((one + alpha / (offset * (omega ^ i))) * cw[i] +
                                (one - alpha / (offset * (omega ^ i))) * cw[n // 2 + i]) for i in range(n // 2)]
        f["fold_%d" % logn] = {"alpha": xfe_triple(alpha), "digest": xfe_digest(nxt),
                               "values": [xfe_triple(x) for x in nxt] if logn == 4 else None}
    fr = Fri(field.generator(), field.primitive_nth_root(1024), 1024, 16, 17, xfield)
    R = random.Random(1200)
    seed = bytes(R.getrandbits(8) for _ in range(32))
    f["sample_indices"] = {"seed": seed.hex(), "size": 512, "reduced": 32, "number": 17,
                           "indices": fr.sample_indices(seed, 512, 32, 17)}
    dump("fold.json", f)


def group_fri_small():
    g = {"gv6": {}}
    # validate the plain-int path against reference xevaluate + identity at 2^8
    g["gv6"]["8"] = run_fri(8, 4, 8, 208, check_xevaluate=True)
    for logn in (4, 5, 6, 10, 12, 14):
        g["gv6"][str(logn)] = run_fri(logn, 4, min(8, 8), 200 + logn)
        print("fri", logn, g["gv6"][str(logn)]["reference_seconds"], flush=True)
    # expansion 32 / 40 checks (BASELINE config 4 as literally worded needs expansion >= 32)
    g["exp32_s40_12"] = run_fri(12, 32, 40, 1312)
    # test_fri.py configuration: degree 63, expansion 16, 17 checks, poly = [xfield(i)]
    n = 1024
    omega = field.primitive_nth_root(n)
    fri = Fri(field.generator(), omega, n, 16, 17, xfield)
    poly = Polynomial([xfield(i) for i in range(64)])
    cw = fri.domain.xevaluate(poly)
    ps = ProofStream()
    top = fri.prove(cw, ps)
    ser = ps.serialize()
    g["test_fri"] = {"codeword_digest": xfe_digest(cw), "top_level_indices": top, "transcript_len": len(ser),
                     "transcript_sha256": hashlib.sha256(ser).hexdigest(), "num_objects": len(ps.objects),
                     "root0": Merkle(cw).root().hex()}
    # GV3: 64 coeffs sampled from 30 random bytes, R = Random(5), n = 256
    R = random.Random(5)
    n = 256
    fri = Fri(field.generator(), field.primitive_nth_root(n), n, 4, 8, xfield)
    poly = Polynomial([xfield.sample(bytes(R.getrandbits(8) for _ in range(30))) for _ in range(64)])
    cw = fri.domain.xevaluate(poly, xfield)
    ps = ProofStream()
    top = fri.prove(cw, ps)
    ser = ps.serialize()
    g["gv3"] = {"codeword_digest": xfe_digest(cw), "top_level_indices": top, "transcript_len": len(ser),
                "transcript_sha256": hashlib.sha256(ser).hexdigest()}
    dump("fri_small.json", g)


def group_fri_big(logn):
    r = run_fri(logn, 4, 8, 200 + logn)
    dump("fri_%d.json" % logn, r)


def group_ntt_big():
    g = {}
    for logn in (16, 18, 20):
        n = 1 << logn
        w = field.primitive_nth_root(n)
        x = rand_bfe_list(logn, n)
        t0 = time.time()
        y = ref_ntt.ntt(w, x)
        dt = time.time() - t0
        g[str(logn)] = {"ntt_digest": bfe_digest(y), "ntt_seconds": round(dt, 2)}
        print("ntt_big", logn, dt, flush=True)
        if logn <= 18:
            t0 = time.time()
            z = ref_ntt.intt(w, x)
            g[str(logn)]["intt_digest"] = bfe_digest(z)
            g[str(logn)]["intt_seconds"] = round(time.time() - t0, 2)
        dump("ntt_big.json", g)


def group_xntt_big():
    g = {}
    for logn in (14, 16, 18):
        n = 1 << logn
        w = xfield.lift(field.primitive_nth_root(n))
        x = rand_xfe_list(100 + logn, n)
        t0 = time.time()
        y = ref_ntt.ntt(w, x)
        dt = time.time() - t0
        g[str(logn)] = {"ntt_digest": xfe_digest(y), "ntt_seconds": round(dt, 2)}
        print("xntt_big", logn, dt, flush=True)
        dump("xntt_big.json", g)


def group_bfs(source="++++", inputs=(), name="bfs.json"):
    """GV7: BrainfuckStark.prove('++++') with seeded urandom; `bfs_io`: a program with a
    loop, input and output symbols (all five tables non-trivial), FRI domain 2048."""
    R = random.Random(1234)
    fake = lambda n: bytes(R.getrandbits(8) for _ in range(n))  # noqa: E731
    os.urandom = fake
    import salted_merkle
    salted_merkle.urandom = fake
    from vm import VirtualMachine
    from brainfuck_stark import BrainfuckStark
    program = VirtualMachine.compile(source)
    running_time, input_symbols, output_symbols = VirtualMachine.run(program, input_data=list(inputs))
    processor_matrix, memory_matrix, instruction_matrix, input_matrix, output_matrix = VirtualMachine.simulate(
        program, input_data=input_symbols)
    bfs = BrainfuckStark(running_time, len(memory_matrix), program, input_symbols, output_symbols)
    t0 = time.time()
    proof = bfs.prove(program, processor_matrix, memory_matrix, instruction_matrix, input_matrix, output_matrix)
    dt = time.time() - t0
    ok = bfs.verify(proof)
    dump(name, {"program": source, "inputs": list(inputs), "urandom_seed": 1234, "proof_len": len(proof),
                      "proof_sha256": hashlib.sha256(proof).hexdigest(), "verify": bool(ok),
                      "fri_domain_length": bfs.fri.domain.length, "prove_seconds": round(dt, 1)})


def group_quotients():
    """Table.boundary/transition/terminal_quotients (code/table.py:155-286) and
    PermutationArgument.quotient (code/permutation_argument.py:11-20) of the unmodified reference on toy
    tables with seeded random constraint polynomials; inputs are stored in full (they are small), the
    constraint dictionaries as (exponent vector, coefficient) lists in dictionary order."""
    from multivariate import MPolynomial
    from permutation_argument import PermutationArgument
    from table import Table
    R = random.Random(31337)
    N, W = 64, 3
    dom = Fri.Domain(field.generator(), field.primitive_nth_root(N), N)

    def rx():
        return X(R.randrange(P), R.randrange(P), R.randrange(P))

    def mpoly(n_vars, n_mono):
        d = {}
        for _ in range(n_mono):
            k = [0] * n_vars
            for _ in range(R.randrange(0, 4)):
                k[R.randrange(n_vars)] += R.randrange(1, 4)
            d[tuple(k)] = rx() if R.random() < 0.8 else X(R.randrange(5))
        return MPolynomial(d)

    class Toy(Table):
        def __init__(self, length):
            super().__init__(xfield, 2, W, length, 1, field.primitive_nth_root(N), N)
            self.b = [mpoly(W, 3), mpoly(W, 1), MPolynomial(dict())]
            self.t = [mpoly(2 * W, 6), mpoly(2 * W, 2)]
            self.e = [mpoly(W, 4)]

        def boundary_constraints_ext(self, challenges):
            return self.b

        def transition_constraints_ext(self, challenges):
            return self.t

        def terminal_constraints_ext(self, challenges, terminals):
            return self.e

    def program(constraints):
        return [[[list(k), xfe_triple(v)] for k, v in c.dictionary.items()] for c in constraints]

    out = {"N": N, "width": W, "offset": dom.offset.value, "omega": dom.omega.value, "tables": []}
    tables = [Toy(5), Toy(0), Toy(16)]
    for t in tables:
        t.codewords = [[rx() for _ in range(N)] for _ in range(W)]
        t.codewords[1][3] = xfield.zero()
        t.codewords[2][7] = X(11)
        res = {"height": t.height, "omicron_inv": t.omicron.inverse().value, "unit_distance": t.unit_distance(N),
               "codewords": [[xfe_triple(x) for x in cw] for cw in t.codewords],
               "boundary": {"program": program(t.b),
                            "out": [[xfe_triple(x) for x in q] for q in t.boundary_quotients(dom, t.codewords, None)]},
               "transition": {"program": program(t.t),
                              "out": [[xfe_triple(x) for x in q] for q in t.transition_quotients(dom, t.codewords, None)]},
               "terminal": {"program": program(t.e),
                            "out": [[xfe_triple(x) for x in q]
                                    for q in t.terminal_quotients(dom, t.codewords, None, None)]}}
        out["tables"].append(res)
    pa = PermutationArgument(tables, (0, 1), (2, 2))
    out["permutation"] = {"lhs": [0, 1], "rhs": [2, 2], "out": [xfe_triple(x) for x in pa.quotient(dom)]}
    dump("quotients.json", out)


def group_air():
    """The constraint polynomials of the Brainfuck AIR as data: boundary / transition / terminal
    `*_constraints_ext` of the reference's five tables (code/processor_table.py:219-357, instruction_table.py,
    memory_table.py, io_table.py) for seeded challenges and terminals, flattened to (exponent vector,
    coefficient) lists in dictionary order.  Realistic programs for b2s_quotients: parity against the oracle and
    the device-time budget of a whole proof (profiles/microbench/prove_device_pipeline.py)."""
    from vm import VirtualMachine
    from brainfuck_stark import BrainfuckStark
    program = VirtualMachine.compile("++[>,.<-]")
    running_time, input_symbols, output_symbols = VirtualMachine.run(program, input_data=["a", "b"])
    _, memory_matrix, _, _, _ = VirtualMachine.simulate(program, input_data=input_symbols)
    bfs = BrainfuckStark(running_time, len(memory_matrix), program, input_symbols, output_symbols)
    R = random.Random(2718)

    def rx():
        return X(R.randrange(P), R.randrange(P), R.randrange(P))
    challenges = [rx() for _ in range(11)]
    terminals = [rx() for _ in range(5)]

    def flat(constraints):
        return [[[list(k), xfe_triple(v)] for k, v in c.dictionary.items()] for c in constraints]
    out = {"challenges": [xfe_triple(c) for c in challenges], "terminals": [xfe_triple(t) for t in terminals],
           "tables": []}
    for t in bfs.tables:
        out["tables"].append({"name": type(t).__name__, "base_width": t.base_width, "full_width": t.full_width,
                              "boundary": flat(t.boundary_constraints_ext(challenges)),
                              "transition": flat(t.transition_constraints_ext(challenges)),
                              "terminal": flat(t.terminal_constraints_ext(challenges, terminals))})
    dump("air.json", out)


def group_lde():
    """Table.interpolate_columns / lde / ldex (code/table.py:112-148) of the unmodified
    reference on toy tables (all heights from 0, 0-3 randomizers, constant and zero columns), randomizers drawn
    from a seeded os.urandom.  Stored: the trace matrices and the resulting codewords."""
    import table as table_mod
    from table import Table
    N = 64
    dom = Fri.Domain(field.generator(), field.primitive_nth_root(N), N)
    out = {"N": N, "offset": dom.offset.value, "omega": dom.omega.value, "cases": []}
    R = random.Random(777)
    for length, nr, bw, fw in ((0, 1, 2, 3), (1, 1, 2, 4), (2, 0, 1, 2), (3, 1, 3, 5), (8, 2, 2, 4), (13, 3, 3, 6),
                               (16, 1, 2, 3), (5, 0, 2, 4), (1, 3, 2, 4), (2, 3, 1, 2)):
        t = Table(field, bw, fw, length, nr, field.primitive_nth_root(N), N)
        h = t.height
        base = [[BaseFieldElement(R.randrange(P), field) for _ in range(bw)] for _ in range(h)]
        for r in range(h):
            base[r][0] = BaseFieldElement(42, field) if bw > 1 else base[r][0]  # a constant column
        seed = R.randrange(1 << 30)
        U = random.Random(seed)
        table_mod.os.urandom = lambda n: bytes(U.getrandbits(8) for _ in range(n))
        t.matrix = [list(row) for row in base]
        base_cw = list(t.lde(dom))  # ldex extends this very list in place (code/table.py:147)
        ext = [[X(R.randrange(P), R.randrange(P), R.randrange(P)) for _ in range(fw - bw)] for _ in range(h)]
        for r in range(h):
            ext[r][0] = X(7, 8, 9)  # constant extension column: shared coefficient objects in the codeword
            if fw - bw > 1:
                ext[r][1] = xfield.zero() if r else X(3)
        t.field = xfield  # what Table.extend does (code/processor_table.py:419)
        t.matrix = [[xfield.lift(v) for v in base[r]] + ext[r] for r in range(h)]
        ext_cw = t.ldex(dom, xfield)
        out["cases"].append({
            "length": length, "height": h, "num_randomizers": nr, "base_width": bw, "full_width": fw,
            "omicron": t.omicron.value, "urandom_seed": seed,
            "base": [[v.value for v in row] for row in base], "ext": [[xfe_triple(v) for v in row] for row in ext],
            "base_codewords": [[v.value for v in cw] for cw in base_cw],
            "ext_codewords": [[xfe_triple(v) for v in cw] for cw in ext_cw],
            "base_pickle_sha256": hashlib.sha256(pickle.dumps(base_cw)).hexdigest(),
            "ext_pickle_sha256": hashlib.sha256(pickle.dumps(ext_cw)).hexdigest()})
    dump("lde.json", out)


def group_combination():
    """the nonlinear combination block of BrainfuckStark.prove
    (code/brainfuck_stark.py:241-298).  The block is inline, so the reference's OWN statements are cut
    out of the source of prove() at run time (between its two marker comments) and executed on toy
    inputs; nothing is restated here.  Inputs are stored in full."""
    import inspect
    import textwrap
    from functools import reduce
    import brainfuck_stark
    src = inspect.getsource(brainfuck_stark.BrainfuckStark.prove)
    a = src.index("# compute terms of nonlinear combination polynomial")
    b = src.index("# commit to combination codeword")
    a = src.rindex("\n", 0, a) + 1
    b = src.rindex("\n", 0, b) + 1
    block = textwrap.dedent(src[a:b])
    R = random.Random(4242)
    out = {"cases": []}
    for N, nb, ne, nq, max_degree in ((64, 3, 2, 3, 40), (32, 1, 0, 1, 9), (16, 0, 1, 0, 5)):
        dom = Fri.Domain(field.generator(), field.primitive_nth_root(N), N)

        def rx():
            return X(R.randrange(P), R.randrange(P), R.randrange(P))

        class Stub:
            pass
        me = Stub()
        me.xfield, me.max_degree, me.fri = xfield, max_degree, Stub()
        me.fri.domain = dom
        base = [[BaseFieldElement(R.randrange(P), field) for _ in range(N)] for _ in range(nb)]
        ext = [[rx() for _ in range(N)] for _ in range(ne)]
        quo = [[rx() for _ in range(N)] for _ in range(nq)]
        rnd = [rx() for _ in range(N)]
        if nb:
            base[0][1] = field.zero()
        if ne:
            ext[0][2] = xfield.zero()
            ext[-1][3] = X(5)
        if nq:
            quo[0][0] = X(0, 7)
        bdb = [R.randrange(0, max_degree + 1) for _ in range(nb)]
        edb = [R.randrange(0, max_degree + 1) for _ in range(ne)]
        qdb = [R.randrange(0, max_degree + 1) for _ in range(nq)]
        if nq:
            qdb[-1] = max_degree  # shift 0
        weights = [rx() for _ in range(1 + 2 * (nb + ne + nq))]
        if nb:
            weights[2] = xfield.zero()
        ns = {"self": me, "os": os, "reduce": reduce, "randomizer_codeword": rnd, "base_codewords": base,
              "num_base_polynomials": nb, "base_degree_bounds": bdb, "extension_codewords": ext,
              "num_extension_polynomials": ne, "extension_degree_bounds": edb, "quotient_codewords": quo,
              "num_quotient_polynomials": nq, "quotient_degree_bounds": qdb, "weights": weights}
        exec(compile(block, "brainfuck_stark.py:241-298", "exec"), ns)
        comb = ns["combination_codeword"]
        out["cases"].append({
            "N": N, "offset": dom.offset.value, "omega": dom.omega.value, "max_degree": max_degree,
            "randomizer": [xfe_triple(x) for x in rnd], "base": [[v.value for v in c] for c in base],
            "extension": [[xfe_triple(x) for x in c] for c in ext], "quotient": [[xfe_triple(x) for x in c] for c in quo],
            "base_degree_bounds": bdb, "extension_degree_bounds": edb, "quotient_degree_bounds": qdb,
            "weights": [xfe_triple(x) for x in weights], "out": [xfe_triple(x) for x in comb],
            "out_pickle_sha256": hashlib.sha256(pickle.dumps(comb)).hexdigest(),
            "root": Merkle(comb).root().hex()})
    dump("combination.json", out)



def group_salted():
    """code/salted_merkle.py over zipped rows, the way BrainfuckStark.prove builds its base and extension trees
    (code/brainfuck_stark.py:178-180, :197-199): tuples of elements that carry DIFFERENT field objects, seeded
    salts.  Columns are stored as plain integers (extension-field columns as coefficient triples); the test
    rebuilds the objects with one field object per `field_id`."""
    import salted_merkle
    edge = [0, 1, 255, 256, 65535, 65536, (1 << 31) - 1, 1 << 31, (1 << 32) - 1, 1 << 32, (1 << 39) - 1, 1 << 39,
            (1 << 63) - 1, 1 << 63, P - 1, 12345678901234567]
    cases = []

    def run(name, n, columns, seed):
        fields = {}
        R = random.Random(seed)
        cols = []
        for c in columns:
            if c["kind"] == "b":
                f = fields.setdefault(c["field_id"], BaseField.main())
                cols.append([BaseFieldElement(v, f) for v in c["values"]])
            else:
                cols.append([X(*t) for t in c["values"]])
        rows = list(zip(*cols))
        salted_merkle.urandom = lambda k: bytes(R.getrandbits(8) for _ in range(k))
        tree = salted_merkle.SaltedMerkle(rows)
        salted_merkle.urandom = os.urandom
        salt, path = tree.open(n // 3)
        assert salted_merkle.SaltedMerkle.verify(tree.root(), n // 3, salt, path, rows[n // 3])
        cases.append({"name": name, "n": n, "columns": columns, "salt_seed": seed,
                      "salts": [s_.hex() for _, s_ in tree.leafs], "root": tree.root().hex(),
                      "nodes_sha256": hashlib.sha256(b"".join(tree.nodes[1:])).hexdigest(),
                      "leaf_digests": [tree.nodes[n + i].hex() for i in range(n)],
                      "preimage_last_row": (pickle.dumps(rows[-1]) + pickle.dumps(tree.leafs[-1][1])).hex(),
                      "open_index": n // 3, "open_path": [d.hex() for d in path]})

    R = random.Random(4242)
    rb = lambda n: [R.randrange(P) for _ in range(n)]  # noqa: E731
    rx = lambda n: [[R.randrange(P) for _ in range(3)] for _ in range(n)]  # noqa: E731
    # the base tree's shape: randomizer (extension field) first, then base columns of several tables
    run("base_rows", 16, [{"kind": "x", "values": rx(16)}, {"kind": "b", "field_id": 0, "values": rb(16)},
                          {"kind": "b", "field_id": 0, "values": edge}, {"kind": "b", "field_id": 1, "values": rb(16)},
                          {"kind": "b", "field_id": 2, "values": [i % 3 for i in range(16)]}], 77)
    # the extension tree's shape: extension-field columns only
    run("extension_rows", 8, [{"kind": "x", "values": rx(8)} for _ in range(4)], 78)
    # rows of different pickle shapes: trimmed coefficients (code/extension_field.py:6-9), zero elements
    tr = rx(8)
    tr[2] = [5, 0, 0]
    tr[3] = [0, 0, 0]
    tr[5] = [7, 9, 0]
    tr2 = rx(8)
    tr2[0] = [0, 0, 0]
    tr2[5] = [1, 0, 0]
    run("trimmed_rows", 8, [{"kind": "x", "values": tr}, {"kind": "b", "field_id": 0, "values": rb(8)},
                            {"kind": "x", "values": tr2}], 79)
    # a first column that is constant one (one coefficient in every row), two rows only, one row only
    run("constant_first", 4, [{"kind": "x", "values": [[1, 0, 0]] * 4}, {"kind": "x", "values": rx(4)}], 80)
    run("two_rows", 2, [{"kind": "b", "field_id": 0, "values": rb(2)}, {"kind": "x", "values": rx(2)}], 81)
    run("one_row", 1, [{"kind": "b", "field_id": 0, "values": rb(1)}], 82)
    dump("salted.json", {"cases": cases})

if __name__ == "__main__":
    for grp in sys.argv[1:]:
        if grp == "small":
            group_small()
        elif grp == "fri_small":
            group_fri_small()
        elif grp.startswith("fri_"):
            group_fri_big(int(grp[4:]))
        elif grp == "ntt_big":
            group_ntt_big()
        elif grp == "xntt_big":
            group_xntt_big()
        elif grp == "bfs":
            group_bfs()
        elif grp == "bfs_io":
            group_bfs("++[>,.<-]", ("a", "b"), "bfs_io.json")
        elif grp == "bfs_nested":  # nested loops, FRI domain 8192
            group_bfs("+++[>+++[>+<-]<-]>>.", (), "bfs_nested.json")
        elif grp == "bfs_A":  # prints "A": 109 cycles, FRI domain 16384
            group_bfs("++++++++[>++++++++<-]>+.", (), "bfs_A.json")
        elif grp == "bfs_He":  # prints "He": 249 cycles, FRI domain 32768 (about 1.7 h of reference time)
            group_bfs("++++++++++[>+++++++>++++++++++<<-]>++.>+.", (), "bfs_He.json")
        elif grp == "bfs_echo":
            group_bfs("+++++[>,.<-]", tuple("hello"), "bfs_echo.json")
        elif grp == "bfs_cat":  # one input, one output symbol: both IO tables of height 1 (unit distance = the domain)
            group_bfs(",.", ("x",), "bfs_cat.json")
        elif grp == "bfs_two":  # two output symbols, then one input symbol
            group_bfs("++..,", ("q",), "bfs_two.json")
        elif grp == "air":
            group_air()
        elif grp == "lde":
            group_lde()
        elif grp == "combination":
            group_combination()
        elif grp == "quotients":
            group_quotients()
        elif grp == "salted":
            group_salted()
        else:
            raise SystemExit("unknown group " + grp)
