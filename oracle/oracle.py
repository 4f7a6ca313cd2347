"""ctypes/numpy front end of the CPU oracle (oracle/b2s_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(stark_brainfuck_b200/) never imports this module.

Vectors are numpy uint64 arrays: base field shape (n,), extension field shape (3, n)
(planes c0,c1,c2; trimmed coefficients zero-filled).
"""
import ctypes as C
import os
import subprocess

import numpy as np

P = 18446744069414584321
_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

TPL_MAX_BYTES = 2048


class LeafTemplates(C.Structure):
    _fields_ = [("n_slots", C.c_uint32), ("trim", C.c_uint32),
                ("seg_off", (C.c_uint32 * 5) * 4), ("bytes", C.c_uint8 * TPL_MAX_BYTES)]


def build(force=False):
    src = os.path.join(_HERE, "b2s_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u64, vp = C.c_uint64, C.c_void_p
        for name in ("orc_add", "orc_sub", "orc_mul", "orc_pow"):
            getattr(L, name).restype = u64
            getattr(L, name).argtypes = [u64, u64]
        for name in ("orc_neg", "orc_inv"):
            getattr(L, name).restype = u64
            getattr(L, name).argtypes = [u64]
        L.orc_sample.restype = u64
        L.orc_sample.argtypes = [C.c_char_p, C.c_uint32]
        L.orc_ntt.argtypes = [u64, vp, vp, u64]
        L.orc_intt.argtypes = [u64, vp, vp, u64]
        L.orc_scale.argtypes = [u64, vp, vp, u64]
        L.orc_scale.restype = None
        L.orc_coset_evaluate.argtypes = [u64, u64, vp, u64, vp, u64]
        L.orc_coset_interpolate.argtypes = [u64, u64, vp, vp, u64]
        for name in ("orc_xmul", "orc_xadd", "orc_xsub", "orc_xdiv"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [vp, vp, vp]
        L.orc_xinv.restype = None
        L.orc_xinv.argtypes = [vp, vp]
        L.orc_xsample.restype = None
        L.orc_xsample.argtypes = [C.c_char_p, C.c_uint32, vp]
        L.orc_xntt.argtypes = [u64, vp, u64, vp, u64, u64, C.c_int]
        L.orc_xscale.restype = None
        L.orc_xscale.argtypes = [vp, vp, u64, vp, u64, u64]
        L.orc_eval_points.restype = None
        L.orc_eval_points.argtypes = [vp, u64, vp, vp, u64]
        L.orc_xeval_points.restype = None
        L.orc_xeval_points.argtypes = [vp, u64, u64, vp, u64, vp, u64, u64]
        L.orc_fri_fold.restype = None
        L.orc_fri_fold.argtypes = [vp, u64, u64, vp, u64, u64, vp, u64]
        L.orc_quotients.argtypes = [vp, u64, C.c_uint32, u64, C.c_uint32, vp, vp, vp, C.c_uint32, C.c_uint32, u64, u64,
                                    u64, u64, vp]
        L.orc_quotients.restype = C.c_int
        L.orc_merkle_upper.argtypes = [vp, u64]
        L.orc_combination.argtypes = [vp, vp, vp, vp, vp, vp, C.c_uint32, u64, u64, u64, vp, u64]
        L.orc_combination.restype = None
        L.orc_merkle_upper.restype = None
        L.orc_blake2b.restype = None
        L.orc_blake2b.argtypes = [C.c_char_p, u64, vp]
        L.orc_pickle_uint.restype = C.c_uint32
        L.orc_pickle_uint.argtypes = [u64, vp]
        L.orc_leaf_preimage.restype = C.c_uint32
        L.orc_leaf_preimage.argtypes = [C.POINTER(LeafTemplates), vp, vp]
        L.orc_merkle_field.argtypes = [C.POINTER(LeafTemplates), vp, u64, u64, vp]
        L.orc_merkle_blobs.argtypes = [C.c_char_p, vp, u64, u64, vp]
        L.orc_merkle_open.restype = None
        L.orc_merkle_open.argtypes = [vp, u64, u64, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


_ERR = {-1: "cannot compute ntt of non-power-of-two sequence", -2: "primitive root must be nth root of unity",
        -3: "primitive root is not primitive nth root of unity", -4: "more coefficients than domain points"}


def _chk(rc):
    if rc:
        raise AssertionError(_ERR.get(rc, "oracle error %d" % rc))


# ---- scalar field ops ------------------------------------------------------
def add(a, b): return lib().orc_add(a, b)
def sub(a, b): return lib().orc_sub(a, b)
def mul(a, b): return lib().orc_mul(a, b)
def neg(a): return lib().orc_neg(a)
def inv(a): return lib().orc_inv(a)
def fpow(a, e): return lib().orc_pow(a, e)
def sample(b): return lib().orc_sample(bytes(b), len(b))


def _x3(t):
    return (C.c_uint64 * 3)(*[int(v) for v in t])


def _xop(name, a, b):
    r = (C.c_uint64 * 3)()
    getattr(lib(), name)(_x3(a), _x3(b), r)
    return [int(v) for v in r]


def xmul(a, b): return _xop("orc_xmul", a, b)
def xadd(a, b): return _xop("orc_xadd", a, b)
def xsub(a, b): return _xop("orc_xsub", a, b)
def xdiv(a, b): return _xop("orc_xdiv", a, b)


def xinv(a):
    r = (C.c_uint64 * 3)()
    lib().orc_xinv(_x3(a), r)
    return [int(v) for v in r]


def xsample(b):
    r = (C.c_uint64 * 3)()
    lib().orc_xsample(bytes(b), len(b), r)
    return [int(v) for v in r]


# ---- vectors ---------------------------------------------------------------
def ntt(omega, v):
    v = _u64(v)
    out = np.empty_like(v)
    _chk(lib().orc_ntt(omega, _p(v), _p(out), v.size))
    return out


def intt(omega, v):
    v = _u64(v)
    out = np.empty_like(v)
    _chk(lib().orc_intt(omega, _p(v), _p(out), v.size))
    return out


def xntt(omega, v, inverse=False):
    v = _u64(v)
    assert v.ndim == 2 and v.shape[0] == 3
    out = np.empty_like(v)
    n = v.shape[1]
    _chk(lib().orc_xntt(omega, _p(v), n, _p(out), n, n, 1 if inverse else 0))
    return out


def scale(factor, v):
    v = _u64(v)
    out = np.empty_like(v)
    lib().orc_scale(factor, _p(v), _p(out), v.size)
    return out


def xscale(factor, v):
    v = _u64(v)
    out = np.empty_like(v)
    n = v.shape[1]
    lib().orc_xscale(_x3(factor), _p(v), n, _p(out), n, n)
    return out


def coset_evaluate(offset, omega, coeffs, n):
    c = _u64(coeffs)
    if c.ndim == 2:
        return np.stack([coset_evaluate(offset, omega, c[i], n) for i in range(3)])
    out = np.empty(n, dtype=np.uint64)
    _chk(lib().orc_coset_evaluate(offset, omega, _p(c), c.size, _p(out), n))
    return out


def coset_interpolate(offset, omega, values):
    v = _u64(values)
    if v.ndim == 2:
        return np.stack([coset_interpolate(offset, omega, v[i]) for i in range(3)])
    out = np.empty_like(v)
    _chk(lib().orc_coset_interpolate(offset, omega, _p(v), _p(out), v.size))
    return out


def eval_points(coeffs, points):
    c, p = _u64(coeffs), _u64(points)
    if c.ndim == 1:
        out = np.empty_like(p)
        lib().orc_eval_points(_p(c), c.size, _p(p), _p(out), p.size)
        return out
    out = np.empty_like(p)
    lib().orc_xeval_points(_p(c), c.shape[1], c.shape[1], _p(p), p.shape[1], _p(out), p.shape[1], p.shape[1])
    return out


def fri_fold(cw, alpha, offset, omega):
    cw = _u64(cw)
    n = cw.shape[1]
    out = np.empty((3, n // 2), dtype=np.uint64)
    lib().orc_fri_fold(_p(cw), n, n, _x3(alpha), offset, omega, _p(out), n // 2)
    return out


def combination(columns, wa, wb, shifts, offset, omega):
    """nonlinear combination codeword; columns: list of (1, N) or (3, N) uint64 arrays -> (3, N)"""
    cols = [np.ascontiguousarray(c, dtype=np.uint64).reshape(-1, np.shape(c)[-1]) for c in columns]
    n, N = len(cols), cols[0].shape[1]
    ptrs = (C.c_void_p * n)(*[c.ctypes.data for c in cols])
    strides = np.full(n, N, dtype=np.uint64)
    planes = np.array([c.shape[0] for c in cols], dtype=np.uint32)
    wa = np.ascontiguousarray(wa, dtype=np.uint64).reshape(n, 3)
    wb = np.ascontiguousarray(wb, dtype=np.uint64).reshape(n, 3)
    shifts = np.ascontiguousarray(shifts, dtype=np.uint64).reshape(n)
    out = np.empty((3, N), dtype=np.uint64)
    lib().orc_combination(ptrs, _p(strides), _p(planes), _p(wa), _p(wb), _p(shifts), n, N, offset, omega, _p(out), N)
    return out


def merkle_upper(nodes):
    """(2 npo2, 64) uint8 with the digests of level npo2 filled in -> all inner nodes, in place"""
    nodes = np.ascontiguousarray(nodes, dtype=np.uint8)
    lib().orc_merkle_upper(_p(nodes), nodes.shape[0] // 2)
    return nodes


def quotients(cw, shift, mono_off, coeffs, factors, kind, height, omicron_inv, offset, omega):
    """code/table.py:155-286 / code/permutation_argument.py:11-20 for one table.
    cw: (width, 3, N) uint64; mono_off (C+1,) uint32; coeffs (M, 3) uint64; factors (M, F) uint32
    ((variable << 8) | exponent).  Returns ((C, 3, N) uint64, zerofier_vanishes flag)."""
    cw = np.ascontiguousarray(cw, dtype=np.uint64)
    width, _, n = cw.shape
    mono_off = np.ascontiguousarray(mono_off, dtype=np.uint32)
    coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 3)
    factors = np.ascontiguousarray(factors, dtype=np.uint32).reshape(len(coeffs), -1)
    nc = len(mono_off) - 1
    out = np.zeros((nc, 3, n), dtype=np.uint64)
    flag = lib().orc_quotients(_p(cw), n, width, shift, nc, mono_off.ctypes.data_as(C.c_void_p),
                               coeffs.ctypes.data_as(C.c_void_p), factors.ctypes.data_as(C.c_void_p),
                               factors.shape[1] if factors.size else 0, kind, height, omicron_inv, offset, omega, _p(out))
    return out, bool(flag)


# ---- hashing / pickle / Merkle ---------------------------------------------
def blake2b(msg):
    out = (C.c_uint8 * 64)()
    lib().orc_blake2b(bytes(msg), len(msg), out)
    return bytes(out)


def pickle_uint(v):
    out = (C.c_uint8 * 16)()
    n = lib().orc_pickle_uint(v, out)
    return bytes(out[:n])


def templates_from_marker_pickles(pickles, n_slots, trim):
    """pickles[k] = pickle.dumps(element with k coefficients 0xA1,0xA2,0xA3); for
    n_slots == 1 (BaseFieldElement) pass a single pickle.  Splits each body at the
    BININT1 markers `K\\xA1`, `K\\xA2`, `K\\xA3`."""
    t = LeafTemplates()
    t.n_slots, t.trim = n_slots, 1 if trim else 0
    blob = b""
    ks = range(n_slots + 1) if trim else [n_slots]
    for k, pk in zip(ks, pickles):
        pk = bytes(pk)
        assert pk[:3] == b"\x80\x04\x95" and int.from_bytes(pk[3:11], "little") == len(pk) - 11
        body = pk[11:]
        segs, rest = [], body
        for j in range(k):
            mark = bytes([0x4B, 0xA1 + j])
            assert rest.count(mark) == 1, "ambiguous marker"
            a, rest = rest.split(mark)
            segs.append(a)
        segs.append(rest)
        for j, s in enumerate(segs):
            t.seg_off[k][j] = len(blob)
            blob += s
            t.seg_off[k][j + 1] = len(blob)
    assert len(blob) <= TPL_MAX_BYTES
    C.memmove(t.bytes, blob, len(blob))
    return t


def leaf_preimage(tpl, coeffs):
    out = (C.c_uint8 * 640)()
    c = (C.c_uint64 * 3)(*([int(v) for v in coeffs] + [0] * (3 - len(coeffs))))
    n = lib().orc_leaf_preimage(C.byref(tpl), c, out)
    return bytes(out[:n])


def merkle_field(tpl, planes):
    """returns nodes as (2n, 64) uint8 array; planes shape (n,) or (3,n)"""
    v = _u64(planes)
    n = v.shape[-1]
    nodes = np.zeros((2 * n, 64), dtype=np.uint8)
    _chk(lib().orc_merkle_field(C.byref(tpl), _p(v), n, n, _p(nodes)))
    return nodes


def merkle_blobs(blobs):
    n = len(blobs)
    npo2 = 1
    while npo2 < n:
        npo2 *= 2
    offs = np.zeros(n + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(b) for b in blobs])
    data = b"".join(blobs)
    nodes = np.zeros((2 * npo2, 64), dtype=np.uint8)
    _chk(lib().orc_merkle_blobs(data, _p(offs), n, npo2, _p(nodes)))
    return nodes


def merkle_open(nodes, index):
    npo2 = nodes.shape[0] // 2
    depth = npo2.bit_length() - 1
    path = np.zeros((max(depth, 1), 64), dtype=np.uint8)
    lib().orc_merkle_open(_p(nodes), npo2, index, _p(path))
    return [bytes(path[j]) for j in range(depth)]


def row_leaves(planes, modes, tpl, seg_off, n, nodes, salts=None, salt_pre=b"", salt_suf=b"", rows=None):
    """digests of zipped-row leaves (code/salted_merkle.py:25-35) into slots [n, 2n) of `nodes`
    ((2n, 64) uint8, modified in place).  planes: list of (n,) uint64 arrays.  Returns the exception rows."""
    planes = [np.ascontiguousarray(a, dtype=np.uint64) for a in planes]
    ptrs = (C.c_void_p * len(planes))(*[a.ctypes.data for a in planes])
    modes = np.ascontiguousarray(modes, dtype=np.uint8)
    seg = np.ascontiguousarray(seg_off, dtype=np.uint32)
    n_slots = len(seg) - 2
    exc = np.zeros(n + 1, dtype=np.uint32)
    rws = None if rows is None else np.ascontiguousarray(rows, dtype=np.uint32)
    sl = None if salts is None else np.ascontiguousarray(salts, dtype=np.uint8)
    f = lib().orc_row_leaves
    f.restype = C.c_long
    got = f(ptrs, _p(modes), len(planes), C.c_uint64(n), bytes(tpl), _p(seg), n_slots,
            _p(sl) if sl is not None else None, sl.shape[1] if sl is not None else 0, bytes(salt_pre), len(salt_pre),
            bytes(salt_suf), len(salt_suf), _p(rws) if rws is not None else None,
            C.c_uint64(len(rws) if rws is not None else 0), _p(nodes), _p(exc))
    if got < 0:
        raise ValueError("orc_row_leaves: bad arguments")
    return exc[:got].copy()
