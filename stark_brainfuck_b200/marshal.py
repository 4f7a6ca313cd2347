"""Python objects <-> u64 planes, and the pickle leaf templates.

The reference's wire format is pickle (code/merkle.py:30, code/ip.py:19), so objects the
engine hands back must be instances of the caller's own classes with the same attribute
order and the same shared `field` objects the reference's arithmetic would have produced
(SURVEY.md Appendix B5):
  * BaseFieldElement results of ntt/intt carry values[0].field (code/ntt.py:11, :23)
  * ExtensionFieldElement results carry the caller's xfield, their coefficients carry
    xfield.modulus.coefficients[0].field (code/extension_field.py:62-66, :93-97)
  * coefficient lists are trimmed like code/extension_field.py:6-9
Objects are built with cls.__new__ + __dict__ in the constructor's attribute order, which
pickles byte-identically to objects built by __init__.
"""
import contextlib
import ctypes as C
import gc
import pickle

import numpy as np

from ._lib import TPL_MAX_BYTES, LeafTemplates

P = 18446744069414584321


@contextlib.contextmanager
def bulk_allocation():
    """Building millions of small container objects with the cyclic collector enabled costs ~10x:
    every generation-0 overflow rescans the growing heap (measured: 10.1 -> 1.1 us per
    ExtensionFieldElement).  Nothing built here is cyclic, so the collector is paused."""
    was = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was:
            gc.enable()


class Binding:
    """The element classes (the reference's, or the standalone mirror's) the glue instantiates."""

    def __init__(self, BaseFieldElement, BaseField, Polynomial, ExtensionFieldElement, ExtensionField):
        self.BaseFieldElement = BaseFieldElement
        self.BaseField = BaseField
        self.Polynomial = Polynomial
        self.ExtensionFieldElement = ExtensionFieldElement
        self.ExtensionField = ExtensionField

    @classmethod
    def from_modules(cls, algebra, univariate, extension_field):
        return cls(algebra.BaseFieldElement, algebra.BaseField, univariate.Polynomial,
                   extension_field.ExtensionFieldElement, extension_field.ExtensionField)

    # ---- classification -----------------------------------------------------------
    def is_xfe(self, v):
        return type(v) is self.ExtensionFieldElement

    def is_bfe(self, v):
        return type(v) is self.BaseFieldElement

    def inner_field(self, xfield):
        """the BaseField object every coefficient produced by extension arithmetic carries"""
        return xfield.modulus.coefficients[0].field

    # ---- objects -> planes --------------------------------------------------------
    def bfe_to_np(self, values):
        return np.fromiter((v.value for v in values), dtype=np.uint64, count=len(values))

    def xfe_to_np(self, values):
        """(3, n) planes; trimmed coefficients are zero"""
        n = len(values)
        out = np.zeros((3, n), dtype=np.uint64)
        c0, c1, c2 = out[0], out[1], out[2]
        for i, x in enumerate(values):
            co = x.polynomial.coefficients
            k = len(co)
            if k > 0:
                c0[i] = co[0].value
                if k > 1:
                    c1[i] = co[1].value
                    if k > 2:
                        c2[i] = co[2].value
                        if k > 3:
                            raise ValueError("extension field element with more than 3 coefficients")
        return out

    def xfe_canonical(self, values, xfield=None):
        """True iff every element has the identity pattern the device leaf templates assume:
        .field is one shared xfield, every coefficient .field is its inner base field, the
        coefficient list is trimmed and has no aliased entries (SURVEY B5 rules 1-2)."""
        if not values:
            return True
        xf = values[0].field if xfield is None else xfield
        if type(xf) is not self.ExtensionField:
            return False
        bf = self.inner_field(xf)
        X = self.ExtensionFieldElement
        for x in values:
            if type(x) is not X or x.field is not xf:
                return False
            co = x.polynomial.coefficients
            k = len(co)
            if k > 3:
                return False
            for c in co:
                if c.field is not bf:
                    return False
            if k:
                if co[k - 1].value == 0:
                    return False
                if k > 1 and (co[0] is co[1] or (k > 2 and (co[0] is co[2] or co[1] is co[2]))):
                    return False
        return True

    # ---- planes -> objects --------------------------------------------------------
    def np_to_bfe(self, arr, field):
        B = self.BaseFieldElement
        new = B.__new__
        out = []
        with bulk_allocation():
            for v in arr.tolist():
                o = new(B)
                o.__dict__ = {"value": v, "field": field}
                out.append(o)
        return out

    def make_xfe(self, c0, c1, c2, xfield, bf=None):
        if bf is None:
            bf = self.inner_field(xfield)
        B, Pn, X = self.BaseFieldElement, self.Polynomial, self.ExtensionFieldElement
        vals = [c0, c1, c2]
        while vals and vals[-1] == 0:
            vals.pop()
        co = []
        for v in vals:
            o = B.__new__(B)
            o.__dict__ = {"value": v, "field": bf}
            co.append(o)
        p = Pn.__new__(Pn)
        p.__dict__ = {"coefficients": co}
        x = X.__new__(X)
        x.__dict__ = {"polynomial": p, "field": xfield}
        return x

    def np_to_xfe(self, planes, xfield):
        bf = self.inner_field(xfield)
        mk = self.make_xfe
        B, Pn, X = self.BaseFieldElement, self.Polynomial, self.ExtensionFieldElement
        nB, nP, nX = B.__new__, Pn.__new__, X.__new__
        out = []
        app = out.append
        with bulk_allocation():
            for a, b, c in zip(planes[0].tolist(), planes[1].tolist(), planes[2].tolist()):
                if c:  # the common case spelled out: three coefficients, nothing to trim
                    o0 = nB(B)
                    o0.__dict__ = {"value": a, "field": bf}
                    o1 = nB(B)
                    o1.__dict__ = {"value": b, "field": bf}
                    o2 = nB(B)
                    o2.__dict__ = {"value": c, "field": bf}
                    p = nP(Pn)
                    p.__dict__ = {"coefficients": [o0, o1, o2]}
                    x = nX(X)
                    x.__dict__ = {"polynomial": p, "field": xfield}
                    app(x)
                else:
                    app(mk(a, b, c, xfield, bf))
        return out

    # ---- leaf templates -----------------------------------------------------------
    def xfe_templates(self, xfield):
        """pickle marker elements with 0..3 coefficients and split at the markers"""
        pk = [pickle.dumps(self.make_xfe(*([0xA1, 0xA2, 0xA3][:k] + [0] * (3 - k)), xfield)) for k in range(4)]
        tpl = templates_from_marker_pickles(pk, 3, True)
        tpl._pickles = pk
        return tpl

    def bfe_templates(self, field):
        B = self.BaseFieldElement
        o = B.__new__(B)
        o.__dict__ = {"value": 0xA1, "field": field}
        return templates_from_marker_pickles([pickle.dumps(o)], 1, False)


def templates_from_marker_pickles(pickles, n_slots, trim):
    """Build struct b2s_leaf_templates from pickles of marker elements whose coefficients
    are 0xA1, 0xA2, 0xA3 (pickled as BININT1 `K\\xA1` ...).  pickles[k] has k coefficients
    (trim) or exactly n_slots (no trim, single pickle)."""
    t = LeafTemplates()
    t.n_slots, t.trim = n_slots, 1 if trim else 0
    blob = bytearray()
    ks = list(range(n_slots + 1)) if trim else [n_slots]
    if len(pickles) != len(ks):
        raise ValueError("expected %d marker pickles" % len(ks))
    for k, pk in zip(ks, pickles):
        pk = bytes(pk)
        if pk[:3] != b"\x80\x04\x95" or int.from_bytes(pk[3:11], "little") != len(pk) - 11:
            raise ValueError("marker pickle is not a single protocol-4 frame")
        rest = pk[11:]
        segs = []
        for j in range(k):
            mark = bytes([0x4B, 0xA1 + j])
            if rest.count(mark) != 1:
                raise ValueError("ambiguous marker in leaf template")
            head, rest = rest.split(mark)
            segs.append(head)
        segs.append(rest)
        for j, s in enumerate(segs):
            t.seg_off[k][j] = len(blob)
            blob += s
            t.seg_off[k][j + 1] = len(blob)
    if len(blob) > TPL_MAX_BYTES:
        raise ValueError("leaf templates need %d bytes (max %d)" % (len(blob), TPL_MAX_BYTES))
    C.memmove(t.bytes, bytes(blob), len(blob))
    return t


# ---- row templates: pickle.dumps(tuple of field elements) with the integers cut out ------------------------
def pickle_uint(v):
    """CPython _pickle.c save_long, protocol 4, for 0 <= v < 2^64"""
    if v < 256:
        return b"K" + bytes([v])
    if v < 65536:
        return b"M" + v.to_bytes(2, "little")
    if v < 1 << 31:
        return b"J" + v.to_bytes(4, "little")
    nb = (v.bit_length() >> 3) + 1
    return b"\x8a" + bytes([nb]) + v.to_bytes(nb, "little")


def _row_marker(j):
    return 0xC3A5000000005AC3 | (j << 16)  # >= 2^63: pickled as LONG1 with 9 payload bytes


class RowTemplate:
    """Byte template of pickle.dumps(row) for the rows of a salted tree (code/salted_merkle.py:31): `tpl` cut into
    n_slots + 1 segments (`seg_off`) around the integers; `modes` has one entry per device plane of the row's
    columns in order (0 integer slot, 1 integer slot that must be non-zero, 2 trimmed coefficient that must be
    zero -- include/b2s.h b2s_merkle_rows)."""

    def __init__(self, signature, modes, tpl, seg_off):
        self.signature, self.modes, self.tpl, self.seg_off = signature, modes, tpl, seg_off

    def render(self, values):
        """the pickle for the given slot integers (host check of the template)"""
        body = b"".join(self.tpl[self.seg_off[j]:self.seg_off[j + 1]] + pickle_uint(v) for j, v in enumerate(values))
        body += self.tpl[self.seg_off[len(values)]:self.seg_off[len(values) + 1]]
        return b"\x80\x04\x95" + len(body).to_bytes(8, "little") + body


def row_signature(binding, row):
    """per element: -1 for a BaseFieldElement, the number of coefficients for an ExtensionFieldElement;
    None when the row holds anything else or has an object graph the template cannot express"""
    if type(row) is not tuple:
        return None
    sig, seen = [], set()
    for e in row:
        if binding.is_bfe(e):
            if list(e.__dict__) != ["value", "field"] or not isinstance(e.value, int) or not 0 <= e.value < 1 << 64:
                return None
            sig.append(-1)
            objs = [e]
        elif binding.is_xfe(e):
            if list(e.__dict__) != ["polynomial", "field"] or type(e.polynomial) is not binding.Polynomial or \
                    list(e.polynomial.__dict__) != ["coefficients"] or type(e.polynomial.coefficients) is not list:
                return None
            co = e.polynomial.coefficients
            if len(co) > 3 or (co and co[-1].value == 0):
                return None
            for c in co:
                if not binding.is_bfe(c) or list(c.__dict__) != ["value", "field"] or not isinstance(c.value, int) \
                        or not 0 <= c.value < 1 << 64:
                    return None
            sig.append(len(co))
            objs = [e, e.polynomial, co] + list(co)
        else:
            return None
        for o in objs:  # an object met twice would be pickled as a memo reference
            if id(o) in seen:
                return None
            seen.add(id(o))
    return tuple(sig)


def row_values(row):
    """the slot integers of a row, in pickle order"""
    vals = []
    for e in row:
        if "value" in e.__dict__:
            vals.append(e.value)
        else:
            vals += [c.value for c in e.polynomial.coefficients]
    return vals


def row_template(binding, row):
    """RowTemplate of `row`'s shape and identity pattern (same classes, same `field` objects), checked against
    pickle.dumps(row) itself; None if the row cannot be expressed."""
    sig = row_signature(binding, row)
    if sig is None:
        return None
    B, Pn, X = binding.BaseFieldElement, binding.Polynomial, binding.ExtensionFieldElement

    def bfe(value, field):
        o = B.__new__(B)
        o.__dict__ = {"value": value, "field": field}
        return o
    fake, modes, j = [], [], 0
    for e, k in zip(row, sig):
        if k < 0:
            fake.append(bfe(_row_marker(j), e.field))
            modes.append(0)
            j += 1
        else:
            co = [bfe(_row_marker(j + i), c.field) for i, c in enumerate(e.polynomial.coefficients)]
            j += k
            p = Pn.__new__(Pn)
            p.__dict__ = {"coefficients": co}
            x = X.__new__(X)
            x.__dict__ = {"polynomial": p, "field": e.field}
            fake.append(x)
            modes += [0] * (k - 1) + [1] * (k > 0) + [2] * (3 - k)
    pk = pickle.dumps(tuple(fake))
    if pk[:3] != b"\x80\x04\x95" or int.from_bytes(pk[3:11], "little") != len(pk) - 11:
        return None  # not a single protocol-4 frame (another interpreter default): the caller pickles on the host
    rest, blob, seg_off = pk[11:], bytearray(), [0]
    for i in range(j):
        mark = pickle_uint(_row_marker(i))
        if rest.count(mark) != 1:
            return None
        head, rest = rest.split(mark)
        blob += head
        seg_off.append(len(blob))
    blob += rest
    seg_off.append(len(blob))
    t = RowTemplate(sig, np.array(modes, dtype=np.uint8), bytes(blob), np.array(seg_off, dtype=np.uint32))
    if t.render(row_values(row)) != pickle.dumps(row):
        return None
    return t


def salt_frame(salt):
    """(prefix, suffix) of pickle.dumps(salt) around the salt bytes"""
    pk = pickle.dumps(salt)
    i = pk.find(salt)
    if type(salt) is not bytes or i < 0 or pk.count(salt) != 1:
        return None
    return pk[:i], pk[i + len(salt):]
