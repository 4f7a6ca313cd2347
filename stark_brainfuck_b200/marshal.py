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
