"""Object-level host glue: the reference's call surface implemented on the engine.

One `Glue` serves both front ends:
  * dropin.py patches these functions into the real reference modules
    (ntt.ntt/intt/fast_coset_*, Polynomial.scale/evaluate_domain, Fri.Domain.*,
    Fri.commit/query/query_last/prove, Merkle) so brainfuck_stark.prove() runs unmodified;
  * mirror/ (the standalone host-side mirror of that interface) calls them directly.

Argument meaning, return types, error behaviour (AssertionError / IndexError) and the
object identity graph that pickle observes follow the reference; see the per-function
citations.  All heavy arithmetic happens in libb2s.so through `Engine`.
"""
import contextlib
import pickle

import numpy as np
import torch

from .engine import P, default_engine


def _ilog2(n):
    return n.bit_length() - 1


class Glue:
    def __init__(self, binding, engine=None):
        self.B = binding
        self._engine = engine
        self._xtpl = {}  # id(xfield) -> (xfield, templates)
        self._btpl = {}  # id(field)  -> (field, templates)
        self._kept = None  # id(list) -> (list, device planes, probe positions, probe elements) inside keep_planes()
        self._lazy = False  # inside keep_planes(lazy=True): codewords are handed out as DeviceCodewords
        self.coset_routed = 0  # evaluate_domain calls that became one coset transform

    @property
    def engine(self):
        if self._engine is None:
            self._engine = default_engine()
        return self._engine

    # ------------------------------------------------------------------ helpers
    def base_value(self, e, what="element"):
        """int value of a BaseFieldElement or of a lifted ExtensionFieldElement (code/fri.py:37)"""
        if self.B.is_bfe(e):
            return e.value
        if self.B.is_xfe(e):
            co = e.polynomial.coefficients
            if len(co) == 0:
                return 0
            if len(co) == 1:
                return co[0].value
            return None
        raise TypeError("%s must be a field element, got %r" % (what, type(e)))

    def xfe_templates(self, xfield):
        ent = self._xtpl.get(id(xfield))
        if ent is None or ent[0] is not xfield:
            ent = (xfield, self.B.xfe_templates(xfield))
            self._xtpl[id(xfield)] = ent
        return ent[1]

    def bfe_templates(self, field):
        ent = self._btpl.get(id(field))
        if ent is None or ent[0] is not field:
            ent = (field, self.B.bfe_templates(field))
            self._btpl[id(field)] = ent
        return ent[1]

    def _to_device(self, values):
        """list of field elements -> (device planes, kind, field object for the results)"""
        if type(values) is DeviceCodeword and values.kind in "xb":
            return values._planes, values.kind, values._field
        first = values[0]
        if self.B.is_xfe(first):
            return self.engine.upload(self.B.xfe_to_np(values)), "x", first.field
        if self.B.is_bfe(first):
            return self.engine.upload(self.B.bfe_to_np(values)), "b", first.field
        raise TypeError("cannot transform a list of %r" % type(first))

    def _from_device(self, t, kind, field, keep=False, lazy=False):
        if lazy and self._lazy and self._kept is not None:
            return DeviceCodeword(self, t, field, kind)
        a = self.engine.download(t)
        values = self.B.np_to_xfe(a, field) if kind == "x" else self.B.np_to_bfe(a[0], field)
        if keep:  # a codeword that a later device op (nonlinear combination) can read without marshalling
            self.remember_planes(values, t)
        return values

    # Codeword lists must stay plain lists (pickle writes a subclass differently), so the link from a
    # returned list to the device planes it was read from is a side table.  It only exists inside
    # keep_planes() -- the drop-in opens it around BrainfuckStark.prove, whose code is known not to
    # mutate its codewords in place -- and holds the lists strongly, so ids cannot be recycled.
    @contextlib.contextmanager
    def keep_planes(self, lazy=False):
        """lazy=True (the drop-in's BrainfuckStark.prove): lde / ldex / Domain.xevaluate hand their codewords out as
        DeviceCodewords -- list-like views whose elements are built on first access.  A proof reads a few hundred
        rows of the 46 codewords it makes (the openings); building every element object is most of its host time."""
        outer, self._kept = self._kept, {}
        outer_lazy, self._lazy = self._lazy, bool(lazy)
        try:
            yield self
        finally:
            self._kept = outer
            self._lazy = outer_lazy

    def remember_planes(self, values, planes):
        if self._kept is not None and len(values):
            n = len(values)
            probe = sorted({(n * k) // 8 for k in range(8)} | {n - 1})  # positions whose identity is re-checked
            ent = (values, planes, probe, [values[i] for i in probe])
            self._kept[id(values)] = ent
            self._kept[("first", id(values[0]))] = ent  # rows zipped from codewords find their columns again
            for i in probe:  # ... and so do element-wise lifts of a base-field codeword (see lifted_planes_of)
                self._kept[("elem", id(values[i]))] = (ent, i)

    def planes_of(self, codeword):
        """device planes behind a codeword, or None"""
        if isinstance(codeword, DeviceCodeword):
            return codeword._planes
        ent = self._kept.get(id(codeword)) if self._kept is not None else None
        if ent is None or ent[0] is not codeword or len(codeword) != ent[1].shape[1]:
            return None
        # the list is the caller's: a replaced element (first, last or a few in between) drops the link
        if any(codeword[i] is not e for i, e in zip(ent[2], ent[3])):
            return None
        return ent[1]

    def lifted_planes_of(self, codeword):
        """The base-field plane behind `[xfield.lift(c) for c in base_codeword]` -- what every Table.extend of the
        reference does to its base codewords (e.g. code/io_table.py:106-107) -- or None.  lift wraps the base
        element ITSELF as the only coefficient (code/extension_field.py:113-116; a zero is trimmed away,
        :6-9), so the lifted list is recognised by identity at the probe positions of the kept base codeword."""
        if self._kept is None or type(codeword) is not list or not codeword:
            return None
        ent = None
        for i in sorted({0, len(codeword) - 1, len(codeword) // 2}):
            e = codeword[i]
            if not self.B.is_xfe(e):
                return None
            co = e.polynomial.coefficients
            if len(co) == 1:
                hit = self._kept.get(("elem", id(co[0])))
                if hit is not None and hit[1] == i:
                    ent = hit[0]
                    break
        if ent is None or len(ent[0]) != len(codeword) or ent[1].shape[0] != 1:
            return None
        for i, base in zip(ent[2], ent[3]):
            co = codeword[i].polynomial.coefficients if self.B.is_xfe(codeword[i]) else None
            if co is None or ent[0][i] is not base or not ((len(co) == 1 and co[0] is base) or
                                                             (len(co) == 0 and base.value == 0)):
                return None
        return ent[1]

    # ------------------------------------------------------------------ code/ntt.py
    def _transform(self, primitive_root, values, inverse, offset=1, n_out=None, lone_source=None, res_field=None,
                   lazy=False):
        n = len(values) if n_out is None else n_out
        w = self.base_value(primitive_root, "primitive_root")
        # every 2-power root of unity of F_p^3 lies in F_p, so a root with higher
        # coefficients can never pass the order asserts of code/ntt.py:13-16
        assert w is not None, "primitive root must be nth root of unity, where n is %d" % n
        first = values[0]
        if self.B.is_xfe(first):
            arr, kind = self.B.xfe_to_np(values), "x"
        elif self.B.is_bfe(first):
            arr, kind = self.B.bfe_to_np(values).reshape(1, -1), "b"
        else:
            raise TypeError("cannot transform a list of %r" % type(first))
        if res_field is None:
            res_field = first.field  # ntt keeps values[0].field (code/ntt.py:20-23)
        share = 0
        if kind == "x" and not inverse:
            share = self._shared_output_period(arr, n)
        out = self.engine.ntt(self.engine.upload(arr), _ilog2(n), w, offset=offset, inverse=inverse)
        if share and lazy and self._lazy and self._kept is not None:
            # the same identity graph as below, built element by element on first access
            if share == 1:
                src = lone_source if lone_source is not None else first
                assert pow(w, n, P) == 1, "primitive root must be nth root of unity, where n is %d" % n
                assert n < 2 or pow(w, n // 2, P) != 1, \
                    "primitive root is not primitive nth root of unity, where n is %d" % n
                return DeviceCodeword(self, out, src.field, "x", share=1, reps={0: src})
            return DeviceCodeword(self, out, res_field, "x", share=share)
        if share == 1:
            values = self._lone_coefficient_ntt(lone_source if lone_source is not None else first, w, n)
        elif share:
            # outputs i and i + share are the same sub-transform output with a zero odd partner all the
            # way up: distinct elements wrapping the SAME coefficient objects (see _lone_coefficient_ntt)
            base = self.B.np_to_xfe(self.engine.download(out)[:, :share], res_field)
            X = self.B.ExtensionFieldElement
            values = [X(base[i % share].polynomial, res_field) for i in range(n)]
        else:
            return self._from_device(out, kind, res_field, keep=True, lazy=lazy)
        self.remember_planes(values, out)  # the values are the device's either way; only the identities differ
        return values

    @staticmethod
    def _shared_output_period(arr, n):
        """Identity structure of the reference's recursive ntt on extension-field inputs
        (code/ntt.py:20-23): `evens[i] + w^i * odds[i]` re-wraps the left operand's coefficient
        objects whenever the odd branch is entirely zero (code/univariate.py:28-31,
        code/extension_field.py:6-9).  With D = the fewest trailing zero bits among the indices
        j >= 1 of non-zero inputs, the odd branches of the first D levels vanish, so outputs i and
        i + n/2^D share their coefficient objects.  Returns that period n/2^D, 1 when only input 0
        is non-zero (all outputs wrap ITS coefficient objects), 0 for the generic case (all fresh)
        and for the zero vector.  (Accidental cancellations inside a non-zero branch are not
        modelled: they need structured inputs of measure zero.)"""
        nz = np.nonzero(arr.any(axis=0))[0]
        if len(nz) == 0:
            return 0
        pos = nz[nz > 0]
        if len(pos) == 0:
            return 1
        low = int(np.bitwise_or.reduce(pos))  # its lowest set bit = 2^D
        period = n // (low & -low)
        return 0 if period == n else period

    def _lone_coefficient_ntt(self, source, w, n):
        """Forward transform of [x, 0, 0, ...] over the extension field.  The value is x everywhere;
        what matters is identity: every odd branch of the reference's recursion is zero, so all n
        results are distinct elements that SHARE x's coefficient objects -- visible in pickles
        (constant columns of brainfuck_stark.prove(): SURVEY App. B5).  Same asserts as the device
        path."""
        assert pow(w, n, P) == 1, "primitive root must be nth root of unity, where n is %d" % n
        assert n < 2 or pow(w, n // 2, P) != 1, "primitive root is not primitive nth root of unity, where n is %d" % n
        X = self.B.ExtensionFieldElement
        return [X(source.polynomial, source.field) for _ in range(n)]

    def _zeros_ntt(self, primitive_root, zero, n, lazy):
        """ntt(primitive_root, [zero] * n): the codeword of an EMPTY polynomial (the input / output tables of a program
        without input / output, code/table.py:117-118 -> code/fri.py:26-37).  The reference's recursion returns n
        distinct zero elements; inside prove() they are a device view over a zero plane (the transform still runs:
        the library checks the root's order like code/ntt.py:13-16)."""
        w = self.base_value(primitive_root, "primitive_root")
        B = self.B
        if not (lazy and self._lazy and self._kept is not None) or n < 2 or n & (n - 1) or w is None or \
                not ((B.is_xfe(zero) and zero.is_zero()) or (B.is_bfe(zero) and zero.value == 0)):
            return self.ntt(primitive_root, [zero] * n)
        kind = "x" if B.is_xfe(zero) else "b"
        out = self.engine.ntt(self.engine.zeros(3 if kind == "x" else 1, n), _ilog2(n), w)
        return DeviceCodeword(self, out, zero.field, kind)

    def ntt(self, primitive_root, values):
        """code/ntt.py:4-23"""
        assert len(values) & (len(values) - 1) == 0, "cannot compute ntt of non-power-of-two sequence"
        if len(values) <= 1:
            return values  # the same list object, code/ntt.py:8-9
        return self._transform(primitive_root, values, False)

    def intt(self, primitive_root, values):
        """code/ntt.py:26-42"""
        n = len(values)
        assert n & (n - 1) == 0, "cannot compute intt of non-power-of-two sequence"
        w = self.base_value(primitive_root, "primitive_root")
        assert w is not None and pow(w, n, P) == 1, "supplied root does not have supplied order"
        if n == 1:
            return values
        assert n >= 2 and pow(w, n // 2, P) != 1, "supplied root is not primitive root of supplied order"
        return self._transform(primitive_root, values, True)

    def fast_coset_evaluate(self, polynomial, offset, generator, order, lazy=False):
        """code/ntt.py:164-168: scale by offset, zero-pad to `order`, ntt -- one fused call"""
        coeffs = polynomial.coefficients
        m = len(coeffs)
        if m == 0:
            # nothing tells us the element class; the reference pads with offset.field.zero()
            zero = offset.field.zero()
            return self._zeros_ntt(generator, zero, order, lazy) if order > 1 else [zero] * order
        assert order & (order - 1) == 0, "cannot compute ntt of non-power-of-two sequence"
        assert m <= order, "polynomial has more coefficients than the domain has points"
        off = self.base_value(offset, "offset")
        if order <= 1:
            return list(coeffs)  # offset^0 * c_0
        if off is None:  # genuinely extension-field offset: scale kernel, then plain ntt
            return self.ntt(generator, self.poly_scale(polynomial, offset).coefficients +
                            [offset.field.zero()] * (order - m))
        lone = None
        if self.B.is_xfe(coeffs[0]) and self.B.is_xfe(offset) and not coeffs[0].is_zero() \
                and all(c.is_zero() for c in coeffs[1:]):
            lone = (offset ^ 0) * coeffs[0]  # the scaled constant term, built by the caller's own classes
        # the scaled coefficients (offset ^ i) * c carry offset.field (code/univariate.py:169), ntt keeps it
        return self._transform(generator, coeffs, False, offset=off, n_out=order, lone_source=lone,
                               res_field=offset.field, lazy=lazy)

    def fast_coset_interpolate(self, offset, generator, values):
        """code/ntt.py:171-174: intt, then scale by offset^-1; all n coefficients are kept"""
        n = len(values)
        assert n & (n - 1) == 0, "cannot compute intt of non-power-of-two sequence"
        w = self.base_value(generator, "generator")
        assert w is not None and pow(w, n, P) == 1, "supplied root does not have supplied order"
        Pn = self.B.Polynomial
        if n == 1:
            return Pn(values)  # offset^0
        assert n >= 2 and pow(w, n // 2, P) != 1, "supplied root is not primitive root of supplied order"
        off = self.base_value(offset, "offset")
        if off is None:
            return self.poly_scale(Pn(self.intt(generator, values)), offset.inverse())
        assert off != 0, "divide by zero"
        return Pn(self._transform(generator, values, True, offset=off))

    # ------------------------------------------------------------------ code/univariate.py
    def poly_scale(self, poly, factor):
        """code/univariate.py:168-169"""
        coeffs = poly.coefficients
        Pn = self.B.Polynomial
        if not coeffs:
            return Pn([])
        d, kind, field = self._to_device(coeffs)
        if self.B.is_xfe(factor):
            f = [c.value for c in factor.polynomial.coefficients]
            f += [0] * (3 - len(f))
            if kind == "b":
                # (factor^i) * c_i with an extension factor and base coefficients: the reference's
                # ExtensionField.multiply would fail on the BaseFieldElement operand
                raise AttributeError("'BaseFieldElement' object has no attribute 'polynomial'")
            out = self.engine.scale(d, f)
            res_field = factor.field  # left operand of `(factor ^ i) * c`
        else:
            if kind == "x":
                raise AttributeError("'ExtensionFieldElement' object has no attribute 'value'")
            out = self.engine.scale(d, factor.value)
            res_field = factor.field
        return Pn(self._from_device(out, kind, res_field))

    def poly_evaluate_domain(self, poly, domain):
        """code/univariate.py:153-154 on an arbitrary list of points"""
        if len(domain) == 0:
            return []
        coeffs = poly.coefficients
        first = domain[0]
        if not coeffs:
            return [d.field.zero() for d in domain]
        ck = "x" if self.B.is_xfe(coeffs[0]) else "b"
        pk = "x" if self.B.is_xfe(first) else "b"
        if ck == "x" and pk == "b" and self.B.is_bfe(first):
            raise AttributeError("'BaseFieldElement' object has no attribute 'polynomial'")
        if ck == "b" and pk == "x" and self.B.is_bfe(coeffs[0]):
            raise AttributeError("'ExtensionFieldElement' object has no attribute 'value'")
        coset = self._as_coset(domain) if len(coeffs) <= len(domain) else None
        if coset is not None:
            # BASELINE config 3's structured case: the points are offset * omega^k, k < n, so the values are ONE
            # coset transform (zero padding and the offset^j scale fused into the pass kernels) instead of
            # n running-power evaluations of code/univariate.py:145-150
            dc, ck, _ = self._to_device(coeffs)
            self.coset_routed += 1
            out = self.engine.ntt(dc, _ilog2(len(domain)), coset[1], offset=coset[0])
            return self._from_device(out, "x" if pk == "x" else "b", first.field)
        dc, ck, _ = self._to_device(coeffs)
        dp, pk, pfield = self._to_device(domain)
        out = self.engine.eval_points(dc, dp)
        # value = point.field.zero() + c * xi ...: results carry the point's field (code/univariate.py:146-150)
        return self._from_device(out, "x" if pk == "x" else "b", pfield if pk == "x" else first.field)

    def _as_coset(self, domain):
        """(offset, omega) if `domain` is [offset * omega^k for k in range(n)] with n >= 2 a power of two, omega a
        primitive n-th root of unity and every point in the base field (plain or lifted, all of one class and one
        field object); else None"""
        n = len(domain)
        if n < 2 or n & (n - 1):
            return None
        first = domain[0]
        if self.B.is_bfe(first):
            cls, f = self.B.BaseFieldElement, first.field
            if any(type(d) is not cls or d.field is not f for d in domain):
                return None
            v = [d.value for d in domain]
        elif self.B.is_xfe(first):
            cls, f = self.B.ExtensionFieldElement, first.field
            v = []
            for d in domain:
                if type(d) is not cls or d.field is not f or len(d.polynomial.coefficients) > 1:
                    return None
                co = d.polynomial.coefficients
                v.append(co[0].value if co else 0)
        else:
            return None
        off = v[0]
        if off == 0:
            return None
        w = v[1] * pow(off, P - 2, P) % P
        if pow(w, n, P) != 1 or pow(w, n // 2, P) == 1:
            return None
        x = off
        for k in range(n):
            if v[k] != x:
                return None
            x = x * w % P
        return off, w

    # ------------------------------------------------------------------ code/fri.py Domain
    def domain_evaluate(self, dom, polynomial):
        """code/fri.py:26-30"""
        coeffs = polynomial.coefficients
        n = dom.length
        if not coeffs:
            zero = dom.omega.field.zero()
            return self._zeros_ntt(dom.omega, zero, n, True)
        assert len(coeffs) <= n, "polynomial has more coefficients than the domain has points"
        if n <= 1:
            return self.B.Polynomial(coeffs).scale(dom.offset).coefficients
        w, off = dom.omega.value, dom.offset.value
        d, kind, _ = self._to_device(coeffs)
        if kind == "x":  # (offset ^ i) * c with a base-field offset, code/univariate.py:169
            raise AttributeError("'ExtensionFieldElement' object has no attribute 'value'")
        out = self.engine.ntt(d, _ilog2(n), w, offset=off)
        # scaled coefficients take offset.field (code/univariate.py:169); ntt keeps values[0].field
        return self._from_device(out, kind, dom.offset.field if kind == "b" else coeffs[0].field, keep=True, lazy=True)

    def domain_xevaluate(self, dom, polynomial, xfield=None):
        """code/fri.py:32-37"""
        if xfield is None:
            assert len(polynomial.coefficients) != 0, "trying to xevaluate zero polynomial with no target field"
            xfield = polynomial.coefficients[0].field
        return self.fast_coset_evaluate(polynomial, xfield.lift(dom.offset), xfield.lift(dom.omega), dom.length,
                                        lazy=True)

    def domain_interpolate(self, dom, values):
        """code/fri.py:39-40"""
        return self.fast_coset_interpolate(dom.offset, dom.omega, values)

    def domain_xinterpolate(self, dom, values):
        """code/fri.py:42-44"""
        xfield = values[0].field
        return self.fast_coset_interpolate(xfield.lift(dom.offset), xfield.lift(dom.omega), values)

    # ------------------------------------------------------------------ code/table.py:112-149 (next row 2)
    def _interpolated_planes(self, table, omega, omega_order, column_indices, urandom):
        """Coefficient planes of Table.interpolate_columns (code/table.py:112-136) for all requested columns
        at once.  The interpolant through the `height` trace points (on the subgroup <omicron>) and the
        `num_randomizers` extra points omega^(2i+1) is unique, so it is computed as
            f = f0 + (x^height - 1) * q,   f0 = intt_omicron(trace),   q through (rho_k, (r_k - f0(rho_k)) / (rho_k^height - 1))
        instead of the reference's recursive fast_interpolate.  Randomizers are drawn exactly like the
        reference draws them (per column, table.field.sample(urandom(24))), so seeded runs stay byte-identical.
        Returns (planes (ncols * pl, height + nr) on the device, pl, kind, first trace element)."""
        B, eng = self.B, self.engine
        # code/table.py:113-114
        assert omega.has_order_po2(omega_order), "omega does not have claimed order"
        h, nr = table.height, table.num_randomizers
        cols = list(column_indices)
        traces, rand_vals = [], []
        for c in cols:
            trace = [row[c] for row in table.matrix]
            randomizers = [table.field.sample(urandom(3 * 8)) for _ in range(nr)]
            # code/table.py:129-130
            assert len(trace) + nr == h + nr, f"length of domain {h + nr} and values {len(trace) + nr} are unequal"
            traces.append(trace)
            rand_vals.append(randomizers)
        first = traces[0][0]
        if B.is_xfe(first):
            kind, pl = "x", 3
            arr = np.concatenate([B.xfe_to_np(t) for t in traces])
            rv = [[[co.value for co in r.polynomial.coefficients] + [0] * (3 - len(r.polynomial.coefficients))
                   for r in rs] for rs in rand_vals]
        elif B.is_bfe(first):
            kind, pl = "b", 1
            arr = np.stack([B.bfe_to_np(t) for t in traces])
            rv = [[[r.value] for r in rs] for rs in rand_vals]
        else:
            raise TypeError("cannot interpolate a column of %r" % type(first))
        omicron = table.omicron.value
        buf = eng.zeros(len(cols) * pl, h + nr)
        f0 = buf[:, :h]
        if h > 1:
            eng.ntt(eng.upload(arr), _ilog2(h), omicron, inverse=True, out=f0)
        else:
            eng.upload_into(f0, arr)
        if nr:
            w = omega.value
            rho = [pow(w, 2 * k + 1, P) for k in range(nr)]
            zinv = [pow((pow(r, h, P) - 1) % P, P - 2, P) for r in rho]  # rho is an odd power: not on the subgroup
            pts = eng.upload(np.array(rho, dtype=np.uint64))
            # basis polynomials of the nr extra points (base field, tiny): L_k(x) = prod_{j != k} (x - rho_j) / (rho_k - rho_j)
            basis = []
            for k in range(nr):
                poly, den = [1], 1
                for j in range(nr):
                    if j != k:
                        poly = [(a - rho[j] * b) % P for a, b in zip([0] + poly, poly + [0])]
                        den = den * (rho[k] - rho[j]) % P
                dinv = pow(den, P - 2, P)
                basis.append([c * dinv % P for c in poly])
            # f = f0 - q + x^h * q touches coefficients [0, nr) and [h, h + nr) (overlapping when h < nr)
            pos = sorted(set(range(nr)) | set(range(h, h + nr)))
            where = {p_: i for i, p_ in enumerate(pos)}
            runs = []  # pos as contiguous column ranges [a, b)
            for p_ in pos:
                if runs and runs[-1][1] == p_:
                    runs[-1][1] = p_ + 1
                else:
                    runs.append([p_, p_ + 1])
            fix = np.concatenate([eng.download(buf[:, a:b]) for a, b in runs], axis=1)
            for ci in range(len(cols)):
                at_rho = eng.download(eng.eval_points(f0[ci * pl:(ci + 1) * pl], pts))  # (pl, nr)
                for s_ in range(pl):
                    t = [(rv[ci][k][s_] - int(at_rho[s_, k])) * zinv[k] % P for k in range(nr)]
                    q = [sum(t[k] * basis[k][i] for k in range(nr)) % P for i in range(nr)]
                    row = fix[ci * pl + s_]
                    for i in range(nr):
                        row[where[i]] = (int(row[where[i]]) - q[i]) % P
                    for i in range(nr):
                        row[where[h + i]] = (int(row[where[h + i]]) + q[i]) % P
            at = 0
            for a, b in runs:
                eng.upload_into(buf[:, a:b], fix[:, at:at + b - a])
                at += b - a
        return buf, pl, kind, first

    def table_interpolate_columns(self, table, omega, omega_order, column_indices, urandom):
        """code/table.py:112-136"""
        Pn = self.B.Polynomial
        if table.height == 0:
            assert omega.has_order_po2(omega_order), "omega does not have claimed order"
            return [Pn([])] * len(column_indices)
        if len(column_indices) == 0:
            assert omega.has_order_po2(omega_order), "omega does not have claimed order"
            return []
        buf, pl, kind, first = self._interpolated_planes(table, omega, omega_order, column_indices, urandom)
        a = self.engine.download(buf)
        if kind == "b":
            return [Pn(self.B.np_to_bfe(a[c], first.field)) for c in range(a.shape[0])]
        return [Pn(self.B.np_to_xfe(a[3 * c:3 * c + 3], first.field)) for c in range(a.shape[0] // 3)]

    def table_lde(self, table, domain, urandom, xfield=None, columns=None):
        """code/table.py:138-148: lde (xfield None: columns [0, base_width), Domain.evaluate) and ldex (columns
        [base_width, full_width), Domain.xevaluate) -- interpolation and coset evaluation of all columns of the
        table as two batched device transforms.  Returns the list of codewords (the caller stores them)."""
        B, eng = self.B, self.engine
        if columns is None:
            columns = range(table.base_width) if xfield is None else range(table.base_width, table.full_width)
        Pn = self.B.Polynomial
        if table.height == 0 or len(columns) == 0 or domain.length <= 1:
            polys = self.table_interpolate_columns(table, domain.omega, domain.length, columns, urandom)
            return [self.domain_evaluate(domain, p) if xfield is None else self.domain_xevaluate(domain, p, xfield)
                    for p in polys]
        buf, pl, kind, first = self._interpolated_planes(table, domain.omega, domain.length, columns, urandom)
        N, m = domain.length, buf.shape[1]
        assert m <= N, "polynomial has more coefficients than the domain has points"
        if xfield is None and kind == "x":  # (offset ^ i) * c with a base-field offset, code/univariate.py:169
            raise AttributeError("'ExtensionFieldElement' object has no attribute 'value'")
        if xfield is not None and kind == "b":
            raise AttributeError("'BaseFieldElement' object has no attribute 'polynomial'")
        coeffs = eng.download(buf) if kind == "x" else None
        out = eng.ntt(buf, _ilog2(N), domain.omega.value, offset=domain.offset.value)
        lazy = self._lazy and self._kept is not None
        a = None if lazy else eng.download(out)
        res = []
        for c in range(len(columns)):
            if kind == "b":
                if lazy:
                    values = DeviceCodeword(self, out[c:c + 1], domain.offset.field, "b")
                else:
                    values = B.np_to_bfe(a[c], domain.offset.field)  # code/fri.py:26-30 via code/univariate.py:169
                    self.remember_planes(values, out[c:c + 1])
            else:
                ca = coeffs[3 * c:3 * c + 3]
                if self._shared_output_period(ca, N):
                    # sparse interpolants (constant columns) share coefficient objects between outputs in the
                    # reference's recursion: take the per-column path that models it (see _transform)
                    values = self.domain_xevaluate(domain, Pn(B.np_to_xfe(ca, first.field)), xfield)
                elif lazy:
                    values = DeviceCodeword(self, out[3 * c:3 * c + 3], xfield, "x")
                else:
                    values = B.np_to_xfe(a[3 * c:3 * c + 3], xfield)  # scale by lift(offset): offset.field = xfield
                    self.remember_planes(values, out[3 * c:3 * c + 3])
            res.append(values)
        return res

    def lift_codewords(self, xfield, codewords):
        """`[[xfield.lift(c) for c in cdwd] for cdwd in codewords]`, the last statement of every Table.extend of the
        reference (e.g. code/io_table.py:106-107).  A base-field DeviceCodeword becomes a lifted VIEW of itself (its
        elements wrap the base codeword's own element objects, as lift does); anything else is lifted as written."""
        return [DeviceCodeword(self, cw._planes, xfield, "l", base=cw) if type(cw) is DeviceCodeword and cw.kind == "b"
                else [xfield.lift(c) for c in cw] for cw in codewords]

    def rows_of(self, codewords):
        """`list(zip(*codewords))` (code/brainfuck_stark.py:178, :196): rows built on first access when the codewords
        are device views, the reference's own list otherwise"""
        if self._lazy and self._kept is not None and codewords and any(type(c) is DeviceCodeword for c in codewords):
            return LazyRows(self, codewords)
        return list(zip(*codewords))

    # ------------------------------------------------------------------ code/table.py quotients
    def compile_constraints(self, constraints, n_vars):
        """Flatten MPolynomials (code/multivariate.py: dict exponent-vector -> coefficient) into the
        monomial program of b2s_quotients.  Same checks as MPolynomial.evaluate (:105-116)."""
        mono_off, coeffs, facs = [0], [], []
        for mpo in constraints:
            for k, v in mpo.dictionary.items():
                assert n_vars == len(k), \
                    f"number of elements in point {n_vars} does not match with number of variables {len(k)} for polynomial {str(mpo)}"
                if self.B.is_xfe(v):
                    c = [co.value for co in v.polynomial.coefficients]
                else:
                    c = [v.value]
                coeffs.append(c + [0] * (3 - len(c)))
                f = [(i << 8) | e for i, e in enumerate(k) if e]
                if any(e > 255 for e in k):
                    raise ValueError("exponent above 255 in a constraint polynomial")
                facs.append(f)
            mono_off.append(len(coeffs))
        width = max([len(f) for f in facs] + [1])
        factors = np.zeros((len(facs), width), dtype=np.uint32)
        for m, f in enumerate(facs):
            factors[m, :len(f)] = f
        return (np.asarray(mono_off, dtype=np.uint32), np.asarray(coeffs, dtype=np.uint64).reshape(-1, 3), factors)

    def _table_planes(self, codewords, width, N):
        """((width, 3, N) device tensor of a table's codewords, flags of the lifted base-field columns among them).  Columns that came out of a device op inside
        keep_planes() are copied on the device, the others marshalled; inside keep_planes() the assembled tensor
        is kept for the table's next call (boundary, transition and terminal quotients read the same lists)."""
        eng = self.engine
        key = None
        if self._kept is not None:
            key = ("table", id(codewords))
            ent = self._kept.get(key)
            if ent is not None and ent[0] is codewords and len(ent[2]) == width and \
                    all(a is b for a, b in zip(ent[2], codewords)):
                return ent[1], ent[3]
        cw = eng.alloc((width, 3, N))
        base_columns = np.zeros(width, dtype=np.uint8)  # columns known to be lifted base-field codewords
        for j in range(width):
            planes = self.planes_of(codewords[j])
            lifted = self.lifted_planes_of(codewords[j]) if planes is None else None
            if planes is not None and planes.shape[0] == 1 and type(codewords[j]) is DeviceCodeword:
                planes, lifted = None, planes  # the lifted view of a base-field codeword (or the codeword itself)
            if planes is not None and planes.shape[0] == 3:
                eng.copy(cw[j], planes)
            elif lifted is not None:  # a base-field codeword lifted element by element: c0 = the plane, c1 = c2 = 0
                eng.copy(cw[j, 0:1], lifted)
                eng.zero(cw[j, 1:3])
                base_columns[j] = 1
            else:
                eng.upload_into(cw[j], self.B.xfe_to_np(codewords[j]))
        if key is not None:
            self._kept[key] = (codewords, cw, list(codewords[:width]), base_columns)
        return cw, base_columns

    @staticmethod
    def _zerofier_may_vanish(domain, kind, height, omicron_inv):
        """False when no zerofier of code/table.py:161-163, :194-201, :256-259 can vanish on offset * <omega>, decided
        on the host: the roots (1, the powers of omicron, omicron^-1) are N-th roots of unity when omicron^N = 1,
        and no point of the domain is one when offset^N != 1.  The engine then skips its read-back of the
        device's own check (code/ntt.py:178-179) and the tables' kernels run back to back."""
        N = domain.length
        if N < 1 or pow(int(domain.omega.value), N, P) != 1 or pow(int(domain.offset.value), N, P) == 1:
            return True
        if kind == ZEROFIER_BOUNDARY:
            return False
        if kind == ZEROFIER_TRANSITION:  # x^height = 1 implies x^N = 1 when height divides N
            return height <= 0 or N % height != 0
        return pow(int(omicron_inv), N, P) != 1

    @staticmethod
    def _element_field(codeword):
        """codeword[0].field without building the element of a device view"""
        return codeword._field if type(codeword) is DeviceCodeword else codeword[0].field

    def quotient_codewords(self, domain, codewords, width, constraints, kind, height=0, omicron_inv=1, shift=0):
        """code/table.py:155-178 / :190-236 / :253-286: [mpo.evaluate(point_i) * lift(zerofier_inverse[i])]
        for every constraint over the whole FRI domain, on the device."""
        N = domain.length
        n_vars = 2 * width if kind == ZEROFIER_TRANSITION else width
        program = self.compile_constraints(constraints, n_vars)
        xfield = self._element_field(codewords[0])  # acc = point[0].field.zero() (code/multivariate.py:106)
        cw, base_columns = self._table_planes(codewords, width, N)
        out, vanishes = self.engine.quotients(cw, shift, *program, kind, height, omicron_inv, domain.offset.value,
                                              domain.omega.value, base_columns=base_columns,
                                              check_zerofier=self._zerofier_may_vanish(domain, kind, height,
                                                                                       omicron_inv))
        # code/ntt.py:178-179
        assert not vanishes, "batch inverse does not work when input contains a zero"
        if self._kept is not None:
            # Inside keep_planes() (the drop-in's BrainfuckStark.prove): the prover only feeds quotient codewords
            # to the nonlinear combination, which reads the planes in place -- materialising N Python elements
            # per constraint would dominate prove().  Elements are built on first access; len / [] / iteration
            # behave like the reference's list.
            return [DeviceCodeword(self, out[c], xfield) for c in range(out.shape[0])]
        a = self.engine.download(out.reshape(-1, N)).reshape(-1, 3, N)
        return [self.B.np_to_xfe(a[c], xfield) for c in range(a.shape[0])]

    # ------------------------------------------------------------------ code/brainfuck_stark.py:241-298
    def combination_codeword(self, domain, xfield, max_degree, randomizer_codeword, base_codewords, base_degree_bounds,
                             extension_codewords, extension_degree_bounds, quotient_codewords,
                             quotient_degree_bounds, weights):
        """The nonlinear combination of code/brainfuck_stark.py:241-298 as one device op: every codeword
        contributes weight * c and weight' * (x^shift * c), shift = max_degree - degree bound.  Codewords
        that came out of device ops inside keep_planes() are read where they are; other lists are
        marshalled.  Returns a DeviceCodeword (the list the reference builds, materialised lazily),
        or None when the result's object graph could not be guaranteed (non-canonical weights): the
        caller then runs the reference's own block."""
        B = self.B
        N = domain.length
        n_terms = 1 + 2 * (len(base_codewords) + len(extension_codewords) + len(quotient_codewords))
        # code/brainfuck_stark.py:246-247, :258-259, :270-271, :294-295
        assert len(base_codewords) == len(base_degree_bounds) and len(extension_codewords) == len(extension_degree_bounds)
        assert len(quotient_codewords) == len(quotient_degree_bounds)
        assert n_terms == len(weights), f"number of terms {n_terms} is not equal to number of weights {len(weights)}"
        if N < 2 or not all(B.is_xfe(w) for w in weights) or not B.xfe_canonical(weights, xfield):
            return None
        columns, shifts = [], [0]
        bounds = list(base_degree_bounds) + list(extension_degree_bounds) + list(quotient_degree_bounds)
        for k, cw in enumerate([randomizer_codeword] + list(base_codewords) + list(extension_codewords) +
                               list(quotient_codewords)):
            if len(cw) != N:
                return None
            planes = self.planes_of(cw)
            if planes is None:
                if B.is_xfe(cw[0]):
                    planes = self.engine.upload(B.xfe_to_np(cw))
                elif B.is_bfe(cw[0]):
                    planes = self.engine.upload(B.bfe_to_np(cw))
                else:
                    return None
            columns.append(planes)
            if k:
                shift = max_degree - bounds[k - 1]
                if shift < 0:
                    return None
                shifts.append(shift)
        w = B.xfe_to_np(weights).T  # (n_terms, 3)
        wa = np.concatenate([w[:1], w[1::2]])
        wb = np.concatenate([np.zeros((1, 3), dtype=np.uint64), w[2::2]])
        out = self.engine.combination(columns, wa, wb, shifts, N, domain.offset.value, domain.omega.value)
        return DeviceCodeword(self, out, xfield)

    def table_boundary_quotients(self, table, fri_domain, codewords, challenges):
        """code/table.py:155-178"""
        assert len(codewords) != 0, "'codewords' argument must have nonzero length"
        return self.quotient_codewords(fri_domain, codewords, table.full_width, table.boundary_constraints_ext(challenges),
                                       ZEROFIER_BOUNDARY)

    def table_transition_quotients(self, table, domain, codewords, challenges):
        """code/table.py:190-236"""
        return self.quotient_codewords(domain, codewords, table.full_width, table.transition_constraints_ext(challenges),
                                       ZEROFIER_TRANSITION, height=table.height,
                                       omicron_inv=table.omicron.inverse().value,
                                       shift=table.unit_distance(domain.length))

    def table_terminal_quotients(self, table, domain, codewords, challenges, terminals):
        """code/table.py:253-286"""
        return self.quotient_codewords(domain, codewords, table.full_width,
                                       table.terminal_constraints_ext(challenges, terminals), ZEROFIER_TERMINAL,
                                       omicron_inv=table.omicron.inverse().value)

    def permutation_quotient(self, pa, fri_domain, MPolynomial):
        """code/permutation_argument.py:11-20: (lhs - rhs) * lift(1 / (x - 1)) as a two-variable program"""
        lhs = pa.all_tables[pa.lhs[0]].codewords[pa.lhs[1]]
        rhs = pa.all_tables[pa.rhs[0]].codewords[pa.rhs[1]]
        xfield = self._element_field(lhs)
        difference = MPolynomial({(1, 0): xfield.one(), (0, 1): -xfield.one()})
        return self.quotient_codewords(fri_domain, [lhs, rhs], 2, [difference], ZEROFIER_BOUNDARY)[0]

    # ------------------------------------------------------------------ code/merkle.py
    def merkle_build(self, tree, data_array, device_planes=None, device_nodes=None, leaf_cache=None,
                     canonical=None):
        """Body of Merkle.__init__ (code/merkle.py:8-41).  Sets the public attributes
        num_leafs, depth, leafs, nodes on `tree`."""
        n = len(data_array)
        if isinstance(data_array, DeviceCodeword) and leaf_cache is None and n > 0 and n & (n - 1) == 0:
            # a codeword that never left the device: its lazily materialised elements ARE the leaves
            leaf_cache, device_planes, canonical = data_array, data_array._planes, True
        tree.num_leafs = n
        npo2 = 1
        while npo2 < n:
            npo2 <<= 1
        if n == 0:
            npo2 = 0
        tree.depth = _ilog2(npo2) if npo2 else 0  # code/merkle.py:16-20
        tree.leafs = leaf_cache if leaf_cache is not None else [leaf for leaf in data_array]
        if n == 0:
            tree.nodes = []  # root() -> IndexError like the reference's nodes[1] on an empty list
            return
        eng = self.engine
        if device_nodes is None:
            first = tree.leafs[0]
            planes = None
            if canonical is None:
                canonical = n == npo2 and self.B.is_xfe(first) and self.B.xfe_canonical(tree.leafs)
            if canonical:
                planes = device_planes if device_planes is not None else eng.upload(self.B.xfe_to_np(tree.leafs))
                device_nodes = eng.merkle_field(planes, self.xfe_templates(first.field))
            elif n == npo2 and self.B.is_bfe(first) and all(
                    type(v) is self.B.BaseFieldElement and v.field is first.field for v in tree.leafs):
                planes = eng.upload(self.B.bfe_to_np(tree.leafs))
                device_nodes = eng.merkle_field(planes, self.bfe_templates(first.field))
            else:  # arbitrary picklable leaves (code/test_merkle.py:57-61) or non-canonical identity
                device_nodes = eng.merkle_blobs([pickle.dumps(leaf) for leaf in tree.leafs])
            tree._planes = planes
        else:
            tree._planes = device_planes
        tree._device_nodes = device_nodes
        tree.nodes = NodeView(eng, device_nodes, npo2, n)

    def merkle_open(self, tree, index):
        """code/merkle.py:46-52"""
        return tree.nodes.open(index, tree.depth)

    def salted_merkle_build(self, tree, data_array, urandom):
        """Body of SaltedMerkle.__init__ (code/salted_merkle.py:8-46): leaf i =
        blake2b(pickle(element) | pickle(salt)), salts drawn with the caller's `urandom` in leaf order.
        The leaves are tuples whose elements carry different field objects, so the host
        pickles them and the device hashes the byte strings and builds the tree (b2s_merkle_blobs).
        open / verify / root stay the reference's: they only index .leafs and .nodes."""
        n = len(data_array)
        tree.num_leafs = n
        npo2 = 1
        while npo2 < n:
            npo2 <<= 1
        tree.depth = _ilog2(npo2)  # code/salted_merkle.py:10-22 (0 leaves -> depth 0 as well)
        from .marshal import bulk_allocation
        nodes = None
        if type(data_array) is LazyRows and n == npo2:
            # rows of device codewords (prove() under the drop-in): the salts are drawn as the reference draws them,
            # the (row, salt) pairs exist only for the leaves somebody opens
            salts = [urandom(24) for _ in range(n)]  # code/salted_merkle.py:25
            tree.leafs = LazyLeafs(data_array, salts, tree)
            nodes = self._row_tree(data_array, salts)
        if nodes is None:
            with bulk_allocation():
                if type(data_array) is LazyRows and n == npo2:
                    tree.leafs = list(tree.leafs)  # same salts, rows materialised
                else:
                    tree.leafs = [(element, urandom(24)) for element in data_array]  # code/salted_merkle.py:25
            # code/salted_merkle.py:23: with no leaves the reference's own consistency assert fires
            assert n != 0, "in SaltedMerkle.__init__, next_power_of_two = 0 =/= 1 << self.depth = 1"
            nodes = self._row_tree([leaf[0] for leaf in tree.leafs], [leaf[1] for leaf in tree.leafs]) if n == npo2 else None
        if nodes is None:  # rows the device templates cannot express: the host pickles, the device hashes
            dumps = pickle.dumps
            nodes = self.engine.merkle_blobs([dumps(e) + dumps(salt) for e, salt in tree.leafs])
        tree._device_nodes = nodes
        tree.nodes = NodeView(self.engine, tree._device_nodes, npo2, n)

    MAX_ROW_SHAPES = 32

    def _row_columns(self, rows):
        """device planes behind every column of zipped rows (list of (n,) views in row order), or None.
        Columns that came out of a device op inside keep_planes() are read in place; others are marshalled
        after checking that the whole column has ONE identity pattern (the row template assumes it)."""
        B, n = self.B, len(rows)
        first = rows[0]
        lazy = type(rows) is LazyRows
        if type(first) is not tuple or not first or \
                (not lazy and any(type(r) is not tuple or len(r) != len(first) for r in rows)):
            return None
        planes = []
        for k, e in enumerate(first):
            ent = self._kept.get(("first", id(e))) if self._kept is not None else None
            if lazy and type(rows.columns[k]) is DeviceCodeword:
                if rows.columns[k].kind == "l" or len(rows.columns[k]) != n:
                    return None
                col = rows.columns[k]._planes
            elif ent is not None and len(ent[0]) == n and ent[1].shape[1] == n and \
                    all(rows[i][k] is v and ent[0][i] is v for i, v in zip(ent[2], ent[3])):
                col = ent[1]
            else:
                column = rows.columns[k] if lazy else [r[k] for r in rows]
                if len(column) != n:
                    return None
                if B.is_bfe(e):
                    f = e.field
                    if not all(type(v) is B.BaseFieldElement and v.field is f for v in column):
                        return None
                    col = self.engine.upload(B.bfe_to_np(column))
                elif B.is_xfe(e) and B.xfe_canonical(column, e.field):
                    col = self.engine.upload(B.xfe_to_np(column))
                else:
                    return None
            if col.shape[0] != (1 if B.is_bfe(e) else 3) or col.stride(1) != 1:
                return None
            planes += [col[q] for q in range(col.shape[0])]
        return planes

    def _row_tree(self, rows, salts):
        """Device tree over salted rows: the row pickle is a per-tree byte template
        derived from a sample row; the device splices the integers of the codeword planes and the salts.
        rows: a list of tuples or LazyRows.  Returns the node tensor, or None when the rows cannot be expressed
        (the caller then pickles)."""
        from . import marshal
        n = len(salts)
        if n == 0 or len(rows) != n:
            return None
        salt0 = salts[0]
        frame = marshal.salt_frame(salt0) if type(salt0) is bytes else None
        if frame is None or set(map(type, salts)) != {bytes} or set(map(len, salts)) != {len(salt0)}:
            return None
        tpl = marshal.row_template(self.B, rows[0])
        if tpl is None:
            return None
        planes = self._row_columns(rows)
        if planes is None or len(planes) != len(tpl.modes):
            return None
        eng = self.engine
        salts = eng.upload_bytes(np.frombuffer(b"".join(salts), dtype=np.uint8).reshape(n, len(salt0)))
        # a couple of rows rendered on the host as well: the template must reproduce the caller's pickler
        for i in {0, n // 2, n - 1}:
            t = tpl if marshal.row_signature(self.B, rows[i]) == tpl.signature else marshal.row_template(self.B, rows[i])
            if t is None or t.render(marshal.row_values(rows[i])) != pickle.dumps(rows[i]):
                return None
        nodes, exc = eng.merkle_rows(planes, tpl.modes, tpl.tpl, tpl.seg_off, n, salts, frame[0], frame[1])
        shapes = 1
        while len(exc):  # rows of another shape (trimmed extension-field coefficients): one more template each
            shapes += 1
            t = marshal.row_template(self.B, rows[int(exc[0])])
            if t is None or shapes > self.MAX_ROW_SHAPES:
                return None
            todo = eng.upload_bytes(np.sort(exc).astype(np.uint32).view(np.uint8)).view(torch.int32)
            nodes, rest = eng.merkle_rows(planes, t.modes, t.tpl, t.seg_off, n, salts, frame[0], frame[1], rows=todo,
                                          nodes=nodes, build_upper=False)
            if len(rest) == len(exc):
                return None  # no progress: the sample row itself does not fit its own template
            exc = rest
            if not len(exc) and n > 1:
                eng.merkle_upper(nodes)
        return nodes

    # ------------------------------------------------------------------ code/fri.py Fri
    def fri_commit(self, fri, codeword, proof_stream, round_index=0, Merkle=None):
        """code/fri.py:91-139.  The round loop stays on the host because each challenge is a
        hash of the pickled proof stream (code/fri.py:120 -> code/ip.py:21-22); per round the
        device folds the codeword and builds the next tree in one fused call."""
        B = self.B
        xfield = fri.field
        eng = self.engine
        num_rounds = fri.num_rounds()
        omega = self.base_value(fri.domain.omega)
        offset = self.base_value(fri.domain.offset)
        tpl = self.xfe_templates(xfield)
        trees, codewords = [], []

        N = len(codeword)
        if isinstance(codeword, DeviceCodeword) and N > 0 and (N & (N - 1)) == 0:
            canonical, planes = True, codeword._planes  # e.g. the nonlinear combination: already on the device
        else:
            canonical = N > 0 and (N & (N - 1)) == 0 and B.is_xfe(codeword[0]) and B.xfe_canonical(codeword, xfield)
            if N > 0 and not B.is_xfe(codeword[0]):
                # code/fri.py:127: (one + alpha / ...) * codeword[i] needs extension-field elements
                raise AttributeError("%r object has no attribute 'polynomial'" % type(codeword[0]).__name__)
            planes = eng.upload(B.xfe_to_np(codeword)) if N > 0 else None
        nodes = None
        cache = None  # identity-stable objects of the current (device) codeword
        for r in range(num_rounds):
            N = len(codeword)
            # code/fri.py:104-105
            assert pow(omega, N - 1, P) == pow(omega, P - 2, P), "error in commit: omega does not have the right order!"
            tree = Merkle.__new__(Merkle)
            if r == 0:
                self.merkle_build(tree, codeword, device_planes=planes, canonical=canonical)
            else:
                self.merkle_build(tree, codeword, device_planes=planes, device_nodes=nodes, leaf_cache=cache)
            root = tree.root()
            if r > 0:
                proof_stream.push(root)
            if r == num_rounds - 1:
                break
            alpha = xfield.sample(proof_stream.prover_fiat_shamir())
            codewords.append(codeword)
            trees.append(tree)
            a = [c.value for c in alpha.polynomial.coefficients]
            a += [0] * (3 - len(a))
            planes, nodes = eng.fri_fold(planes, a, offset, omega, tpl)
            cache = DeviceCodeword(self, planes, xfield)
            codeword = cache
            omega = omega * omega % P
            offset = offset * offset % P
        # send last codeword: a real list of real objects, shared with query_last (code/fri.py:134, :169)
        if isinstance(codeword, DeviceCodeword):
            codeword = codeword.materialize()
        proof_stream.push(codeword)
        codewords.append(codeword)
        return codewords, trees

    def fri_query(self, fri, current_tree, next_tree, c_indices, proof_stream):
        """code/fri.py:141-158"""
        half = len(current_tree.leafs) // 2
        a_indices = [i for i in c_indices]
        b_indices = [i + half for i in c_indices]
        s = fri.num_colinearity_tests
        prefetch(current_tree.leafs, a_indices[:s] + b_indices[:s])
        prefetch(next_tree.leafs, c_indices[:s])
        for k in range(s):
            proof_stream.push((current_tree.leafs[a_indices[k]], current_tree.leafs[b_indices[k]],
                               next_tree.leafs[c_indices[k]]))
        prefetch_paths(current_tree, a_indices[:s] + b_indices[:s])
        prefetch_paths(next_tree, c_indices[:s])
        for k in range(s):
            proof_stream.push(current_tree.open(a_indices[k]))
            proof_stream.push(current_tree.open(b_indices[k]))
            proof_stream.push(next_tree.open(c_indices[k]))
        return a_indices + b_indices

    def fri_query_last(self, fri, current_tree, last_codeword, c_indices, proof_stream):
        """code/fri.py:160-176"""
        half = len(current_tree.leafs) // 2
        a_indices = [i for i in c_indices]
        b_indices = [i + half for i in c_indices]
        s = fri.num_colinearity_tests
        prefetch(current_tree.leafs, a_indices[:s] + b_indices[:s])
        for k in range(s):
            proof_stream.push((current_tree.leafs[a_indices[k]], current_tree.leafs[b_indices[k]],
                               last_codeword[c_indices[k]]))
        prefetch_paths(current_tree, a_indices[:s] + b_indices[:s])
        for k in range(s):
            proof_stream.push(current_tree.open(a_indices[k]))
            proof_stream.push(current_tree.open(b_indices[k]))
        return a_indices + b_indices

    def fri_prove(self, fri, codeword, proof_stream):
        """code/fri.py:178-199"""
        assert fri.domain.length == len(codeword), "initial codeword length does not match length of initial codeword"
        codewords, trees = fri.commit(codeword, proof_stream)
        top_level_indices = fri.sample_indices(proof_stream.prover_fiat_shamir(), len(codewords[1]),
                                               len(codewords[-1]), fri.num_colinearity_tests)
        # every round's indices follow from the top-level ones (code/fri.py:191-197): fetch each tree's leaves
        # and authentication paths with ONE device call per tree before the query loop pushes them
        s = fri.num_colinearity_tests
        idx = [i for i in top_level_indices]
        wanted = [[] for _ in trees]
        for i in range(len(trees)):
            half = len(codewords[i]) // 2
            idx = [index % half for index in idx]
            wanted[i] += idx[:s] + [j + half for j in idx[:s]]
            if i + 1 < len(trees):
                wanted[i + 1] += idx[:s]
        prefetch_openings(self.engine, trees, wanted)
        indices = [i for i in top_level_indices]
        for i in range(len(trees) - 1):
            indices = [index % (len(codewords[i]) // 2) for index in indices]
            fri.query(trees[i], trees[i + 1], indices, proof_stream)
        indices = [index % len(codewords[-1]) for index in indices]
        fri.query_last(trees[-1], codewords[-1], indices, proof_stream)
        return top_level_indices


def prefetch_openings(engine, trees, wanted):
    """leaves and authentication paths of several trees with ONE device call (b2s_open_multi): what the query
    phase of a proof opens.  Trees whose leaves / nodes are not plain device views (host lists, the sharded
    views of dist_fri) go through their own prefetch."""
    sets, fills = [], []
    for tree, w in zip(trees, wanted):
        leafs, nodes = tree.leafs, getattr(tree, "nodes", None)
        if type(leafs) is DeviceCodeword:
            need = leafs.wanted(w)
            if need:
                sets.append((leafs._planes, None, need))
                fills.append(lambda res, leafs=leafs, need=need: leafs.fill(need, res[0]))
        else:
            prefetch(leafs, w)
        if type(nodes) is NodeView:
            need = nodes.wanted_paths(w, tree.depth)
            if need:
                sets.append((None, nodes._dev, need))
                fills.append(lambda res, nodes=nodes, need=need, depth=tree.depth: nodes.fill_paths(need, res[1], depth))
        else:
            prefetch_paths(tree, w)
    if sets:
        for fill, res in zip(fills, engine.open_multi(sets)):
            fill(res)


def prefetch(leafs, indices):
    if isinstance(leafs, DeviceCodeword):
        leafs.prefetch(indices)


def prefetch_paths(tree, indices):
    nodes = getattr(tree, "nodes", None)
    if isinstance(nodes, NodeView):
        nodes.prefetch_paths(indices, tree.depth)


ZEROFIER_BOUNDARY, ZEROFIER_TRANSITION, ZEROFIER_TERMINAL = 1, 2, 3


class DeviceCodeword:
    """A codeword that lives on the device.  Behaves like the list of field elements the reference builds (the
    folded codewords of code/fri.py:127-128, the codewords of Table.lde / ldex, Domain.xevaluate) for the accesses
    the reference makes -- len, indexing, slicing, iteration; elements are materialised on demand and cached, so
    repeated access returns the SAME object (pickle memoises by identity, SURVEY B5 rule 3).
    kind "x": ExtensionFieldElements of `field`, planes (3, n)
         "b": BaseFieldElements of `field`, planes (1, n)
         "l": `[xfield.lift(c) for c in base]` of a kind-"b" codeword `base` (what every Table.extend does to its
              base codewords): element i wraps base[i] ITSELF as its only coefficient (none when it is zero,
              code/extension_field.py:6-9, :113-116); planes = the base codeword's"""

    kind, _base, _share, _reps = "x", None, 0, None  # defaults for subclasses with their own constructor (dist_fri)

    def __init__(self, glue, planes, field, kind="x", base=None, share=0, reps=None):
        assert planes.shape[0] == (3 if kind == "x" else 1) and (not share or kind == "x")
        self._glue = glue
        self._planes = planes
        self._field = field
        self.kind = kind
        self._base = base
        self._n = planes.shape[1]
        self._cache = {}
        # share > 0 (kind "x"): the identity graph of the reference's recursive ntt on a sparse input
        # (Glue._shared_output_period): element i is a distinct object wrapping the coefficient objects of
        # representative i % share; share == 1 with reps = {0: the input's lone coefficient element}
        self._share = share
        self._reps = {} if reps is None else reps

    def __len__(self):
        return self._n

    def wanted(self, indices):
        """the indices that still have to be fetched (range-checked like list indexing)"""
        cache = self._cache if self.kind != "l" else self._base._cache
        need = [i for i in dict.fromkeys(indices) if i not in cache]
        for i in need:
            if not 0 <= i < self._n:
                raise IndexError("list index out of range")
        return need

    def fill(self, need, vals):
        """vals: (len(need), number of planes) uint64 as gathered from the planes"""
        B = self._glue.B
        if self._share:
            mk, xf, X, reps, share = B.make_xfe, self._field, B.ExtensionFieldElement, self._reps, self._share
            for i, v in zip(need, vals.tolist()):
                rep = reps.get(i % share)
                if rep is None:  # outputs i and i % share are equal: either's values build the representative
                    rep = reps[i % share] = mk(v[0], v[1], v[2], xf)
                self._cache[i] = X(rep.polynomial, xf)
        elif self.kind == "x":
            mk, xf = B.make_xfe, self._field
            for i, v in zip(need, vals.tolist()):
                self._cache[i] = mk(v[0], v[1], v[2], xf)
        elif self.kind == "b":
            for i, e in zip(need, B.np_to_bfe(vals[:, 0], self._field)):
                self._cache[i] = e
        else:
            self._base.fill(need, vals)

    def prefetch(self, indices):
        need = self.wanted(indices)
        if need:
            self.fill(need, self._glue.engine.gather(self._planes, need))

    def _lift(self, e):
        B = self._glue.B
        p = B.Polynomial.__new__(B.Polynomial)
        p.__dict__ = {"coefficients": [e] if e.value != 0 else []}
        x = B.ExtensionFieldElement.__new__(B.ExtensionFieldElement)
        x.__dict__ = {"polynomial": p, "field": self._field}
        return x

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self._n))]
        if i < 0:
            i += self._n
        if i not in self._cache:
            if self.kind == "l":
                if not 0 <= i < self._n:
                    raise IndexError("list index out of range")
                self._cache[i] = self._lift(self._base[i])
            else:
                self.prefetch([i])
        return self._cache[i]

    def __iter__(self):
        return iter(self.materialize())

    def materialize(self):
        """the whole codeword as a real list (cached objects are reused)"""
        if len(self._cache) < self._n:
            from .marshal import bulk_allocation
            B = self._glue.B
            with bulk_allocation():
                if self.kind == "l":
                    base = self._base.materialize()
                    for i in range(self._n):
                        if i not in self._cache:
                            self._cache[i] = self._lift(base[i])
                elif self._share:
                    todo = [j for j in range(self._share) if j not in self._reps]
                    if todo:
                        a = self._glue.engine.download(self._planes)[:, :self._share]
                        fresh = B.np_to_xfe(a, self._field)
                        for j in todo:
                            self._reps[j] = fresh[j]
                    X, xf, reps, share = B.ExtensionFieldElement, self._field, self._reps, self._share
                    for i in range(self._n):
                        if i not in self._cache:
                            self._cache[i] = X(reps[i % share].polynomial, xf)
                else:
                    a = self._glue.engine.download(self._planes)
                    fresh = B.np_to_xfe(a, self._field) if self.kind == "x" else B.np_to_bfe(a[0], self._field)
                    for i in range(self._n):
                        if i not in self._cache:
                            self._cache[i] = fresh[i]
        return [self._cache[i] for i in range(self._n)]

    def __eq__(self, other):
        return list(self) == list(other)

    __hash__ = None


class LazyRows:
    """`list(zip(*codewords))` over codewords of which some are DeviceCodewords (code/brainfuck_stark.py:178, :196):
    row i is built on first access -- one device call gathers the row's elements from every device column -- and
    cached, so the tuple the prover pushes twice is the same object twice."""

    def __init__(self, glue, columns):
        self._glue = glue
        self.columns = list(columns)
        self._n = min(len(c) for c in self.columns)  # zip stops at the shortest
        self._cache = {}

    def __len__(self):
        return self._n

    def prefetch(self, indices, tree=None):
        """gather the elements of the given rows from every device column -- and, with `tree`, the authentication
        paths of those leaves (SaltedMerkle.open follows leafs[i] in code/brainfuck_stark.py:316-326) -- with ONE
        device call"""
        sets, fills = [], []
        for c in self.columns:
            if type(c) is DeviceCodeword:
                need = c.wanted(indices)
                if need:
                    for q in range(c._planes.shape[0]):
                        sets.append((c._planes[q:q + 1], None, need))
                    fills.append((c, need, c._planes.shape[0]))
        nodes = getattr(tree, "nodes", None)
        path_need = nodes.wanted_paths(indices, tree.depth) if type(nodes) is NodeView else []
        if path_need:
            sets.append((None, nodes._dev, path_need))
        if not sets:
            return
        res = self._glue.engine.open_multi(sets)
        k = 0
        for c, need, q in fills:
            c.fill(need, np.concatenate([res[k + j][0] for j in range(q)], axis=1))
            k += q
        if path_need:
            nodes.fill_paths(path_need, res[k][1], tree.depth)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self._n))]
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError("list index out of range")
        row = self._cache.get(i)
        if row is None:
            self.prefetch([i])
            row = self._cache[i] = tuple(c[i] for c in self.columns)
        return row

    def __iter__(self):
        cols = [c.materialize() if type(c) is DeviceCodeword else c for c in self.columns]
        for i in range(self._n):
            row = self._cache.get(i)
            if row is None:
                row = self._cache[i] = tuple(c[i] for c in cols)
            yield row


class LazyLeafs:
    """SaltedMerkle.leafs = [(element, salt)] (code/salted_merkle.py:25) over LazyRows: pairs built on first access
    and cached; the salts are the objects drawn at construction"""

    def __init__(self, rows, salts, tree=None):
        self.rows, self.salts, self._tree = rows, salts, tree
        self._cache = {}

    def __len__(self):
        return len(self.salts)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self.salts)))]
        if i < 0:
            i += len(self.salts)
        leaf = self._cache.get(i)
        if leaf is None:
            if 0 <= i < len(self.salts) and type(self.rows) is LazyRows:
                self.rows.prefetch([i], self._tree)
            leaf = self._cache[i] = (self.rows[i], self.salts[i])
        return leaf

    def __iter__(self):
        for i, row in enumerate(self.rows):
            leaf = self._cache.get(i)
            if leaf is None:
                leaf = self._cache[i] = (row, self.salts[i])
            yield leaf


class NodeView:
    """Merkle.nodes for a tree whose digests live on the device (code/merkle.py:26-41 heap
    layout).  Indexing returns `bytes` objects that are cached, so auth paths of different
    leaves share their upper siblings exactly like slices of the reference's list."""

    _ZERO32 = bytes(32)

    def __init__(self, engine, device_nodes, npo2, n_leafs):
        self._eng = engine
        self._dev = device_nodes
        self._npo2 = npo2
        self._n = n_leafs
        self._cache = {}

    def __len__(self):
        return 2 * self._npo2

    def _fetch_all(self):
        raw = self._eng.download_bytes(self._dev)
        for k in range(1, 2 * self._npo2):
            if k not in self._cache and not (k >= self._npo2 + self._n):
                self._cache[k] = raw[64 * k:64 * k + 64]

    def __getitem__(self, k):
        if isinstance(k, slice):
            self._fetch_all()
            return [self[j] for j in range(*k.indices(len(self)))]
        if k < 0:
            k += len(self)
        if not 0 <= k < len(self):
            raise IndexError("list index out of range")
        if k >= self._npo2 + self._n:
            return self._ZERO32  # unused leaf slot, code/merkle.py:26
        v = self._cache.get(k)
        if v is None:
            if k == 0:
                # code/merkle.py:35-41 also overwrites slot 0 with blake2b(placeholder | root)
                from hashlib import blake2b
                v = blake2b(self._ZERO32 + self[1]).digest() if self._npo2 >= 1 else self._ZERO32
            else:
                v = self._eng.download_bytes(self._dev[k])
            self._cache[k] = v
        return v

    def __iter__(self):
        self._fetch_all()
        return (self[k] for k in range(len(self)))

    def wanted_paths(self, indices, depth):
        """the leaf indices whose authentication path is not completely cached yet"""
        if depth == 0:
            return []
        cache, top = self._cache, 1 << depth
        need = [i for i in dict.fromkeys(indices) if 0 <= i < top]
        if len(cache) > 1:  # (a fresh tree holds at most its root: everything is wanted)
            need = [i for i in need if any(((top | i) >> j) ^ 1 not in cache for j in range(depth))]
        return need

    def fill_paths(self, need, paths, depth):
        """paths[q][j]: the 64-byte sibling at level j of leaf need[q] (lists of bytes, or a (len, depth, 64) uint8
        array)"""
        cache, limit = self._cache, self._npo2 + self._n
        raw = paths.tobytes() if hasattr(paths, "tobytes") else None  # one conversion, then slices
        for q, i in enumerate(need):
            k = (1 << depth) | i
            at = q * depth * 64
            for j in range(depth):
                sib = (k >> j) ^ 1
                if sib not in cache and sib < limit:
                    cache[sib] = raw[at + 64 * j:at + 64 * j + 64] if raw is not None else bytes(paths[q][j])

    def prefetch_paths(self, indices, depth):
        need = self.wanted_paths(indices, depth)
        if need:
            self.fill_paths(need, self._eng.merkle_open(self._dev, need), depth)

    def open(self, index, depth):
        """code/merkle.py:46-52"""
        cache = self._cache
        k = (1 << depth) | index
        path = []
        try:  # all siblings already fetched (the usual case after prefetch_paths)
            while k > 1:
                path.append(cache[k ^ 1])
                k >>= 1
            return path
        except KeyError:
            pass
        self.prefetch_paths([index], depth)
        path = []
        k = (1 << depth) | index
        while k > 1:
            path += [self[k ^ 1]]
            k >>= 1
        return path
