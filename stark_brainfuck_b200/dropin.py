"""Drop-in for an existing checkout of the reference: after install(), the reference's own
modules run their hot path on the device and brainfuck_stark.BrainfuckStark.prove() runs
UNMODIFIED (no reference file is edited).

The reference has no plugin layer: its modules star-import each other by bare name
(code/fri.py:1-10, code/table.py:3-4, code/brainfuck_stark.py:1-17), so a binding copied at import
time lives in every importing module's globals.  install() therefore
  1. imports algebra, univariate, extension_field, ntt, merkle, ip, fri from `reference_dir`;
  2. rebinds ntt / intt / fast_coset_evaluate / fast_coset_interpolate in `ntt`'s globals (so
     fast_multiply & co. pick them up) and in every loaded module whose global IS the original;
  3. patches class attributes, which every importer shares: Polynomial.scale/evaluate_domain,
     Fri.Domain.evaluate/xevaluate/interpolate/xinterpolate, Fri.commit/query/query_last/prove,
     Merkle.__init__/open (Merkle.root/verify and Fri.verify stay the reference's), and the per-element host
     arithmetic of SURVEY 8(a) a1 / a3 by direct equivalents with the same values and object graph:
     ExtensionField.lift/sample/multiply/add/subtract/negate, BaseField.sample, Polynomial.degree.
  4. next rows: Table.*_quotients / PermutationArgument.quotient, SaltedMerkle.__init__, and the
     nonlinear combination.  The combination is INLINE in BrainfuckStark.prove
     (code/brainfuck_stark.py:241-298), so there is no attribute to rebind: prove() is recompiled
     at install time from the reference's own source with that one block guarded by a call into
     the glue (the block itself stays as the fallback) -- the in-memory equivalent of the hunk
     shown in INTEGRATION.md.
  5. inside prove() the codewords are device views (glue.DeviceCodeword) whose elements are built on first access:
     Table.lde / ldex / Domain.xevaluate return them, the `list(zip(*codewords))` statements of prove()
     (code/brainfuck_stark.py:178, :196) go through a hook compiled into the same recompiled prove(), and the four
     Table.extend methods are wrapped so that their last statement (`self.codewords = [[xfield.lift(c) ...`, e.g.
     code/io_table.py:106-107) runs over an empty list while the glue lifts the views.  A proof opens a few hundred
     rows of its 46 codewords; building every element object was most of its host time.
uninstall() restores every original (needed to time the CPU reference in the same process).
"""
import importlib
import inspect
import os
import sys
import textwrap
import warnings

from .glue import Glue
from .marshal import Binding, bulk_allocation

_state = None

_NTT_FUNCS = ("ntt", "intt", "fast_coset_evaluate", "fast_coset_interpolate")


def installed():
    return _state is not None


def current_glue():
    return _state["glue"] if _state else None


_COMB_BEGIN = "# compute terms of nonlinear combination polynomial"
_COMB_END = "# commit to combination codeword"
_COMB_HOOK = "__b2s_combination__"
_ROWS_HOOK = "__b2s_rows__"
# the two transpositions of prove() (code/brainfuck_stark.py:178, :196): rows of codewords for the salted trees
_ROWS_STATEMENTS = ("zipped_codeword = list(zip(*all_base_codewords))",
                    "zipped_extension_codeword = list(zip(*extension_codewords))")
# last statement of ProcessorTable / InstructionTable / MemoryTable.extend and IOTable.extend_iotable
_LIFT_STATEMENT = "self.codewords = [[xfield.lift(c) for c in cdwd]"


def lazy_rows_source(src):
    """prove() source with `list(zip(*codewords))` routed through the rows hook (which returns the same list unless
    the codewords are device views).  A statement that is not where this reference version has it stays as it is:
    zip then iterates the device views, which materialises them -- slower, same result."""
    for stmt in _ROWS_STATEMENTS:
        if src.count(stmt) == 1:
            name, arg = stmt.split(" = list(zip(*")
            src = src.replace(stmt, "%s = %s(%s)" % (name, _ROWS_HOOK, arg.rstrip(")")))
    return src


def guarded_prove_source(prove):
    """Source of BrainfuckStark.prove with its nonlinear-combination block (between the two marker
    comments, code/brainfuck_stark.py:241-298) kept as the fallback of a call to the hook.  Raises
    LookupError when the markers are not where this reference version has them."""
    src = textwrap.dedent(inspect.getsource(prove))
    a, b = src.find(_COMB_BEGIN), src.find(_COMB_END)
    if a < 0 or b < a or src.find(_COMB_BEGIN, a + 1) >= 0:
        raise LookupError("nonlinear-combination block of BrainfuckStark.prove not found")
    a, b = src.rindex("\n", 0, a) + 1, src.rindex("\n", 0, b) + 1
    block = src[a:b]
    if "combination_codeword = reduce(" not in block:
        raise LookupError("nonlinear-combination block of BrainfuckStark.prove has an unexpected shape")
    pad = block[:len(block) - len(block.lstrip())]
    call = (pad + "combination_codeword = " + _COMB_HOOK + "(self.fri.domain, self.xfield, self.max_degree, "
            "randomizer_codeword, base_codewords, base_degree_bounds, extension_codewords, extension_degree_bounds, "
            "quotient_codewords, quotient_degree_bounds, weights)\n" + pad + "if combination_codeword is None:\n")
    return lazy_rows_source(src[:a] + call + textwrap.indent(block, "    ") + src[b:])


def install(reference_dir=None, engine=None, quotients=True, salted=True, combination=True, lde=True, lazy=None):
    """Patch the reference modules importable from `reference_dir` (or already on sys.path).
    `quotients=True` also moves the quotient-codeword loops of table.py / permutation_argument.py
    (93 % of prove()) to the device, `salted=True` the trees of salted_merkle.py,
    `combination=True` the nonlinear combination inside BrainfuckStark.prove (see module docstring),
    `lde=True` Table.interpolate_columns / lde / ldex as batched transforms over all columns of a table.
    `lazy` (default: on unless B2S_LAZY_CODEWORDS=0): inside prove() the codewords of lde / ldex / xevaluate are device
    views whose elements are built on first access (glue.DeviceCodeword); needs quotients, salted, combination and lde.
    Returns the Glue in use."""
    global _state
    if _state is not None:
        return _state["glue"]
    import pickle
    if pickle.DEFAULT_PROTOCOL != 4:
        # The reference hashes pickle.dumps(leaf) with the interpreter's default protocol; the device-side leaf
        # emitter (csrc/leaf.cuh) and the templates of marshal.py reproduce protocol 4's framing and integer
        # opcodes (CPython 3.8 - 3.13).  Under another default the reference's own bytes change as well.
        raise RuntimeError("stark_brainfuck_b200 reproduces pickle protocol 4 (CPython 3.8-3.13); this interpreter's "
                           "default protocol is %d" % pickle.DEFAULT_PROTOCOL)
    if reference_dir is not None and reference_dir not in sys.path:
        sys.path.insert(0, reference_dir)
    if lazy is None:
        lazy = os.environ.get("B2S_LAZY_CODEWORDS", "1") != "0"
    lazy = bool(lazy and quotients and salted and combination and lde) and os.environ.get("DEBUG") is None
    mods = {name: importlib.import_module(name)
            for name in ("algebra", "univariate", "extension_field", "ntt", "merkle", "ip", "fri")}
    binding = Binding.from_modules(mods["algebra"], mods["univariate"], mods["extension_field"])
    glue = Glue(binding, engine)
    saved = {"globals": [], "attrs": []}

    def set_attr(obj, name, value):
        saved["attrs"].append((obj, name, obj.__dict__.get(name, _MISSING)))
        setattr(obj, name, value)

    # -- 2. module-level functions ------------------------------------------------------
    replacements = {"ntt": glue.ntt, "intt": glue.intt, "fast_coset_evaluate": glue.fast_coset_evaluate,
                    "fast_coset_interpolate": glue.fast_coset_interpolate}
    originals = {name: getattr(mods["ntt"], name) for name in _NTT_FUNCS}
    for mod in list(sys.modules.values()):
        d = getattr(mod, "__dict__", None)
        if not d:
            continue
        for name in _NTT_FUNCS:
            if d.get(name) is originals[name]:
                saved["globals"].append((mod, name, originals[name]))
                d[name] = replacements[name]

    # -- 3. class attributes --------------------------------------------------------------
    Polynomial, Fri, Merkle = mods["univariate"].Polynomial, mods["fri"].Fri, mods["merkle"].Merkle
    set_attr(Polynomial, "scale", lambda self, factor: glue.poly_scale(self, factor))
    set_attr(Polynomial, "evaluate_domain", lambda self, domain: glue.poly_evaluate_domain(self, domain))
    D = Fri.Domain
    set_attr(D, "evaluate", lambda self, polynomial: glue.domain_evaluate(self, polynomial))
    set_attr(D, "xevaluate", lambda self, polynomial, xfield=None: glue.domain_xevaluate(self, polynomial, xfield))
    set_attr(D, "interpolate", lambda self, values: glue.domain_interpolate(self, values))
    set_attr(D, "xinterpolate", lambda self, values: glue.domain_xinterpolate(self, values))
    set_attr(Fri, "commit", lambda self, codeword, proof_stream, round_index=0:
             glue.fri_commit(self, codeword, proof_stream, round_index, Merkle=Merkle))
    set_attr(Fri, "query", lambda self, current_tree, next_tree, c_indices, proof_stream:
             glue.fri_query(self, current_tree, next_tree, c_indices, proof_stream))
    set_attr(Fri, "query_last", lambda self, current_tree, last_codeword, c_indices, proof_stream:
             glue.fri_query_last(self, current_tree, last_codeword, c_indices, proof_stream))
    set_attr(Fri, "prove", lambda self, codeword, proof_stream: glue.fri_prove(self, codeword, proof_stream))
    set_attr(Merkle, "__init__", lambda self, data_array: glue.merkle_build(self, data_array))
    set_attr(Merkle, "open", lambda self, index: glue.merkle_open(self, index))

    # ExtensionField.lift (code/extension_field.py:113-116).  Every Table.extend lifts its base
    # codewords element by element (e.g. code/io_table.py:106-107): 2.1 M calls in a Hello-World proof, each through
    # three constructors and two Polynomial.degree() scans.  Same object graph, built directly: the element wraps
    # the base element ITSELF as its only coefficient, or no coefficient when it is zero (code/extension_field.py:6-9).
    X, Pn = binding.ExtensionFieldElement, binding.Polynomial

    def lift(self, base_field_element):
        if type(base_field_element) == X:
            return base_field_element
        p = Pn.__new__(Pn)
        p.__dict__ = {"coefficients": [base_field_element] if base_field_element.value != 0 else []}
        x = X.__new__(X)
        x.__dict__ = {"polynomial": p, "field": self}
        return x
    set_attr(binding.ExtensionField, "lift", lift)

    # ExtensionField.sample / BaseField.sample (code/extension_field.py:100-111, code/algebra.py:138-142; SURVEY 8(a)
    # a1 / a3): the prover draws its randomizer polynomial with max_degree + 1 = N/4 calls (code/brainfuck_stark.py:
    # 164-165), each a Python loop over the bytes, three constructors and two Polynomial.degree() scans.  Same
    # values ((acc << 8) ^ b over the bytes is the big-endian integer), same object graph, built directly.
    Bf, Bcls = binding.BaseField, binding.BaseFieldElement
    orig_bsample, orig_xsample = Bf.sample, binding.ExtensionField.sample

    def bsample(self, byte_array):
        if type(byte_array) is not bytes:
            return orig_bsample(self, byte_array)
        o = Bcls.__new__(Bcls)
        o.__dict__ = {"value": int.from_bytes(byte_array, "big") % self.p, "field": self}
        return o

    shape = {}  # id(extension field) -> (the field, degree of its modulus, its inner base field, p)

    def xsample(self, byte_array):
        if type(byte_array) is not bytes:
            return orig_xsample(self, byte_array)
        hit = shape.get(id(self))
        if hit is None or hit[0] is not self or hit[2] is not self.modulus.coefficients[0].field:
            inner = self.modulus.coefficients[0].field
            hit = shape[id(self)] = (self, self.modulus.degree(), inner, inner.p)
        _, parts, inner, p = hit
        width = len(byte_array) // parts
        frm = int.from_bytes
        vals = [frm(byte_array[k * width:(k + 1) * width], "big") % p for k in range(parts)]
        while vals and vals[-1] == 0:  # ExtensionFieldElement.__init__ trims (code/extension_field.py:6-9)
            vals.pop()
        co = []
        for v in vals:
            o = Bcls.__new__(Bcls)
            o.__dict__ = {"value": v, "field": inner}
            co.append(o)
        q = Pn.__new__(Pn)
        q.__dict__ = {"coefficients": co}
        x = X.__new__(X)
        x.__dict__ = {"polynomial": q, "field": self}
        return x
    set_attr(Bf, "sample", bsample)
    set_attr(binding.ExtensionField, "sample", xsample)

    # ExtensionField.multiply / add / subtract / negate (code/extension_field.py:65-75).  The reference
    # multiplies two elements through a schoolbook Polynomial product and a long division by the modulus, dozens of
    # temporary objects and degree() scans per call (85 us; every Table.extend is made of these: 190 000 products in a
    # proof on a 2^20 domain).  Same values AND the same object graph, which pickle sees when a result (a terminal)
    # goes into the proof stream -- derived from code/univariate.py:23-51, :90-111:
    #   product   all coefficients are new elements of left.coefficients[0].field; when the product needs no reduction
    #             (degree < 3) the positions no non-zero left coefficient reached hold ONE shared zero element
    #   sum       an empty operand returns the other one's coefficient OBJECTS (for a difference: their new negations,
    #             each in its own field); otherwise new elements of left.coefficients[0].field
    # Anything unusual (more than three coefficients, untrimmed operands, another modulus or field) takes the
    # reference's own path.
    XF = binding.ExtensionField
    orig_mul, orig_add, orig_sub, orig_neg = XF.multiply, XF.add, XF.subtract, XF.negate
    standard = {}

    def is_standard(xf):
        hit = standard.get(id(xf))
        if hit is None or hit[0] is not xf:
            try:
                m = xf.modulus.coefficients
                q = m[0].field.p
                ok = [c.value for c in m] == [1, q - 1, 0, 1] and all(c.field.p == q for c in m)
            except (AttributeError, IndexError):
                ok, q = False, None
            hit = standard[id(xf)] = (xf, ok, q)
        return hit[1], hit[2]

    def operands(xf, left, right):
        """(coefficient lists, their values, p) when both are trimmed elements of at most three coefficients over
        base fields of the modulus' characteristic; None otherwise"""
        ok, q = is_standard(xf)
        if not ok or type(left) is not X or type(right) is not X:
            return None
        L, R = left.polynomial.coefficients, right.polynomial.coefficients
        if len(L) > 3 or len(R) > 3:
            return None
        try:
            lv, rv = [c.value for c in L], [c.value for c in R]
            if any(c.field.p != q for c in L) or any(c.field.p != q for c in R):
                return None
        except AttributeError:
            return None
        if (lv and not 0 < lv[-1] < q) or (rv and not 0 < rv[-1] < q):
            return None
        return L, R, lv, rv, q

    def element(xf, co):
        poly = Pn.__new__(Pn)
        poly.__dict__ = {"coefficients": co}
        x = X.__new__(X)
        x.__dict__ = {"polynomial": poly, "field": xf}
        return x

    def mk(v, f):
        o = Bcls.__new__(Bcls)
        o.__dict__ = {"value": v, "field": f}
        return o

    def xmul(self, left, right):
        ops = operands(self, left, right)
        if ops is None:
            return orig_mul(self, left, right)
        L, R, lv, rv, q = ops
        a, b = len(lv), len(rv)
        if a == 0 or b == 0:
            return element(self, [])
        n = a + b - 1
        c, touched = [0] * n, [False] * n
        for i, x in enumerate(lv):
            if x == 0:
                continue  # code/univariate.py:46-47
            for j, y in enumerate(rv):
                c[i + j] += x * y
                touched[i + j] = True
        f = L[0].field
        if n <= 3:  # no reduction: the product polynomial itself is the remainder (code/univariate.py:93-94)
            zero = None
            co = []
            for k in range(n):
                if touched[k]:
                    co.append(mk(c[k] % q, f))
                else:
                    if zero is None:
                        zero = mk(0, f)
                    co.append(zero)
            return element(self, co)
        c3, c4 = c[3], (c[4] if n == 5 else 0)  # X^3 = X - 1, X^4 = X^2 - X
        r = [(c[0] - c3) % q, (c[1] + c3 - c4) % q, (c[2] + c4) % q]
        while r and r[-1] == 0:
            r.pop()
        return element(self, [mk(v, f) for v in r])

    def xadd(self, left, right):
        ops = operands(self, left, right)
        if ops is None:
            return orig_add(self, left, right)
        L, R, lv, rv, q = ops
        if not lv:
            return element(self, list(R))  # code/univariate.py:24-25: the other operand's coefficient objects
        if not rv:
            return element(self, list(L))
        n = max(len(lv), len(rv))
        r = [((lv[k] if k < len(lv) else 0) + (rv[k] if k < len(rv) else 0)) % q for k in range(n)]
        while r and r[-1] == 0:
            r.pop()
        f = L[0].field
        return element(self, [mk(v, f) for v in r])

    def xsub(self, left, right):
        ops = operands(self, left, right)
        if ops is None:
            return orig_sub(self, left, right)
        L, R, lv, rv, q = ops
        if not lv:
            return element(self, [mk((q - c.value) % q, c.field) for c in R])  # the negations, each in its own field
        if not rv:
            return element(self, list(L))
        n = max(len(lv), len(rv))
        r = [((lv[k] if k < len(lv) else 0) - (rv[k] if k < len(rv) else 0)) % q for k in range(n)]
        while r and r[-1] == 0:
            r.pop()
        f = L[0].field
        return element(self, [mk(v, f) for v in r])

    def xneg(self, operand):
        ops = operands(self, operand, operand)
        if ops is None:
            return orig_neg(self, operand)
        L, _, lv, _, q = ops
        return element(self, [mk((q - c.value) % q, c.field) for c in L])
    set_attr(XF, "multiply", xmul)
    set_attr(XF, "add", xadd)
    set_attr(XF, "subtract", xsub)
    set_attr(XF, "negate", xneg)

    # Polynomial.degree (code/univariate.py:8-18): the index of the last non-zero coefficient, found by building a
    # zero element, a list of n copies of it, comparing the lists element by element and then scanning all n
    # coefficients -- and every ExtensionFieldElement constructor calls it (code/extension_field.py:7-8): 2.7 M calls
    # in a proof on a 2^20 domain, a third of its host time.  Same integer, found from the top; coefficients that are
    # not base-field elements (no `.value`) take the reference's own path.
    orig_degree = Polynomial.degree

    def degree(self):
        co = self.coefficients
        try:
            for i in range(len(co) - 1, -1, -1):
                if co[i].value != 0:
                    return i
        except AttributeError:
            return orig_degree(self)
        return -1
    set_attr(Polynomial, "degree", degree)

    # -- 4. next row: quotient codewords of the AIR ---------------------------
    # DEBUG keeps the reference's own loops (they print and assert degree bounds on the way).
    if quotients:
        table = importlib.import_module("table")
        pa = importlib.import_module("permutation_argument")
        mv = importlib.import_module("multivariate")
        T = table.Table
        orig_b, orig_t, orig_e = T.boundary_quotients, T.transition_quotients, T.terminal_quotients
        orig_q = pa.PermutationArgument.quotient
        debug = lambda: os.environ.get("DEBUG") is not None  # noqa: E731
        set_attr(T, "boundary_quotients", lambda self, fri_domain, codewords, challenges:
                 orig_b(self, fri_domain, codewords, challenges) if debug()
                 else glue.table_boundary_quotients(self, fri_domain, codewords, challenges))
        set_attr(T, "transition_quotients", lambda self, domain, codewords, challenges:
                 orig_t(self, domain, codewords, challenges) if debug()
                 else glue.table_transition_quotients(self, domain, codewords, challenges))
        set_attr(T, "terminal_quotients", lambda self, domain, codewords, challenges, terminals:
                 orig_e(self, domain, codewords, challenges, terminals) if debug()
                 else glue.table_terminal_quotients(self, domain, codewords, challenges, terminals))
        set_attr(pa.PermutationArgument, "quotient", lambda self, fri_domain:
                 orig_q(self, fri_domain) if debug() else glue.permutation_quotient(self, fri_domain, mv.MPolynomial))

    # -- 5. next row: SaltedMerkle (two of the three trees of prove()) -----------
    if salted:
        sm = importlib.import_module("salted_merkle")
        # `urandom` is looked up in the module at call time: tests patch salted_merkle.urandom
        set_attr(sm.SaltedMerkle, "__init__",
                 lambda self, data_array: glue.salted_merkle_build(self, data_array, lambda k: sm.urandom(k)))

    # -- 5b. next row: interpolate_columns + lde/ldex, all columns of a table at once ---
    if lde:
        table = importlib.import_module("table")
        T = table.Table
        draw = lambda k: table.os.urandom(k)  # noqa: E731  (looked up at call time: tests seed os.urandom)
        set_attr(T, "interpolate_columns", lambda self, omega, omega_order, column_indices:
                 glue.table_interpolate_columns(self, omega, omega_order, column_indices, draw))

        def lde_(self, domain):
            self.codewords = glue.table_lde(self, domain, draw)
            return self.codewords

        def ldex_(self, domain, xfield):
            codewords = glue.table_lde(self, domain, draw, xfield=xfield)
            self.codewords += codewords
            return codewords
        set_attr(T, "lde", lde_)
        set_attr(T, "ldex", ldex_)

    # -- 5c. Table.extend ends in `self.codewords = [[xfield.lift(c) for c in cdwd] for cdwd in self.codewords]`
    # (e.g. code/io_table.py:106-107): 2.1 M lift calls in a Hello-World proof for lists only the quotient kernels
    # read.  With device views as codewords the statement runs over an empty list and the glue lifts the views.
    if lazy:
        from .glue import DeviceCodeword
        XField = binding.ExtensionField

        def lifting(orig):
            def extend(self, *args, **kwargs):
                codewords = getattr(self, "codewords", None)
                if not glue._lazy or type(codewords) is not list or \
                        not any(type(c) is DeviceCodeword for c in codewords):
                    return orig(self, *args, **kwargs)
                self.codewords = []
                try:
                    result = orig(self, *args, **kwargs)
                    as_expected = self.codewords == [] and type(self.field) is XField
                finally:
                    self.codewords = codewords
                if not as_expected:  # ruled out by the source check at install time
                    raise RuntimeError("Table.extend did not treat its codewords as this reference version does")
                self.codewords = glue.lift_codewords(self.field, codewords)
                return result
            extend.__wrapped__ = orig
            return extend
        for mod_name, cls_name, meth in (("processor_table", "ProcessorTable", "extend"),
                                         ("instruction_table", "InstructionTable", "extend"),
                                         ("memory_table", "MemoryTable", "extend"),
                                         ("io_table", "IOTable", "extend_iotable")):
            try:
                cls = getattr(importlib.import_module(mod_name), cls_name)
                orig = cls.__dict__[meth]
                src = inspect.getsource(orig)
                if src.count(_LIFT_STATEMENT) != 1 or src.count("codewords") != 2:
                    raise LookupError("%s.%s has an unexpected shape" % (cls_name, meth))
                set_attr(cls, meth, lifting(orig))
            except (ImportError, KeyError, LookupError, OSError, TypeError) as e:
                lazy = False  # without the wrapper a device view would be lifted element by element: no gain
                warnings.warn("codewords stay host lists inside prove(): %s" % e)
                break

    # -- 6. next row: the nonlinear combination inside BrainfuckStark.prove -------
    if combination:
        try:
            bs = importlib.import_module("brainfuck_stark")
            prove = bs.BrainfuckStark.prove
            code = compile(guarded_prove_source(prove), prove.__code__.co_filename, "exec")
            scope = {}
            exec(code, bs.__dict__, scope)
            saved["globals"].append((bs, _COMB_HOOK, bs.__dict__.get(_COMB_HOOK, _MISSING)))
            # DEBUG keeps the reference's own block (it prints and asserts degree bounds on the way)
            bs.__dict__[_COMB_HOOK] = lambda *args: (None if os.environ.get("DEBUG") is not None
                                                     else glue.combination_codeword(*args))
            saved["globals"].append((bs, _ROWS_HOOK, bs.__dict__.get(_ROWS_HOOK, _MISSING)))
            bs.__dict__[_ROWS_HOOK] = glue.rows_of
            guarded = scope["prove"]
            lazy_views = lazy

            def prove(self, *args, **kwargs):
                # codewords stay readable on the device for the length of one proof; the cyclic collector is paused:
                # a proof allocates tens of millions of acyclic element objects and every generation-0 overflow
                # would rescan them (marshal.bulk_allocation)
                with glue.keep_planes(lazy=lazy_views and os.environ.get("DEBUG") is None), bulk_allocation():
                    return guarded(self, *args, **kwargs)
            prove.__wrapped__ = guarded
            prove.__doc__ = guarded.__doc__
            set_attr(bs.BrainfuckStark, "prove", prove)
        except (ImportError, LookupError, OSError, SyntaxError) as e:
            warnings.warn("nonlinear combination stays on the host: %s" % e)

    _state = {"glue": glue, "saved": saved, "mods": mods}
    return glue


class _Missing:
    pass


_MISSING = _Missing()


def uninstall():
    """restore the reference's own functions and methods"""
    global _state
    if _state is None:
        return
    saved = _state["saved"]
    for mod, name, orig in saved["globals"]:
        if orig is _MISSING:
            mod.__dict__.pop(name, None)
        else:
            mod.__dict__[name] = orig
    for obj, name, orig in reversed(saved["attrs"]):
        if orig is _MISSING:
            delattr(obj, name)
        else:
            setattr(obj, name, orig)
    _state = None
