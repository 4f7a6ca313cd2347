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
     Merkle.__init__/open  (Merkle.root/verify and Fri.verify stay the reference's).
uninstall() restores every original (needed to time the CPU reference in the same process).
"""
import importlib
import os
import sys

from .glue import Glue
from .marshal import Binding

_state = None

_NTT_FUNCS = ("ntt", "intt", "fast_coset_evaluate", "fast_coset_interpolate")


def installed():
    return _state is not None


def current_glue():
    return _state["glue"] if _state else None


def install(reference_dir=None, engine=None, quotients=True, salted=True):
    """Patch the reference modules importable from `reference_dir` (or already on sys.path).
    `quotients=True` also moves the quotient-codeword loops of table.py / permutation_argument.py
    (93 % of prove()) to the device, `salted=True` the trees of salted_merkle.py.
    Returns the Glue in use."""
    global _state
    if _state is not None:
        return _state["glue"]
    if reference_dir is not None and reference_dir not in sys.path:
        sys.path.insert(0, reference_dir)
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
        mod.__dict__[name] = orig
    for obj, name, orig in reversed(saved["attrs"]):
        if orig is _MISSING:
            delattr(obj, name)
        else:
            setattr(obj, name, orig)
    _state = None
