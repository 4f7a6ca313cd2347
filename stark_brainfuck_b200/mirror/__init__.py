"""Standalone host-side mirror of the reference's interface for the hot path.

Where the reference checkout is available, use stark_brainfuck_b200.dropin.install() and keep
calling the reference's own modules.  Where it is not (the GPU box), this package offers the
same module / class / function names -- algebra, univariate, extension_field, ntt, merkle,
salted_merkle, ip, fri -- so code and tests written against the reference read the same.

Pickle is the reference's wire format and records module names, so byte-identical Merkle
leaves and transcripts need these classes to be importable under the reference's BARE module
names: register() installs them in sys.modules (refusing to shadow a loaded reference),
unregister() removes them.
"""
import sys
import types

from . import fri, hostmodel, merkle, ntt, salted_merkle
from ..glue import Glue
from ..marshal import Binding


def _module(name, exported):
    """a module object that exposes `exported` names of hostmodel (what `from X import *` of the
    reference's module X would bring in, including what X itself star-imported)"""
    mod = types.ModuleType(name, "stark_brainfuck_b200.mirror view of the reference's `%s` module" % name)
    mod.__file__ = hostmodel.__file__
    for n in exported:
        setattr(mod, n, getattr(hostmodel, n))
    return mod


_ALGEBRA = ("P", "xgcd", "BaseFieldElement", "BaseField")
_UNIVARIATE = _ALGEBRA + ("Polynomial", "test_colinearity")
_EXTENSION = _UNIVARIATE + ("ExtensionFieldElement", "ExtensionField")
algebra = _module("algebra", _ALGEBRA)
univariate = _module("univariate", _UNIVARIATE)
extension_field = _module("extension_field", _EXTENSION)
ip = _module("ip", ("ProofStream", "pickle", "shake_256"))

NAMES = ("algebra", "univariate", "extension_field", "ntt", "merkle", "salted_merkle", "ip", "fri")
_MODS = {"algebra": algebra, "univariate": univariate, "extension_field": extension_field, "ntt": ntt,
         "merkle": merkle, "salted_merkle": salted_merkle, "ip": ip, "fri": fri}

binding = Binding.from_modules(algebra, univariate, extension_field)
field = algebra.BaseField.main()
xfield = extension_field.ExtensionField.main()

_glue = None


def glue():
    """the Glue all mirror modules call (created on first use; needs the CUDA engine)"""
    global _glue
    if _glue is None:
        register()  # pickle needs the classes under the reference's bare module names
        _glue = Glue(binding)
    return _glue


def set_glue(g):
    global _glue
    _glue = g


def register(force=False):
    for name in NAMES:
        cur = sys.modules.get(name)
        if cur is not None and cur is not _MODS[name] and not force:
            raise RuntimeError("module %r is already loaded from %s; refusing to shadow it" %
                               (name, getattr(cur, "__file__", "?")))
    for name in NAMES:
        sys.modules[name] = _MODS[name]


def unregister():
    for name in NAMES:
        if sys.modules.get(name) is _MODS[name]:
            del sys.modules[name]
