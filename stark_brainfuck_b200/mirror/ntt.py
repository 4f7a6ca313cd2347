"""Host-side mirror of the hot-path names of the reference's `ntt` module (code/ntt.py:4-42, :164-174): ntt, intt,
fast_coset_evaluate, fast_coset_interpolate, all on the device.  The reference's divide-and-conquer helpers built on
top of them (fast_multiply, fast_zerofier, ...) are not part of the path and are not mirrored; under the drop-in the
reference's own versions ride on the patched ntt / intt."""
from .hostmodel import *  # noqa: F401,F403


def _g():
    from . import glue
    return glue()


def ntt(primitive_root, values):
    return _g().ntt(primitive_root, values)


def intt(primitive_root, values):
    return _g().intt(primitive_root, values)


def fast_coset_evaluate(polynomial, offset, generator, order):
    return _g().fast_coset_evaluate(polynomial, offset, generator, order)


def fast_coset_interpolate(offset, generator, values):
    return _g().fast_coset_interpolate(offset, generator, values)
