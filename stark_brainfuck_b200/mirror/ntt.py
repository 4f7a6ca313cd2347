"""Host-side mirror of the reference's `ntt` module (code/ntt.py).  ntt / intt /
fast_coset_evaluate / fast_coset_interpolate run on the device; the remaining fast
polynomial routines are the reference's divide-and-conquer compositions of those."""
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


def _check_root(primitive_root, root_order):
    one = primitive_root.field.one()
    assert primitive_root ^ root_order == one, "supplied root does not have supplied order"
    assert primitive_root ^ (root_order // 2) != one, "supplied root is not primitive root of supplied order"


def fast_multiply(lhs, rhs, primitive_root, root_order):
    """code/ntt.py:45-79: pointwise product of two forward transforms"""
    _check_root(primitive_root, root_order)
    if lhs.is_zero() or rhs.is_zero():
        return Polynomial([])
    field = lhs.coefficients[0].field
    degree = lhs.degree() + rhs.degree()
    if degree < 8:
        return lhs * rhs
    root, order = primitive_root, root_order
    while degree < order // 2:
        root, order = root ^ 2, order // 2
    a = lhs.coefficients[:lhs.degree() + 1]
    b = rhs.coefficients[:rhs.degree() + 1]
    a = a + [field.zero()] * (order - len(a))
    b = b + [field.zero()] * (order - len(b))
    prod = [x * y for x, y in zip(ntt(root, a), ntt(root, b))]
    return Polynomial(intt(root, prod)[:degree + 1])


def batch_inverse(array):
    """Montgomery's trick (code/ntt.py:177-188)"""
    assert all(not a.is_zero() for a in array), "batch inverse does not work when input contains a zero"
    prefix = list(array)
    for i in range(1, len(array)):
        prefix[i] = prefix[i - 1] * array[i]
    acc = prefix[-1].inverse()
    for i in range(len(array) - 1, 0, -1):
        prefix[i] = acc * prefix[i - 1]
        acc = acc * array[i]
    prefix[0] = acc
    return prefix
