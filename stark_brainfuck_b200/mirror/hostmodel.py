"""Host-side object model of the standalone front end: the element, field, polynomial and
proof-stream classes that code written against the reference expects to find in its modules
`algebra`, `univariate`, `extension_field` and `ip` (class names, method names, argument
meaning, assertion behaviour and -- because pickle is the reference's wire format -- instance
attribute names and their order).  mirror/__init__.py exposes these classes under those module
names.

Nothing here is on the hot path: vectors of elements go through the engine (glue.py).  The
scalar arithmetic below works on Python ints directly; results are canonical integers in
[0, p) exactly like the reference's `% p` arithmetic (code/algebra.py:89-108), an extension
element is a Polynomial with trailing zeros trimmed (code/extension_field.py:6-9), and every
result carries the LEFT operand's field object (pickle memoises by identity).
"""
import pickle
from hashlib import shake_256

P = (1 << 64) - (1 << 32) + 1
_ROOT_2_32 = 1753635133440165772  # 7^(2^32 - 1): order 2^32  (code/algebra.py:129)


def xgcd(x, y):
    """Bezout coefficients over the integers: (a, b, g) with a*x + b*y == g."""
    prev, cur = (x, 1, 0), (y, 0, 1)
    while cur[0]:
        q = prev[0] // cur[0]
        prev, cur = cur, tuple(u - q * v for u, v in zip(prev, cur))
    return prev[1], prev[2], prev[0]


def _square_and_multiply(base, exponent, unit):
    acc = unit
    for bit in format(exponent, "b"):
        acc = acc * acc
        if bit == "1":
            acc = acc * base
    return acc


# ------------------------------------------------------------------------------- algebra
class BaseFieldElement:
    __module__ = "algebra"

    def __init__(self, value, field):
        self.value = value
        self.field = field

    def _like(self, integer):
        return BaseFieldElement(integer % self.field.p, self.field)

    def __add__(self, right):
        return self._like(self.value + right.value)

    def __sub__(self, right):
        return self._like(self.value - right.value)

    def __neg__(self):
        return self._like(-self.value)

    def __mul__(self, right):
        return self._like(self.value * right.value)

    def inverse(self):
        return self._like(xgcd(self.value, self.field.p)[0])  # 0 -> 0, as in the reference

    def __truediv__(self, right):
        assert not right.is_zero(), "divide by zero"
        return self._like(self.value * xgcd(right.value, self.field.p)[0])

    def __xor__(self, exponent):  # `^` is exponentiation in the reference (code/algebra.py:39-46)
        return BaseFieldElement(pow(self.value, exponent, self.field.p), self.field)

    def __eq__(self, other):
        return self.value == other.value

    def __neq__(self, other):
        return self.value != other.value

    def __hash__(self):
        return self.value

    def __str__(self):
        return str(self.value)

    def __bytes__(self):
        return str(self.value).encode()

    def is_zero(self):
        return self.value == 0

    def has_order_po2(self, order):
        assert order & (order - 1) == 0
        if order == 1 and self.value == 1:
            return True
        return pow(self.value, order, self.field.p) == 1 and pow(self.value, order // 2, self.field.p) != 1


class BaseField:
    __module__ = "algebra"

    def __init__(self, p):
        self.p = p

    def main():
        return BaseField(P)

    def __call__(self, integer):
        return BaseFieldElement(integer % self.p, self)

    def zero(self):
        return BaseFieldElement(0, self)

    def one(self):
        return BaseFieldElement(1, self)

    def lift(self, bfe):
        return bfe

    # field-level spellings of the element operators; the result belongs to THIS field object
    def add(self, left, right):
        return self(left.value + right.value)

    def subtract(self, left, right):
        return self(left.value - right.value)

    def negate(self, operand):
        return self(-operand.value)

    def multiply(self, left, right):
        return self(left.value * right.value)

    def inverse(self, operand):
        return self(xgcd(operand.value, self.p)[0])

    def divide(self, left, right):
        assert not right.is_zero(), "divide by zero"
        return self(left.value * xgcd(right.value, self.p)[0])

    def generator(self):
        assert self.p == P, "Do not know generator for other fields beyond 2^64 - 2^32 + 1"
        return BaseFieldElement(7, self)

    def primitive_nth_root(self, n):
        assert self.p == P, "Unknown field, can't return root of unity."
        assert n <= 1 << 32 and n & (n - 1) == 0, \
            "Field does not have nth root of unity where n > 2^32 or not power of two."
        halvings = 32 - (n.bit_length() - 1)
        return BaseFieldElement(pow(_ROOT_2_32, 1 << halvings, P), self)

    def sample(self, byte_array):
        return self(int.from_bytes(bytes(byte_array), "big"))


# ---------------------------------------------------------------------------- univariate
class Polynomial:
    __module__ = "univariate"

    def __init__(self, coefficients):
        self.coefficients = list(coefficients)

    # -- structure
    def degree(self):
        d = len(self.coefficients) - 1
        while d >= 0 and self.coefficients[d].is_zero():
            d -= 1
        return d

    def is_zero(self):
        return self.degree() < 0

    def leading_coefficient(self):
        return self.coefficients[self.degree()]

    def __str__(self):
        return "[%s]" % ",".join(map(str, self.coefficients))

    def __eq__(self, other):
        assert type(self) == type(other), \
            f"type of self {type(self)} must be equal to type of other which is {type(other)}"
        d = self.degree()
        return d == other.degree() and self.coefficients[:d + 1] == other.coefficients[:d + 1]

    def __neq__(self, other):
        return not self == other

    # -- ring operations (dense, schoolbook: only small polynomials stay on the host)
    def __neg__(self):
        return Polynomial(-c for c in self.coefficients)

    def __add__(self, other):
        if self.is_zero():
            return other
        if other.is_zero():
            return self
        zero = self.coefficients[0].field.zero()
        total = []
        for i in range(max(len(self.coefficients), len(other.coefficients))):
            acc = zero  # sums start from the left operand's zero, so they inherit its field object
            for operand in (self.coefficients, other.coefficients):
                if i < len(operand):
                    acc = acc + operand[i]
            total.append(acc)
        return Polynomial(total)

    def __sub__(self, other):
        return self + -other

    def __mul__(self, other):
        a, b = self.coefficients, other.coefficients
        if not a or not b:
            return Polynomial([])
        zero = a[0].field.zero()
        prod = [zero] * (len(a) + len(b) - 1)
        for i, ai in enumerate(a):
            if not ai.is_zero():
                for j, bj in enumerate(b):
                    prod[i + j] = prod[i + j] + ai * bj
        return Polynomial(prod)

    def divide(numerator, denominator):
        """(quotient, remainder) by long division, or None for the zero denominator."""
        dd = denominator.degree()
        if dd < 0:
            return None
        if numerator.degree() < dd:
            return Polynomial([]), numerator
        field = denominator.coefficients[0].field
        lead = denominator.leading_coefficient()
        rem = Polynomial(numerator.coefficients)
        quo = [field.zero() for _ in range(numerator.degree() - dd + 1)]
        while rem.degree() >= dd:
            shift = rem.degree() - dd
            quo[shift] = rem.leading_coefficient() / lead
            rem = rem - Polynomial([field.zero()] * shift + [quo[shift]]) * denominator
        return Polynomial(quo), rem

    def __truediv__(self, other):
        quo, rem = Polynomial.divide(self, other)
        assert rem.is_zero(), "cannot perform polynomial division because remainder is not zero"
        return quo

    def __floordiv__(self, other):
        return Polynomial.divide(self, other)[0]

    def __mod__(self, other):
        return Polynomial.divide(self, other)[1]

    def __xor__(self, exponent):
        if self.is_zero():
            return Polynomial([])
        return _square_and_multiply(self, exponent, Polynomial([self.coefficients[0].field.one()]))

    def xgcd(x, y):
        """Bezout coefficients over polynomials, normalised to a monic gcd: (a, b, g)."""
        field = x.coefficients[0].field
        prev = (x, Polynomial([field.one()]), Polynomial([field.zero()]))
        cur = (y, Polynomial([field.zero()]), Polynomial([field.one()]))
        while not cur[0].is_zero():
            q = prev[0] // cur[0]
            prev, cur = cur, tuple(u - q * v for u, v in zip(prev, cur))
        g, a, b = prev
        k = g.leading_coefficient().inverse()
        return tuple(Polynomial(c * k for c in f.coefficients) for f in (a, b, g))

    # -- evaluation / interpolation
    def evaluate(self, point):
        """sum of c_k * x^k with a running power of the point (code/univariate.py:145-151)"""
        power, total = point.field.one(), point.field.zero()
        for c in self.coefficients:
            total = total + c * power
            power = power * point
        return total

    def evaluate_domain(self, domain):  # hot path: engine (code/univariate.py:153-154)
        from . import glue
        return glue().poly_evaluate_domain(self, domain)

    def scale(self, factor):  # hot path: engine (code/univariate.py:168-169)
        from . import glue
        return glue().poly_scale(self, factor)

    def interpolate_domain(domain, values):
        """Lagrange interpolation through (domain[i], values[i])."""
        assert len(domain) == len(values), \
            "number of elements in domain does not match number of values -- cannot interpolate"
        assert len(domain) > 0, "cannot interpolate between zero points"
        field = domain[0].field
        x = Polynomial([field.zero(), field.one()])
        total = Polynomial([])
        for i, (xi, yi) in enumerate(zip(domain, values)):
            basis = Polynomial([yi])
            for j, xj in enumerate(domain):
                if j != i:
                    basis = basis * (x - Polynomial([xj])) * Polynomial([(xi - xj).inverse()])
            total = total + basis
        return total

    def zerofier_domain(domain):
        field = domain[0].field
        x = Polynomial([field.zero(), field.one()])
        product = Polynomial([field.one()])
        for point in domain:
            product = product * (x - Polynomial([point]))
        return product


def test_colinearity(points):
    xs, ys = zip(*points)
    return Polynomial.interpolate_domain(list(xs), list(ys)).degree() == 1


# ----------------------------------------------------------------------- extension_field
class ExtensionFieldElement:
    __module__ = "extension_field"

    def __init__(self, polynomial, field):
        self.polynomial = Polynomial(polynomial.coefficients[:polynomial.degree() + 1])  # trimmed
        self.field = field

    def _reduced(self, polynomial):
        return ExtensionFieldElement(polynomial % self.field.modulus, self.field)

    def __add__(self, right):
        return ExtensionFieldElement(self.polynomial + right.polynomial, self.field)

    def __sub__(self, right):
        return ExtensionFieldElement(self.polynomial - right.polynomial, self.field)

    def __neg__(self):
        return ExtensionFieldElement(-self.polynomial, self.field)

    def __mul__(self, right):
        return self._reduced(self.polynomial * right.polynomial)

    def inverse(self):
        a, b, g = Polynomial.xgcd(self.polynomial, self.field.modulus)
        assert a * self.polynomial + b * self.field.modulus == g, "bezout relation fails"
        return self._reduced(a)

    def __truediv__(self, right):
        assert not right.is_zero(), "divide by zero"
        return self._reduced(self.polynomial * Polynomial.xgcd(right.polynomial, self.field.modulus)[0])

    def __xor__(self, exponent):
        return _square_and_multiply(self, exponent, self.field.one())

    def __eq__(self, other):
        return self.polynomial == other.polynomial

    def __neq__(self, other):
        return self.polynomial != other.polynomial

    def __str__(self):
        return str(self.polynomial)

    def is_zero(self):
        return self.polynomial.is_zero()


class ExtensionField:
    __module__ = "extension_field"

    def __init__(self, modulus):
        self.modulus = modulus

    def main():
        base = BaseField(P)
        one = BaseFieldElement(1, base)
        # X^3 - X + 1; both `one` entries are the same object (code/extension_field.py:94-97)
        return ExtensionField(Polynomial([one, BaseFieldElement(P - 1, base), base.zero(), one]))

    def _base(self):
        return self.modulus.coefficients[0].field

    def __call__(self, integer):
        return ExtensionFieldElement(Polynomial([BaseFieldElement(integer, self._base())]), self)

    def zero(self):
        return ExtensionFieldElement(Polynomial([]), self)

    def one(self):
        return ExtensionFieldElement(Polynomial([self._base().one()]), self)

    def lift(self, base_field_element):
        if type(base_field_element) == ExtensionFieldElement:
            return base_field_element
        return ExtensionFieldElement(Polynomial([base_field_element]), self)

    def sample(self, byte_array):
        """one coefficient per big-endian chunk of len/3 bytes (code/extension_field.py:100-111)"""
        parts = self.modulus.degree()
        width = len(byte_array) // parts
        base = self._base()
        return ExtensionFieldElement(
            Polynomial(base.sample(byte_array[k * width:(k + 1) * width]) for k in range(parts)), self)

    # field-level spellings; the result belongs to THIS field object
    def add(self, left, right):
        return ExtensionFieldElement(left.polynomial + right.polynomial, self)

    def subtract(self, left, right):
        return ExtensionFieldElement(left.polynomial - right.polynomial, self)

    def negate(self, operand):
        return ExtensionFieldElement(-operand.polynomial, self)

    def multiply(self, left, right):
        return ExtensionFieldElement(left.polynomial * right.polynomial % self.modulus, self)

    def inverse(self, operand):
        return ExtensionFieldElement(operand.inverse().polynomial, self)

    def divide(self, left, right):
        return ExtensionFieldElement((left / right).polynomial, self)


# ------------------------------------------------------------------------------------- ip
class ProofStream:
    """The transcript: a list of pushed objects whose pickle is both the proof and the
    Fiat-Shamir input (code/ip.py)."""
    __module__ = "ip"

    def __init__(self):
        self.objects = []
        self.read_index = 0

    def push(self, obj):
        self.objects.append(obj)

    def pull(self):
        assert self.read_index < len(self.objects), "ProofStream: cannot pull object; queue empty."
        obj = self.objects[self.read_index]
        self.read_index += 1
        return obj

    def serialize(self):
        return pickle.dumps(self.objects)

    def deserialize(self, bb):
        stream = ProofStream()
        stream.objects = pickle.loads(bb)
        return stream

    def _challenge(self, objects, num_bytes):
        return shake_256(pickle.dumps(objects)).digest(num_bytes)

    def prover_fiat_shamir(self, num_bytes=32):
        return self._challenge(self.objects, num_bytes)

    def verifier_fiat_shamir(self, num_bytes=32):
        return self._challenge(self.objects[:self.read_index], num_bytes)
