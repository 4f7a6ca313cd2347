"""Host-side mirror of the reference's `univariate` module (code/univariate.py).  The two
methods on the hot path -- scale (code/univariate.py:168-169) and evaluate_domain (:153-154)
-- go to the engine; everything else is small symbolic arithmetic on the host."""
from .algebra import *  # noqa: F401,F403


class Polynomial:
    __module__ = "univariate"

    def __init__(self, coefficients):
        self.coefficients = [c for c in coefficients]

    def degree(self):
        for i in range(len(self.coefficients) - 1, -1, -1):
            if not self.coefficients[i].is_zero():
                return i
        return -1

    def is_zero(self):
        return self.degree() == -1

    def leading_coefficient(self):
        return self.coefficients[self.degree()]

    def __neg__(self):
        return Polynomial([-c for c in self.coefficients])

    def __add__(self, other):
        if self.degree() == -1:
            return other
        if other.degree() == -1:
            return self
        zero = self.coefficients[0].field.zero()
        n = max(len(self.coefficients), len(other.coefficients))
        out = [zero] * n
        for i, c in enumerate(self.coefficients):
            out[i] = out[i] + c
        for i, c in enumerate(other.coefficients):
            out[i] = out[i] + c
        return Polynomial(out)

    def __sub__(self, other):
        return self + (-other)

    def __mul__(self, other):
        if not self.coefficients or not other.coefficients:
            return Polynomial([])
        zero = self.coefficients[0].field.zero()
        out = [zero] * (len(self.coefficients) + len(other.coefficients) - 1)
        for i, a in enumerate(self.coefficients):
            if a.is_zero():
                continue
            for j, b in enumerate(other.coefficients):
                out[i + j] = out[i + j] + a * b
        return Polynomial(out)

    def divide(numerator, denominator):
        """long division -> (quotient, remainder); None for a zero denominator"""
        dd = denominator.degree()
        if dd == -1:
            return None
        if numerator.degree() < dd:
            return Polynomial([]), numerator
        field = denominator.coefficients[0].field
        rem = Polynomial(numerator.coefficients)
        quo = [field.zero() for _ in range(numerator.degree() - dd + 1)]
        lead = denominator.leading_coefficient()
        for _ in range(len(quo)):
            rd = rem.degree()
            if rd < dd:
                break
            c = rem.leading_coefficient() / lead
            shift = rd - dd
            quo[shift] = c
            rem = rem - Polynomial([field.zero()] * shift + [c]) * denominator
        return Polynomial(quo), rem

    def __truediv__(self, other):
        quo, rem = Polynomial.divide(self, other)
        assert rem.is_zero(), "cannot perform polynomial division because remainder is not zero"
        return quo

    def __floordiv__(self, other):
        return Polynomial.divide(self, other)[0]

    def __mod__(self, other):
        return Polynomial.divide(self, other)[1]

    def __eq__(self, other):
        assert type(self) == type(other), \
            f"type of self {type(self)} must be equal to type of other which is {type(other)}"
        d = self.degree()
        if d != other.degree():
            return False
        return all(self.coefficients[i] == other.coefficients[i] for i in range(d + 1))

    def __neq__(self, other):
        return not self.__eq__(other)

    def __str__(self):
        return "[" + ",".join(str(c) for c in self.coefficients) + "]"

    def interpolate_domain(domain, values):
        """Lagrange interpolation (code/univariate.py:119-135)"""
        assert len(domain) == len(values), \
            "number of elements in domain does not match number of values -- cannot interpolate"
        assert len(domain) > 0, "cannot interpolate between zero points"
        field = domain[0].field
        x = Polynomial([field.zero(), field.one()])
        acc = Polynomial([])
        for i, (xi, yi) in enumerate(zip(domain, values)):
            term = Polynomial([yi])
            for j, xj in enumerate(domain):
                if j != i:
                    term = term * (x - Polynomial([xj])) * Polynomial([(xi - xj).inverse()])
            acc = acc + term
        return acc

    def zerofier_domain(domain):
        field = domain[0].field
        x = Polynomial([field.zero(), field.one()])
        acc = Polynomial([field.one()])
        for d in domain:
            acc = acc * (x - Polynomial([d]))
        return acc

    def evaluate(self, point):
        """running power of the point, not Horner (code/univariate.py:145-151)"""
        xi = point.field.one()
        value = point.field.zero()
        for c in self.coefficients:
            value = value + c * xi
            xi = xi * point
        return value

    def evaluate_domain(self, domain):
        from . import glue
        return glue().poly_evaluate_domain(self, domain)

    def __xor__(self, exponent):
        if self.is_zero():
            return Polynomial([])
        acc = Polynomial([self.coefficients[0].field.one()])
        for bit in bin(exponent)[2:]:
            acc = acc * acc
            if bit == "1":
                acc = acc * self
        return acc

    def scale(self, factor):
        from . import glue
        return glue().poly_scale(self, factor)

    def xgcd(x, y):
        """monic extended Euclid over polynomials -> (a, b, g)  (code/univariate.py:171-187)"""
        one = Polynomial([x.coefficients[0].field.one()])
        zero = Polynomial([x.coefficients[0].field.zero()])
        r0, r1, s0, s1, t0, t1 = x, y, one, zero, zero, one
        while not r1.is_zero():
            q = r0 // r1
            r0, r1 = r1, r0 - q * r1
            s0, s1 = s1, s0 - q * s1
            t0, t1 = t1, t0 - q * t1
        k = r0.coefficients[r0.degree()].inverse()
        return (Polynomial([c * k for c in s0.coefficients]), Polynomial([c * k for c in t0.coefficients]),
                Polynomial([c * k for c in r0.coefficients]))


def test_colinearity(points):
    domain = [p[0] for p in points]
    values = [p[1] for p in points]
    return Polynomial.interpolate_domain(domain, values).degree() == 1
