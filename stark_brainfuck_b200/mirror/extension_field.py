"""Host-side mirror of the reference's `extension_field` module (code/extension_field.py):
F_p[X]/(X^3 - X + 1) with elements stored as trimmed Polynomials."""
from .algebra import *  # noqa: F401,F403
from .univariate import *  # noqa: F401,F403


class ExtensionFieldElement:
    __module__ = "extension_field"

    def __init__(self, polynomial, field):
        # trailing zero coefficients are dropped (code/extension_field.py:6-9)
        self.polynomial = Polynomial(polynomial.coefficients[:polynomial.degree() + 1])
        self.field = field

    def __add__(self, right):
        return self.field.add(self, right)

    def __sub__(self, right):
        return self.field.subtract(self, right)

    def __mul__(self, right):
        return self.field.multiply(self, right)

    def __truediv__(self, right):
        return self.field.divide(self, right)

    def __neg__(self):
        return self.field.negate(self)

    def inverse(self):
        return self.field.inverse(self)

    def __xor__(self, exponent):
        acc = self.field.one()
        for bit in bin(exponent)[2:]:
            acc = acc * acc
            if bit == "1":
                acc = acc * self
        return acc

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

    def _bf(self):
        return self.modulus.coefficients[0].field

    def zero(self):
        return ExtensionFieldElement(Polynomial([]), self)

    def one(self):
        return ExtensionFieldElement(Polynomial([self._bf().one()]), self)

    def add(self, left, right):
        return ExtensionFieldElement(left.polynomial + right.polynomial, self)

    def subtract(self, left, right):
        return ExtensionFieldElement(left.polynomial - right.polynomial, self)

    def negate(self, operand):
        return ExtensionFieldElement(-operand.polynomial, self)

    def multiply(self, left, right):
        return ExtensionFieldElement((left.polynomial * right.polynomial) % self.modulus, self)

    def inverse(self, operand):
        a, b, g = Polynomial.xgcd(operand.polynomial, self.modulus)
        assert a * operand.polynomial + b * self.modulus == g, "bezout relation fails"
        return ExtensionFieldElement(a % self.modulus, self)

    def divide(self, left, right):
        assert not right.is_zero(), "divide by zero"
        a, _, _ = Polynomial.xgcd(right.polynomial, self.modulus)
        return ExtensionFieldElement(left.polynomial * a % self.modulus, self)

    def main():
        field = BaseField(P)
        one = BaseFieldElement(1, field)
        # X^3 - X + 1; the two `one` entries are the same object (code/extension_field.py:94-97)
        return ExtensionField(Polynomial([one, BaseFieldElement(P - 1, field), field.zero(), one]))

    def sample(self, byte_array):
        """three big-endian chunks of len/3 bytes (code/extension_field.py:100-111)"""
        deg = self.modulus.degree()
        chunk = len(byte_array) // deg
        bf = self._bf()
        return ExtensionFieldElement(
            Polynomial([bf.sample(byte_array[i * chunk:(i + 1) * chunk]) for i in range(deg)]), self)

    def lift(self, base_field_element):
        if type(base_field_element) == ExtensionFieldElement:
            return base_field_element
        return ExtensionFieldElement(Polynomial([base_field_element]), self)

    def __call__(self, integer):
        return ExtensionFieldElement(Polynomial([BaseFieldElement(integer, self._bf())]), self)
