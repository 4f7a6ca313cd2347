"""Host-side mirror of the reference's `algebra` module (code/algebra.py): same class and
method names, same argument meaning, same canonical-integer semantics.  Written for the
standalone (reference-less) front end; the per-element arithmetic here is NOT the hot path --
vectors go through the engine.  Pickles byte-identically to the reference's classes when the
mirror is registered under the bare module names (mirror.register())."""

P = (1 << 64) - (1 << 32) + 1


def xgcd(x, y):
    """extended Euclid: returns (a, b, g) with a*x + b*y == g  (code/algebra.py:1-12)"""
    r0, r1, s0, s1, t0, t1 = x, y, 1, 0, 0, 1
    while r1:
        q = r0 // r1
        r0, r1 = r1, r0 - q * r1
        s0, s1 = s1, s0 - q * s1
        t0, t1 = t1, t0 - q * t1
    return s0, t0, r0


class BaseFieldElement:
    __module__ = "algebra"

    def __init__(self, value, field):
        self.value = value
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
        """modular exponentiation, like the reference's `^` (code/algebra.py:39-46)"""
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
        return str(self).encode()

    def is_zero(self):
        return self.value == 0

    def has_order_po2(self, order):
        assert order & (order - 1) == 0
        if self.value == 1 and order == 1:
            return True
        p = self.field.p
        return pow(self.value, order, p) == 1 and pow(self.value, order // 2, p) != 1


class BaseField:
    __module__ = "algebra"

    def __init__(self, p):
        self.p = p

    def lift(self, bfe):
        return bfe

    def zero(self):
        return BaseFieldElement(0, self)

    def one(self):
        return BaseFieldElement(1, self)

    def add(self, left, right):
        return BaseFieldElement((left.value + right.value) % self.p, self)

    def subtract(self, left, right):
        return BaseFieldElement((left.value - right.value) % self.p, self)

    def negate(self, operand):
        return BaseFieldElement(-operand.value % self.p, self)

    def multiply(self, left, right):
        return BaseFieldElement(left.value * right.value % self.p, self)

    def inverse(self, operand):
        a, _, _ = xgcd(operand.value, self.p)  # inverse(0) == 0 like code/algebra.py:101-103
        return BaseFieldElement(a % self.p, self)

    def divide(self, left, right):
        assert not right.is_zero(), "divide by zero"
        a, _, _ = xgcd(right.value, self.p)
        return BaseFieldElement(left.value * a % self.p, self)

    def main():
        return BaseField(P)

    def generator(self):
        assert self.p == P, "Do not know generator for other fields beyond 2^64 - 2^32 + 1"
        return BaseFieldElement(7, self)

    def primitive_nth_root(self, n):
        assert self.p == P, "Unknown field, can't return root of unity."
        assert n <= 1 << 32 and (n & (n - 1)) == 0, \
            "Field does not have nth root of unity where n > 2^32 or not power of two."
        # 7^(2^32 - 1) has order 2^32; square down to order n  (code/algebra.py:122-136)
        root, order = 1753635133440165772, 1 << 32
        while order != n:
            root = root * root % P
            order >>= 1
        return BaseFieldElement(root, self)

    def sample(self, byte_array):
        return BaseFieldElement(int.from_bytes(bytes(byte_array), "big") % self.p, self)

    def __call__(self, integer):
        return BaseFieldElement(integer % self.p, self)
