"""Host-side mirror of the reference's `fri` module (code/fri.py): class Fri with nested
Domain.  Prover-side work (Domain.evaluate/xevaluate/interpolate/xinterpolate, commit, query,
query_last, prove) runs on the device through the glue; verify is the host-side checker."""
from hashlib import blake2b

from .algebra import *  # noqa: F401,F403
from .extension_field import ExtensionField, ExtensionFieldElement  # noqa: F401
from .ip import *  # noqa: F401,F403
from .merkle import *  # noqa: F401,F403
from .ntt import *  # noqa: F401,F403
from .univariate import *  # noqa: F401,F403


def _g():
    from . import glue
    return glue()


class Fri:
    class Domain:
        def __init__(self, offset, omega, length):
            self.offset = offset
            self.omega = omega
            self.length = length

        def __call__(self, index):
            return (self.omega ^ index) * self.offset

        def list(self):
            return [self(i) for i in range(self.length)]

        def evaluate(self, polynomial):
            return _g().domain_evaluate(self, polynomial)

        def xevaluate(self, polynomial, xfield=None):
            return _g().domain_xevaluate(self, polynomial, xfield)

        def interpolate(self, values):
            return _g().domain_interpolate(self, values)

        def xinterpolate(self, values):
            return _g().domain_xinterpolate(self, values)

    def __init__(self, offset, omega, initial_domain_length, expansion_factor, num_colinearity_tests, xfield):
        self.domain = Fri.Domain(offset, omega, initial_domain_length)
        self.field = xfield
        self.expansion_factor = expansion_factor
        self.num_colinearity_tests = num_colinearity_tests
        assert self.num_rounds() >= 1, "cannot do FRI with less than one round"

    def num_rounds(self):
        length, rounds = self.domain.length, 0
        while length > self.expansion_factor:
            length //= 2
            rounds += 1
        return rounds

    def sample_index(byte_array, size):
        return int.from_bytes(bytes(byte_array), "big") % size

    def sample_indices(self, seed, size, reduced_size, number):
        """code/fri.py:68-86: blake2b(seed | counter NUL bytes) mod size, distinct modulo reduced_size"""
        assert number <= reduced_size, \
            f"cannot sample more indices than available in last codeword; requested: {number}, available: {reduced_size}"
        assert number <= 2 * reduced_size, "not enough entropy in indices wrt last codeword"
        indices, seen, counter = [], set(), 0
        while len(indices) < number:
            index = Fri.sample_index(blake2b(seed + bytes(counter)).digest(), size)
            counter += 1
            if index % reduced_size not in seen:
                seen.add(index % reduced_size)
                indices.append(index)
        return indices

    def eval_domain(self):
        return [self.domain(i) for i in range(self.domain.length)]

    def commit(self, codeword, proof_stream, round_index=0):
        return _g().fri_commit(self, codeword, proof_stream, round_index, Merkle=Merkle)

    def query(self, current_tree, next_tree, c_indices, proof_stream):
        return _g().fri_query(self, current_tree, next_tree, c_indices, proof_stream)

    def query_last(self, current_tree, last_codeword, c_indices, proof_stream):
        return _g().fri_query_last(self, current_tree, last_codeword, c_indices, proof_stream)

    def prove(self, codeword, proof_stream):
        return _g().fri_prove(self, codeword, proof_stream)

    def verify(self, proof_stream, root):
        """code/fri.py:201-319.  Host-side; returns False (after printing why) on rejection."""
        xf = self.field
        omega, offset = xf.lift(self.domain.omega), xf.lift(self.domain.offset)
        rounds = self.num_rounds()
        roots, alphas = [root], []
        for r in range(rounds):
            if r > 0:
                roots.append(proof_stream.pull())
            alphas.append(xf.sample(proof_stream.verifier_fiat_shamir()))
        last_codeword = proof_stream.pull()
        if roots[-1] != Merkle(last_codeword).root():
            print("last codeword is not well formed")
            return False
        max_degree = len(last_codeword) // self.expansion_factor - 1
        last_omega, last_offset = omega, offset
        for _ in range(rounds - 1):
            last_omega, last_offset = last_omega ^ 2, last_offset ^ 2
        assert last_omega.inverse() == last_omega ^ (len(last_codeword) - 1), "omega does not have right order"
        last_domain = [last_offset * (last_omega ^ i) for i in range(len(last_codeword))]
        poly = Polynomial.interpolate_domain(last_domain, last_codeword)
        assert poly.evaluate_domain(last_domain) == last_codeword, "re-evaluated codeword does not match original!"
        if poly.degree() > max_degree:
            return False
        n = self.domain.length
        top = self.sample_indices(proof_stream.verifier_fiat_shamir(), n >> 1, n >> (rounds - 1),
                                  self.num_colinearity_tests)
        s = self.num_colinearity_tests
        for r in range(rounds - 1):
            half = n >> (r + 1)
            c_idx = [i % half for i in top]
            a_idx = list(c_idx)
            b_idx = [i + half for i in a_idx]
            aa, bb, cc = [], [], []
            for k in range(s):
                ay, by, cy = proof_stream.pull()
                aa.append(ay)
                bb.append(by)
                cc.append(cy)
                ax, bx = offset * (omega ^ a_idx[k]), offset * (omega ^ b_idx[k])
                if not test_colinearity([(ax, ay), (bx, by), (alphas[r], cy)]):
                    print("colinearity check failure")
                    return False
            for k in range(s):
                if not Merkle.verify(roots[r], a_idx[k], proof_stream.pull(), aa[k]):
                    print("merkle authentication path verification fails for aa")
                    return False
                if not Merkle.verify(roots[r], b_idx[k], proof_stream.pull(), bb[k]):
                    print("merkle authentication path verification fails for bb")
                    return False
                if r + 1 != rounds - 1:
                    if not Merkle.verify(roots[r + 1], c_idx[k], proof_stream.pull(), cc[k]):
                        print("merkle authentication path verification fails for cc")
                        return False
            if r + 1 == rounds - 1:
                for k in range(s):
                    if cc[k] != last_codeword[c_idx[k]]:
                        print("leafs in last round do not correspond to last codeword")
                        return False
            omega, offset = omega ^ 2, offset ^ 2
        return True
