"""Host-side mirror of the reference's `fri` module (code/fri.py): class Fri with nested
Domain.  Prover-side work (Domain.evaluate/xevaluate/interpolate/xinterpolate, commit, query,
query_last, prove) runs on the device through the glue; verify is the host-side checker."""
from hashlib import blake2b

from .hostmodel import *  # noqa: F401,F403  (the classes of algebra / univariate / extension_field / ip)
from .merkle import *  # noqa: F401,F403
from .ntt import *  # noqa: F401,F403


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

    # ---- verifier (host side; the acceptance check of code/fri.py:201-319) -----------------
    def _read_commitments(self, proof_stream, root):
        """Merkle roots of all rounds and the folding challenges re-derived from the transcript."""
        roots, alphas = [root], []
        for r in range(self.num_rounds()):
            if r:
                roots.append(proof_stream.pull())
            alphas.append(self.field.sample(proof_stream.verifier_fiat_shamir()))
        return roots, alphas

    def _last_codeword_ok(self, last_codeword, last_root, omega, offset):
        """the final codeword matches its root and has degree < length / expansion_factor"""
        if last_root != Merkle(last_codeword).root():
            print("last codeword is not well formed")
            return False
        assert omega.inverse() == omega ^ (len(last_codeword) - 1), "omega does not have right order"
        points = [offset * (omega ^ i) for i in range(len(last_codeword))]
        poly = Polynomial.interpolate_domain(points, last_codeword)
        assert poly.evaluate_domain(points) == last_codeword, "re-evaluated codeword does not match original!"
        return poly.degree() <= len(last_codeword) // self.expansion_factor - 1

    def _layer_ok(self, proof_stream, this_root, next_root, alpha, omega, offset, c_idx, half, last_codeword):
        """colinearity of the s opened triples of one layer and their authentication paths;
        next_root is None on the final layer, whose c values are read off the last codeword"""
        triples = [proof_stream.pull() for _ in c_idx]
        for c, (ay, by, cy) in zip(c_idx, triples):
            ax, bx = offset * (omega ^ c), offset * (omega ^ (c + half))
            if not test_colinearity([(ax, ay), (bx, by), (alpha, cy)]):
                print("colinearity check failure")
                return False
        for c, (ay, by, cy) in zip(c_idx, triples):
            checks = [("aa", this_root, c, ay), ("bb", this_root, c + half, by)]
            if next_root is not None:
                checks.append(("cc", next_root, c, cy))
            for label, tree_root, index, leaf in checks:
                if not Merkle.verify(tree_root, index, proof_stream.pull(), leaf):
                    print("merkle authentication path verification fails for " + label)
                    return False
        if next_root is None and any(cy != last_codeword[c] for c, (_, _, cy) in zip(c_idx, triples)):
            print("leafs in last round do not correspond to last codeword")
            return False
        return True

    def verify(self, proof_stream, root):
        """Returns False (after printing why) on rejection."""
        rounds, n = self.num_rounds(), self.domain.length
        omega, offset = self.field.lift(self.domain.omega), self.field.lift(self.domain.offset)
        roots, alphas = self._read_commitments(proof_stream, root)
        last_codeword = proof_stream.pull()
        squarings = 1 << (rounds - 1)
        if not self._last_codeword_ok(last_codeword, roots[-1], omega ^ squarings, offset ^ squarings):
            return False
        top = self.sample_indices(proof_stream.verifier_fiat_shamir(), n >> 1, n >> (rounds - 1),
                                  self.num_colinearity_tests)
        for r in range(rounds - 1):
            half = n >> (r + 1)
            final = r + 2 == rounds
            if not self._layer_ok(proof_stream, roots[r], None if final else roots[r + 1], alphas[r], omega, offset,
                                  [i % half for i in top], half, last_codeword):
                return False
            omega, offset = omega ^ 2, offset ^ 2
        return True
