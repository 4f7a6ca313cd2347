"""Host-side mirror of the reference's `ip` module (code/ip.py): the proof stream IS the
transcript format (pickle of the pushed objects) and the Fiat-Shamir source."""
import pickle
from hashlib import shake_256


class ProofStream:
    __module__ = "ip"

    def __init__(self):
        self.objects = []
        self.read_index = 0

    def push(self, obj):
        self.objects += [obj]

    def pull(self):
        assert self.read_index < len(self.objects), "ProofStream: cannot pull object; queue empty."
        self.read_index += 1
        return self.objects[self.read_index - 1]

    def serialize(self):
        return pickle.dumps(self.objects)

    def prover_fiat_shamir(self, num_bytes=32):
        return shake_256(self.serialize()).digest(num_bytes)

    def verifier_fiat_shamir(self, num_bytes=32):
        return shake_256(pickle.dumps(self.objects[:self.read_index])).digest(num_bytes)

    def deserialize(self, bb):
        ps = ProofStream()
        ps.objects = pickle.loads(bb)
        return ps
