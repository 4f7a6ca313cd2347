"""Host-side mirror of the reference's `merkle` module (code/merkle.py): same constructor,
root/open/verify and public attributes (leafs, nodes, depth, num_leafs); the tree is built
on the device."""
import pickle
from hashlib import blake2b


class Merkle:
    __module__ = "merkle"

    def __init__(self, data_array):
        from . import glue
        glue().merkle_build(self, data_array)

    def root(self):
        return self.nodes[1]

    def open(self, index):
        from . import glue
        return glue().merkle_open(self, index)

    @staticmethod
    def verify(root, index, path, element):
        """recompute the root from a leaf and its siblings (code/merkle.py:54-63); host-side,
        it is the verifier's check"""
        h = blake2b(pickle.dumps(element)).digest()
        for sibling in path:
            h = blake2b(h + sibling).digest() if index % 2 == 0 else blake2b(sibling + h).digest()
            index >>= 1
        return h == root
