"""Host-side mirror of the reference's `salted_merkle` module (code/salted_merkle.py): same constructor,
root/open/verify and public attributes; leaves are (element, salt) pairs, hashed as
blake2b(pickle(element) | pickle(salt)).  The tree is built on the device: rows of field elements are
pickled there from a per-tree byte template (glue.Glue._row_tree), anything else is pickled here and
hashed there."""
import os
import pickle
from hashlib import blake2b

urandom = os.urandom  # looked up at call time, like the reference's `from os import urandom`


class SaltedMerkle:
    __module__ = "salted_merkle"

    def __init__(self, data_array):
        from . import glue
        import sys
        me = sys.modules[__name__]
        glue().salted_merkle_build(self, data_array, lambda k: me.urandom(k))

    def root(self):
        return self.nodes[1]

    def open(self, index):
        """(salt, authentication path), code/salted_merkle.py:49-56"""
        from . import glue
        return self.leafs[index][1], glue().merkle_open(self, index)

    @staticmethod
    def verify(root, index, salt, path, element):
        """host-side check of one opening (code/salted_merkle.py:58-68)"""
        h = blake2b(pickle.dumps(element) + pickle.dumps(salt)).digest()
        for sibling in path:
            h = blake2b(sibling + h).digest() if index & 1 else blake2b(h + sibling).digest()
            index >>= 1
        return h == root
