"""Acceptance of GPU-produced proofs by the UNMODIFIED reference verifier.

`tests/test_gpu_frontend.py::test_fri_big_transcripts` (run on the B200) writes the FRI transcripts
it produced to gpurun_out/; the copies committed under profiles/artifacts/ are fed here to the real
reference's Fri.verify (authoring container only -- the reference never travels to the GPU box)."""
import hashlib
import os
import sys

import pytest

from util import golden, have_golden, rand_xfe, root_of_unity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("B2S_REFERENCE_DIR", "/root/reference/code")


@pytest.mark.parametrize("logn", (16, 18, 20))
def test_reference_verifier_accepts_gpu_transcript(logn):
    path = os.path.join(ROOT, "profiles", "artifacts", "fri_%d_transcript.bin" % logn)
    if not os.path.isdir(REF) or not os.path.exists(path) or not have_golden("fri_%d.json" % logn):
        pytest.skip("needs the reference checkout, the GPU artifact and the golden digest")
    from conftest import _purge
    saved = _purge()
    sys.path.insert(0, REF)
    old = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        from algebra import BaseField
        from extension_field import ExtensionField
        from fri import Fri
        from ip import ProofStream
        from oracle import oracle as orc
        ser = open(path, "rb").read()
        e = golden("fri_%d.json" % logn)
        assert hashlib.sha256(ser).hexdigest() == e["transcript_sha256"]  # byte-identical to the reference's own proof
        field, xfield = BaseField.main(), ExtensionField.main()
        n = 1 << logn
        fri = Fri(field.generator(), field.primitive_nth_root(n), n, 4, 8, xfield)
        # root of the committed codeword, recomputed independently (oracle) from the seeded polynomial
        import pickle
        from extension_field import ExtensionFieldElement
        from univariate import Polynomial
        from algebra import BaseFieldElement
        bf = xfield.modulus.coefficients[0].field
        mk = lambda c: ExtensionFieldElement(Polynomial([BaseFieldElement(v, bf) for v in c]), xfield)  # noqa: E731
        tpl = orc.templates_from_marker_pickles([pickle.dumps(mk([0xA1, 0xA2, 0xA3][:k])) for k in range(4)], 3, True)
        cw = orc.coset_evaluate(7, root_of_unity(logn), rand_xfe(200 + logn, n // 4), n)
        root0 = bytes(orc.merkle_field(tpl, cw)[1])
        ps = ProofStream().deserialize(ser)
        assert fri.verify(ps, root0) is True
    finally:
        sys.dont_write_bytecode = old
        sys.path.remove(REF)
        _purge()
        sys.modules.update(saved)
