"""Freeze the reference's OWN encoder file as a test fixture (run in the build container, where /root/reference exists).

    python tests/golden/make_reference_fixture.py

Writes tests/golden/reference_models_SparseConvNet.py.fixture = /root/reference/models/SparseConvNet.py byte for byte
behind a three-line provenance header.  It is TEST DATA, not product code: `tests/test_reference_models.py` executes it
against this repository's `sparseconvnet` package (and against the CPU oracle) with the stubs SURVEY.md Appendix A lists
(`easydict`, `utils.registry`), which is how "existing configs run as a drop-in" (BASELINE.json north_star) is checked
on a GPU box that has no /root/reference.  Nothing under 3d-weakly-supervised-semantic-segmentation_b200/ reads it.
"""
import hashlib
import os

SRC = "/root/reference/models/SparseConvNet.py"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_models_SparseConvNet.py.fixture")

if __name__ == "__main__":
    body = open(SRC, "rb").read()
    head = ("# TEST FIXTURE -- unmodified copy of the reference file models/SparseConvNet.py (timsu1104/3D-Weakly-Supervised-"
            "Semantic-Segmentation),\n# made by tests/golden/make_reference_fixture.py; sha256 of the original: %s\n"
            "# Not product code: executed only by tests/test_reference_models.py to prove the drop-in claim.\n"
            % hashlib.sha256(body).hexdigest()).encode()
    open(DST, "wb").write(head + body)
    print(DST, len(body))
