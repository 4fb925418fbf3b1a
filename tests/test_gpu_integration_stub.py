"""The ctypes stub that INTEGRATION.md shows a kereHMM maintainer is executed as written (only the library path is
pointed at the in-tree build) and checked against the reference's golden outputs -- the document cannot drift from
the ABI."""
import os
import re
import types

import numpy as np
import pytest

from conftest import gcase

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_module():
    with open(os.path.join(ROOT, "INTEGRATION.md")) as f:
        text = f.read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    code = next(b for b in blocks if "kerehmm/_khmm.py" in b)
    lib = os.path.join(ROOT, "kerehmm_b200", "libkhmm.so")
    assert 'ctypes.CDLL("libkhmm.so")' in code
    code = code.replace('ctypes.CDLL("libkhmm.so")', "ctypes.CDLL(%r)" % lib)
    mod = types.ModuleType("_khmm_stub")
    exec(compile(code, "INTEGRATION.md:_khmm.py", "exec"), mod.__dict__)
    return mod


class _RefLikeModel:
    """Just the attributes the stub reads from a reference DiscreteHMM."""

    def __init__(self, pi, A, Bm):
        self.nStates, self.alphabetSize = A.shape[0], Bm.shape[1]
        self.initialProbabilities, self.transitionMatrix = pi, A
        self.emissionDistributions = [types.SimpleNamespace(probabilities=Bm[i]) for i in range(A.shape[0])]


@pytest.mark.parametrize("name", ["D1", "D3"])
def test_integration_stub_matches_reference_golden(golden, name):
    g = gcase(golden, name)
    stub = _stub_module()
    hmm = _RefLikeModel(g["pi"], g["A"], g["B"])
    alpha, scale = stub.forward(hmm, g["obs"])
    np.testing.assert_allclose(alpha, g["alpha"], rtol=1e-10, atol=1e-300)
    np.testing.assert_allclose(scale, g["scale"], rtol=1e-10)
    path, logp = stub.viterbi_path(hmm, g["obs"])
    assert np.array_equal(path, g["path"]) and logp == float(g["logp"])
    assert path.dtype == np.asarray(g["obs"]).dtype                      # abstract_hmm.py:234 dtype rule
