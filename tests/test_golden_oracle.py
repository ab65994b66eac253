"""Oracle vs the committed scipy-generated fixtures (tests/golden/make_golden.py)."""
import os

import numpy as np

import _oracle as orc
from pipe_b200 import design

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "dsp_golden.npz"))


def _chunked(chain, x, size):
    outs = [chain.process(x[i:i + size]) for i in range(0, len(x), size)]
    return np.concatenate(outs), [len(o) for o in outs]


def test_case_a_gain_biquad():
    y, _ = _chunked(orc.Chain(8, design.config_stages("gain_biquad")), G["a_x"].astype(np.float64), 256)
    np.testing.assert_allclose(y, G["a_y"], rtol=0, atol=1e-12)


def test_case_b_chain4():
    y, counts = _chunked(orc.Chain(4, design.config_stages("chain4")), G["b_x"].astype(np.float64), 640)
    assert counts == list(G["b_counts"])
    np.testing.assert_allclose(y, G["b_y"], rtol=0, atol=1e-12)


def test_case_c_fir_impulse_dc_sweep():
    st = design.config_stages("chain4")[1]
    y, _ = _chunked(orc.Chain(2, [st]), G["c_x"].astype(np.float64), 300)
    np.testing.assert_allclose(y, G["c_y"], rtol=0, atol=1e-13)
