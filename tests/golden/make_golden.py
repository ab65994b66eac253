"""Generate tests/golden/dsp_golden.npz.

The reference (Go) cannot be built or imported in this image and ships none of
the DSP Processors, so these vectors do NOT come from the reference: they are
produced by scipy.signal (lfilter / upfirdn), an implementation independent of
both oracle/pipe_oracle.c and the CUDA kernels, from seeded inputs.  The
reference-derived integer goldens live in tests/test_oracle_plumbing.py with
their pipe_test.go / mock_test.go line numbers.

Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
from scipy import signal

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from pipe_b200 import design  # noqa: E402


def upfirdn_ours(proto, x, up, down, n_out):
    return signal.upfirdn(np.concatenate([[0.0], proto]), x, up, down, axis=0)[1:1 + n_out]


def main():
    rng = np.random.default_rng(20261017)
    out = {}
    # case A: configs[1] shape in miniature -- gain + biquad, 8 ch, 3 buffers of 256
    st = design.config_stages("gain_biquad")
    x = rng.uniform(-1, 1, (768, 8)).astype(np.float32).astype(np.float64)
    y = signal.lfilter(st[1]["b"], np.concatenate([[1.0], st[1]["a"]]), st[0]["gain"] * x, axis=0)
    out["a_x"], out["a_y"] = x.astype(np.float32), y
    # case B: configs[2] in miniature -- the 4-stage chain, 4 ch, 3 buffers of 640
    st = design.config_stages("chain4")
    x = rng.uniform(-1, 1, (1920, 4)).astype(np.float32).astype(np.float64)
    s1 = signal.lfilter(st[1]["taps"], [1.0], st[0]["gain"] * x, axis=0)
    s2 = signal.lfilter(st[2]["b"], np.concatenate([[1.0], st[2]["a"]]), s1, axis=0)
    n_out = 1920 * 147 // 160
    out["b_x"], out["b_y"] = x.astype(np.float32), upfirdn_ours(st[3]["taps"], s2, 147, 160, n_out)
    out["b_counts"] = np.array([588, 588, 588], dtype=np.int64)  # (640*147)//160 per buffer, acc returns to 0
    # case C: impulse + DC + sine sweep through the 257-tap FIR only, 2 ch
    n = 1024
    t = np.arange(n) / 48000.0
    sweep = np.sin(2 * np.pi * (20.0 * t + (20000.0 - 20.0) / (2 * t[-1]) * t * t))
    x = np.stack([np.r_[1.0, np.zeros(n - 1)], 0.5 * sweep + 0.25], axis=1)
    x = x.astype(np.float32).astype(np.float64)
    out["c_x"], out["c_y"] = x.astype(np.float32), signal.lfilter(st[1]["taps"], [1.0], x, axis=0)
    np.savez_compressed(os.path.join(HERE, "dsp_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
