"""Coefficient design helpers for the build-defined Processors (host side, numpy only).

The reference ships no DSP Processors; a component
library would bring its own coefficients.  These helpers produce the float64
coefficient sets used by the configs in BASELINE.json so that the oracle, the
tests and bench.py all feed the GPU chain and the CPU oracle identical numbers.
Nothing here runs per buffer.
"""
from __future__ import annotations

import numpy as np


def lowpass_fir(n_taps: int, cutoff: float, beta: float = 8.6) -> np.ndarray:
    """Kaiser-windowed sinc low-pass, unity DC gain.  cutoff in cycles/sample (0..0.5)."""
    n = np.arange(n_taps, dtype=np.float64) - (n_taps - 1) / 2.0
    h = 2.0 * cutoff * np.sinc(2.0 * cutoff * n) * np.kaiser(n_taps, beta)
    return h / h.sum()


def resampler_prototype(up: int, down: int, taps_per_phase: int = 16, beta: float = 8.6,
                        rolloff: float = 0.9) -> np.ndarray:
    """Prototype low-pass for the rational up/down polyphase resampler.

    Length up*taps_per_phase, designed at the up-sampled rate with cutoff
    rolloff*0.5/max(up, down) and pass-band gain `up` (so that each polyphase
    branch has roughly unity DC gain).
    """
    n_taps = up * taps_per_phase
    cutoff = rolloff * 0.5 / max(up, down)
    return lowpass_fir(n_taps, cutoff, beta) * up


def biquad(kind: str, f0: float, fs: float, q: float = 0.7071067811865476, gain_db: float = 0.0):
    """RBJ audio-EQ-cookbook biquad.  Returns (b[3], a[2]) normalised to a0 == 1."""
    w0 = 2.0 * np.pi * f0 / fs
    cw, sw = np.cos(w0), np.sin(w0)
    alpha = sw / (2.0 * q)
    if kind == "lowpass":
        b = np.array([(1 - cw) / 2, 1 - cw, (1 - cw) / 2])
        a = np.array([1 + alpha, -2 * cw, 1 - alpha])
    elif kind == "highpass":
        b = np.array([(1 + cw) / 2, -(1 + cw), (1 + cw) / 2])
        a = np.array([1 + alpha, -2 * cw, 1 - alpha])
    elif kind == "peaking":
        A = 10.0 ** (gain_db / 40.0)
        b = np.array([1 + alpha * A, -2 * cw, 1 - alpha * A])
        a = np.array([1 + alpha / A, -2 * cw, 1 - alpha / A])
    else:
        raise ValueError(f"unknown biquad kind {kind!r}")
    return (b / a[0]).astype(np.float64), (a[1:] / a[0]).astype(np.float64)


def config_stages(name: str, fs: float = 48000.0) -> list[dict]:
    """Stage lists for the BASELINE.json configs (coefficients are build-defined)."""
    if name == "passthrough":          # configs[0]: mock.Processor
        return [{"kind": "copy"}]
    if name == "gain_biquad":          # configs[1]
        b, a = biquad("lowpass", 8000.0, fs, q=0.9)
        return [{"kind": "gain", "gain": 0.5}, {"kind": "biquad", "b": b, "a": a}]
    if name == "chain4":               # configs[2..3]: gain, 257-tap FIR, biquad, 48k->44.1k
        b, a = biquad("peaking", 1000.0, fs, q=1.2, gain_db=-3.0)
        return [
            {"kind": "gain", "gain": 0.8},
            {"kind": "fir", "taps": lowpass_fir(257, 20000.0 / fs)},
            {"kind": "biquad", "b": b, "a": a},
            {"kind": "resample", "up": 147, "down": 160, "taps": resampler_prototype(147, 160, 16)},
        ]
    raise ValueError(name)
