"""Cross-check the oracle's DSP specification (parity UNPINNED against the
reference: these Processors are build-defined) against
independent implementations (scipy.signal) and closed forms, in float64."""
import numpy as np
import pytest
from scipy import signal

import _oracle as orc
from pipe_b200 import design

RNG = np.random.default_rng(7)


def _chunks(x, sizes):
    i = 0
    for s in sizes:
        yield x[i:i + s]
        i += s


def _run_chunked(chain, x, sizes, threads=1):
    outs, counts = [], []
    for blk in _chunks(x, sizes):
        y = chain.process(blk, threads=threads)
        outs.append(y)
        counts.append(len(y))
    return np.concatenate(outs, axis=0), counts


def test_gain_exact():
    x = RNG.standard_normal((100, 3))
    y = orc.Chain(3, [{"kind": "gain", "gain": 0.37}]).process(x)
    assert np.array_equal(y, 0.37 * x)


def test_fir_impulse_returns_taps_exactly():
    taps = design.lowpass_fir(257, 20000 / 48000)
    x = np.zeros((600, 2))
    x[0, 0] = 1.0
    x[5, 1] = 1.0
    y = orc.Chain(2, [{"kind": "fir", "taps": taps}]).process(x)
    assert np.array_equal(y[:257, 0], taps)
    assert np.array_equal(y[5:262, 1], taps)
    assert not y[257:, 0].any()


def test_fir_matches_scipy_and_carries_history():
    taps = design.lowpass_fir(257, 0.2)
    x = RNG.uniform(-1, 1, (5000, 4))
    ref = signal.lfilter(taps, [1.0], x, axis=0)
    y, _ = _run_chunked(orc.Chain(4, [{"kind": "fir", "taps": taps}]), x, [1, 255, 256, 257, 1000, 3231])
    np.testing.assert_allclose(y, ref, rtol=0, atol=1e-13)


def test_fir_dc_gain_unity():
    taps = design.lowpass_fir(257, 0.2)
    y = orc.Chain(1, [{"kind": "fir", "taps": taps}]).process(np.ones((1000, 1)))
    np.testing.assert_allclose(y[300:], 1.0, atol=1e-13)


@pytest.mark.parametrize("kind,f0,q", [("lowpass", 8000.0, 0.9), ("highpass", 100.0, 0.707), ("peaking", 1000.0, 1.2)])
def test_biquad_matches_scipy_and_carries_state(kind, f0, q):
    b, a = design.biquad(kind, f0, 48000.0, q=q, gain_db=-3.0)
    x = RNG.uniform(-1, 1, (4096, 3))
    ref = signal.lfilter(b, np.concatenate([[1.0], a]), x, axis=0)
    y, _ = _run_chunked(orc.Chain(3, [{"kind": "biquad", "b": b, "a": a}]), x, [7, 512, 1, 3576])
    np.testing.assert_allclose(y, ref, rtol=0, atol=1e-11)


def test_biquad_impulse_against_analytic_recursion():
    b, a = design.biquad("lowpass", 8000.0, 48000.0, q=0.9)
    n = 64
    h = np.zeros(n)
    for i in range(n):  # direct form I difference equation
        acc = b[i] if i < 3 else 0.0
        if i >= 1:
            acc -= a[0] * h[i - 1]
        if i >= 2:
            acc -= a[1] * h[i - 2]
        h[i] = acc
    x = np.zeros((n, 1))
    x[0] = 1.0
    y = orc.Chain(1, [{"kind": "biquad", "b": b, "a": a}]).process(x)
    np.testing.assert_allclose(y[:, 0], h, atol=1e-14)


def _upfirdn_ref(proto, x, up, down):
    # ours[m] == upfirdn([0]+h, x, up, down)[m+1]
    full = signal.upfirdn(np.concatenate([[0.0], proto]), x, up, down, axis=0)
    return full[1:]


@pytest.mark.parametrize("up,down,tpp", [(147, 160, 16), (1, 2, 8), (2, 3, 12), (1, 1, 4)])
def test_resampler_matches_upfirdn(up, down, tpp):
    proto = design.resampler_prototype(up, down, tpp)
    x = RNG.uniform(-1, 1, (3000, 2))
    chain = orc.Chain(2, [{"kind": "resample", "up": up, "down": down, "taps": proto}])
    y, counts = _run_chunked(chain, x, [1, 159, 160, 1000, 1680])
    ref = _upfirdn_ref(proto, x, up, down)
    assert len(y) == (3000 * up) // down
    np.testing.assert_allclose(y, ref[:len(y)], rtol=0, atol=1e-13)
    # frame counts follow the integer phase accumulator exactly
    acc, want = 0, []
    for n in [1, 159, 160, 1000, 1680]:
        tot = acc + n * up
        want.append(tot // down)
        acc = tot % down
    assert counts == want


def test_resampler_frame_sequence_48k_to_44k1():
    # 4096-frame buffers -> 3763,3763,3763,3763,3764,...
    proto = design.resampler_prototype(147, 160, 16)
    chain = orc.Chain(1, [{"kind": "resample", "up": 147, "down": 160, "taps": proto}])
    counts = [chain.peek_out_frames(4096) or 0 for _ in range(1)]
    seq = []
    for _ in range(10):
        seq.append(len(chain.process(np.zeros((4096, 1)))))
    assert counts[0] == 3763
    assert seq == [3763, 3763, 3763, 3763, 3764] * 2
    assert sum(seq[:5]) == 5 * 4096 * 147 // 160


def test_resampler_dc_gain():
    proto = design.resampler_prototype(147, 160, 16)
    y = orc.Chain(1, [{"kind": "resample", "up": 147, "down": 160, "taps": proto}]).process(np.ones((2000, 1)))
    np.testing.assert_allclose(y[100:], 1.0, atol=2e-3)


def test_resample_up_gt_down_rejected():
    with pytest.raises(ValueError):
        orc.Chain(1, [{"kind": "resample", "up": 3, "down": 2, "taps": np.ones(6)}])


def test_chain4_equals_composition_and_mt_is_identical():
    stages = design.config_stages("chain4")
    x = RNG.uniform(-1, 1, (4096 * 3, 8))
    y, counts = _run_chunked(orc.Chain(8, stages), x, [4096] * 3)
    # independent composition with scipy
    g = stages[0]["gain"]
    s1 = signal.lfilter(stages[1]["taps"], [1.0], g * x, axis=0)
    s2 = signal.lfilter(stages[2]["b"], np.concatenate([[1.0], stages[2]["a"]]), s1, axis=0)
    ref = _upfirdn_ref(stages[3]["taps"], s2, 147, 160)
    np.testing.assert_allclose(y, ref[:len(y)], rtol=0, atol=1e-11)
    assert counts == [3763, 3763, 3763]
    y_mt, _ = _run_chunked(orc.Chain(8, stages), x, [4096] * 3, threads=3)
    assert np.array_equal(y, y_mt)


def test_set_stage_between_buffers_keeps_state():
    b, a = design.biquad("lowpass", 8000.0, 48000.0, q=0.9)
    stages = [{"kind": "gain", "gain": 1.0}, {"kind": "biquad", "b": b, "a": a}]
    x = RNG.uniform(-1, 1, (200, 1))
    c = orc.Chain(1, stages)
    y0 = c.process(x[:100])
    c.set_stage(0, {"kind": "gain", "gain": 2.0})
    y1 = c.process(x[100:])
    zi = signal.lfiltic(b, np.concatenate([[1.0], a]), y0[::-1, 0], x[:100][::-1, 0])
    ref, _ = signal.lfilter(b, np.concatenate([[1.0], a]), 2.0 * x[100:, 0], zi=zi)
    np.testing.assert_allclose(y1[:, 0], ref, atol=1e-12)


def test_mix_sum_and_meter_and_source():
    a, b, c = (RNG.standard_normal((50, 4)) for _ in range(3))
    np.testing.assert_array_equal(orc.mix_sum([a, b, c]), (a + b) + c)
    peak, sumsq = orc.meter(a)
    np.testing.assert_array_equal(peak, np.abs(a).max(axis=0))
    np.testing.assert_allclose(sumsq, (a * a).sum(axis=0), rtol=1e-14)
    x = orc.source_fill(0, 4096, seed=1234, line=0)
    assert x.min() >= -1.0 and x.max() < 1.0
    assert np.array_equal(x.astype(np.float32).astype(np.float64), x)  # exact in float32
    assert np.array_equal(orc.source_fill(100, 10), x[100:110])
    assert abs(x.mean()) < 0.05 and not np.array_equal(x[:10], orc.source_fill(0, 10, line=1))


def test_empty_buffer():
    c = orc.Chain(2, design.config_stages("chain4"))
    assert c.process(np.zeros((0, 2))).shape == (0, 2)
