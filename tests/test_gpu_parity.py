"""GPU parity tests: the CUDA path, called through the C-ABI (pipe_b200.abi ->
libpipe_b200.so), against the CPU oracle on identical seeded inputs.

Bar (BASELINE.json north_star / SURVEY.md H2):
  * integer bookkeeping (frames per buffer, message counts): bit-exact;
  * float32 path vs the float64 oracle: max|y - ref| <= 1e-6 * max|ref| per
    channel per buffer (REL_F32 below);
  * float64 path vs the float64 oracle: 1e-12 on the same measure.
  * buffers shorter than SHORT_BUFFER frames are a degenerate case of "per
    buffer" (a 1-frame buffer's peak is one sample, possibly at a zero
    crossing): for those the per-channel peak is taken over the stream so far.
"""
import os

import numpy as np
import pytest

import _oracle as orc
from pipe_b200 import abi, design

pytestmark = pytest.mark.gpu

REL_F32 = 1e-6
REL_F64 = 1e-12
SHORT_BUFFER = 256


def assert_parity(y, ref, rel, what="", floor=None):
    assert y.shape == ref.shape, f"{what}: shape {y.shape} vs {ref.shape}"
    if ref.size == 0:
        return
    ref = np.asarray(ref, dtype=np.float64)
    scale = np.abs(ref).max(axis=0)
    if floor is not None and len(ref) < SHORT_BUFFER:
        scale = np.maximum(scale, floor)
    err = np.abs(np.asarray(y, dtype=np.float64) - ref).max(axis=0)
    bad = err > rel * scale + 1e-300
    assert not bad.any(), (f"{what}: worst channel {int(np.argmax(err / (scale + 1e-300)))} "
                           f"err {err.max():.3e} vs scale {scale[np.argmax(err)]:.3e} "
                           f"(ratio {np.max(err / (scale + 1e-300)):.3e}, allowed {rel:.1e})")


def signal_input(frames, channels, seed=0, line=0):
    x = orc.source_fill(seed * 7919, frames * channels, seed=1234, line=line)
    return x.reshape(frames, channels)


def run_both(stages, channels, sizes, dtype=np.float32, buffer_frames=None, seed=0, rel=None, max_batch=1):
    rel = rel if rel is not None else (REL_F32 if dtype == np.float32 else REL_F64)
    bf = buffer_frames or max(max(sizes), 1)
    gpu = abi.Chain(channels, stages, buffer_frames=bf, dtype=dtype, max_batch=max_batch)
    cpu = orc.Chain(channels, stages)
    x = signal_input(sum(sizes), channels, seed)
    pos = 0
    run_peak = np.zeros(channels)
    for i, n in enumerate(sizes):
        blk = x[pos:pos + n]
        pos += n
        ref = cpu.process(blk)
        if len(ref):
            run_peak = np.maximum(run_peak, np.abs(ref).max(axis=0))
        assert gpu.peek_out_frames(n) == len(ref)
        y = gpu.process(blk.astype(dtype))
        assert len(y) == len(ref), f"buffer {i}: frames {len(y)} vs {len(ref)}"  # bit-exact bookkeeping
        assert_parity(y, ref, rel, f"buffer {i} ({n} frames)", floor=run_peak)
    gpu.close()


# ------------------------------------------------------------------ configs --

def test_config1_passthrough_f64_is_bit_exact():
    # configs[0]: mock.Processor pass-through, 2 ch float64, 512-frame buffers
    gpu = abi.Chain(2, [{"kind": "copy"}], buffer_frames=512, dtype=np.float64)
    x = signal_input(512 * 3 + 17, 2)
    for i in range(0, len(x), 512):
        blk = x[i:i + 512]
        assert np.array_equal(gpu.process(blk), blk)
    # the reference's own value goldens (mock_test.go:133-146)
    for vals in ([1, 1, 1, 1], [1, 1, 1, 1, 2, 2, 2, 2]):
        g1 = abi.Chain(1, [{"kind": "copy"}], buffer_frames=8, dtype=np.float64)
        v = np.asarray(vals, dtype=np.float64).reshape(-1, 1)
        assert np.array_equal(g1.process(v), v)


def test_config2_gain_biquad_64ch_4096():
    run_both(design.config_stages("gain_biquad"), 64, [4096] * 4)


def test_config3_chain4_small_channels():
    run_both(design.config_stages("chain4"), 8, [1024] * 6)


def test_config3_chain4_1024ch_4096_counts_and_values():
    stages = design.config_stages("chain4")
    gpu = abi.Chain(1024, stages, buffer_frames=4096)
    cpu = orc.Chain(1024, stages)
    counts = []
    for b in range(2):
        x = signal_input(4096, 1024, seed=b + 1)
        ref = cpu.process(x, threads=os.cpu_count() or 1)
        y = gpu.process(x.astype(np.float32))
        counts.append(len(y))
        assert_parity(y, ref, REL_F32, f"buffer {b}")
    assert counts == [3763, 3763]


def test_resampler_frame_sequence_bit_exact():
    # 4096-frame buffers -> 3763,3763,3763,3763,3764 repeating
    stages = [design.config_stages("chain4")[3]]
    gpu = abi.Chain(4, stages, buffer_frames=4096)
    seq = [len(gpu.process(np.zeros((4096, 4), np.float32))) for _ in range(10)]
    assert seq == [3763, 3763, 3763, 3763, 3764] * 2


# --------------------------------------------------------------- edge cases --

@pytest.mark.parametrize("channels", [1, 3, 33, 100])
def test_odd_channel_counts(channels):
    run_both(design.config_stages("chain4"), channels, [700, 300])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_ragged_buffer_sizes_carry_state(dtype):
    sizes = [1, 7, 255, 256, 257, 1000, 0, 3, 15, 16, 17, 2048]
    run_both(design.config_stages("chain4"), 5, sizes, dtype=dtype, buffer_frames=2048)


def test_short_final_buffer_like_pipe_test():
    # pipe_test.go:337: Limit 1040 with bufferSize 512 -> 512, 512, 16
    run_both(design.config_stages("gain_biquad"), 1, [512, 512, 16], buffer_frames=512)


def test_empty_buffer_is_a_no_op():
    gpu = abi.Chain(2, design.config_stages("chain4"), buffer_frames=64)
    assert gpu.process(np.zeros((0, 2), np.float32)).shape == (0, 2)


@pytest.mark.parametrize("n_taps", [1, 2, 16, 33, 257, 600])
def test_fir_lengths(n_taps):
    taps = design.lowpass_fir(n_taps, 0.2) if n_taps > 2 else np.array([0.75, -0.25][:n_taps])
    run_both([{"kind": "fir", "taps": taps}], 6, [500, 500, 41])


def test_fir_impulse_returns_taps():
    taps = design.lowpass_fir(257, 20000 / 48000)
    gpu = abi.Chain(2, [{"kind": "fir", "taps": taps}], buffer_frames=600, dtype=np.float64)
    x = np.zeros((600, 2))
    x[0, 0] = 1.0
    y = gpu.process(x)
    assert np.array_equal(y[:257, 0], taps)
    assert not y[:, 1].any()


@pytest.mark.parametrize("kind,f0,q", [("lowpass", 8000.0, 0.9), ("highpass", 200.0, 0.707), ("peaking", 1000.0, 4.0)])
def test_biquad_kinds(kind, f0, q):
    b, a = design.biquad(kind, f0, 48000.0, q=q, gain_db=6.0)
    run_both([{"kind": "biquad", "b": b, "a": a}], 40, [4096, 4096, 100])


@pytest.mark.parametrize("up,down,tpp", [(147, 160, 16), (1, 2, 8), (2, 3, 12), (1, 1, 4), (147, 160, 32)])
def test_resampler_ratios(up, down, tpp):
    proto = design.resampler_prototype(up, down, tpp)
    run_both([{"kind": "resample", "up": up, "down": down, "taps": proto}], 7, [1000, 1, 159, 160, 1680])


def test_multi_segment_chains():
    b1, a1 = design.biquad("lowpass", 6000.0, 48000.0, q=0.8)
    b2, a2 = design.biquad("highpass", 300.0, 48000.0, q=0.7)
    proto = design.resampler_prototype(1, 2, 8)
    stages = [
        {"kind": "biquad", "b": b1, "a": a1}, {"kind": "gain", "gain": 1.5},
        {"kind": "biquad", "b": b2, "a": a2},                       # second biquad -> second segment
        {"kind": "resample", "up": 1, "down": 2, "taps": proto},
        {"kind": "fir", "taps": design.lowpass_fir(31, 0.2)},       # FIR after resample -> third segment
        {"kind": "copy"}, {"kind": "gain", "gain": -0.5},
    ]
    run_both(stages, 9, [999, 1000, 1, 500], buffer_frames=1000)


def test_batched_launch_equals_per_buffer_launches():
    stages = design.config_stages("chain4")
    ch, bf, nb = 64, 1024, 5
    x = signal_input(bf * nb - 100, ch).astype(np.float32)   # last buffer short
    sizes = [bf] * (nb - 1) + [bf - 100]
    cpu = orc.Chain(ch, stages)
    refs = [cpu.process(x[i * bf:i * bf + n]) for i, n in enumerate(sizes)]
    gpu = abi.Chain(ch, stages, buffer_frames=bf, max_batch=nb)
    d_in, d_out = abi.DeviceBuffer(x.nbytes), abi.DeviceBuffer(x.nbytes)
    d_in.upload(x)
    counts = gpu.process_batch_device(d_in.ptr, sizes, d_out.ptr, len(x))
    gpu.sync()
    assert counts == [len(r) for r in refs]
    y = d_out.download((sum(counts), ch), np.float32)
    pos = 0
    for i, r in enumerate(refs):
        assert_parity(y[pos:pos + len(r)], r, REL_F32, f"batched buffer {i}")
        pos += len(r)
    assert gpu.last_path()[0] in (1, 2, 3)


def test_reset_restarts_the_stream():
    stages = design.config_stages("chain4")
    gpu = abi.Chain(4, stages, buffer_frames=512)
    x = signal_input(512, 4).astype(np.float32)
    y0 = gpu.process(x)
    gpu.process(x)
    gpu.reset()
    assert np.array_equal(gpu.process(x), y0)


def test_mutations_between_buffers():
    # pipe.go:433: parameter changes land between buffers, state is kept
    b, a = design.biquad("lowpass", 8000.0, 48000.0, q=0.9)
    b2, a2 = design.biquad("lowpass", 2000.0, 48000.0, q=0.7)
    proto = design.resampler_prototype(2, 3, 12)
    stages = [{"kind": "gain", "gain": 1.0}, {"kind": "fir", "taps": design.lowpass_fir(65, 0.3)},
              {"kind": "biquad", "b": b, "a": a}, {"kind": "gain", "gain": 1.0},
              {"kind": "resample", "up": 2, "down": 3, "taps": proto}]
    gpu, cpu = abi.Chain(3, stages, buffer_frames=300), orc.Chain(3, stages)
    x = signal_input(1200, 3)
    edits = {1: (0, {"kind": "gain", "gain": 2.0}), 2: (2, {"kind": "biquad", "b": b2, "a": a2}),
             3: (3, {"kind": "gain", "gain": 0.25})}
    for i in range(4):
        if i in edits:
            gpu.set_stage(*edits[i])
            cpu.set_stage(*edits[i])
        blk = x[i * 300:(i + 1) * 300]
        assert_parity(gpu.process(blk.astype(np.float32)), cpu.process(blk), REL_F32, f"buffer {i}")


def test_fused_meter_sink_and_standalone_meter():
    stages = design.config_stages("gain_biquad")
    gpu = abi.Chain(40, stages, buffer_frames=1000, flags=abi.CHAIN_METER)
    cpu = orc.Chain(40, stages)
    x = signal_input(3000, 40)
    refs = np.concatenate([cpu.process(x[i:i + 1000]) for i in range(0, 3000, 1000)])
    ys = np.concatenate([gpu.process(x[i:i + 1000].astype(np.float32)) for i in range(0, 3000, 1000)])
    peak, sumsq, frames = gpu.meter_read()
    rp, rs = orc.meter(refs)
    assert frames == 3000
    np.testing.assert_allclose(peak, rp, rtol=2e-6)
    np.testing.assert_allclose(sumsq, rs, rtol=2e-6)
    # stand-alone meter kernel on the GPU output: exact peak, double-accumulated sum of squares
    d = abi.DeviceBuffer(ys.nbytes)
    d.upload(ys)
    acc = abi.DeviceBuffer(2 * 40 * 8)
    acc.upload(np.zeros(80))
    abi.meter_device(d.ptr, abi.PB_F32, 3000, 40, acc.ptr, acc.ptr + 40 * 8)
    got = acc.download((2, 40), np.float64)
    p2, s2 = orc.meter(ys.astype(np.float64))
    assert np.array_equal(got[0], p2)
    np.testing.assert_allclose(got[1], s2, rtol=1e-12)


def test_source_fill_matches_oracle_bit_exact():
    for dtype, pb in ((np.float32, abi.PB_F32), (np.float64, abi.PB_F64)):
        n = 100003
        d = abi.DeviceBuffer(n * np.dtype(dtype).itemsize)
        abi.source_fill(d.ptr, pb, 12345, n, seed=1234, line=3)
        got = d.download((n,), dtype)
        assert np.array_equal(got.astype(np.float64), orc.source_fill(12345, n, seed=1234, line=3))


def test_mix_sum_fan_in_bit_exact():
    # configs[4] in miniature on one device: 4 Lines x 256 ch, fan-in sum
    n = 256 * 1000
    ins = [signal_input(1000, 256, seed=i, line=i).astype(np.float32) for i in range(4)]
    bufs = []
    for a in ins:
        d = abi.DeviceBuffer(a.nbytes)
        d.upload(a)
        bufs.append(d)
    out = abi.DeviceBuffer(ins[0].nbytes)
    abi.mix_sum([b.ptr for b in bufs], abi.PB_F32, n, out.ptr)
    got = out.download((1000, 256), np.float32)
    ref = ((ins[0] + ins[1]) + ins[2]) + ins[3]          # float32, same order
    assert np.array_equal(got, ref)
    assert_parity(got, orc.mix_sum([a.astype(np.float64) for a in ins]), REL_F32)


def test_pipelined_submit_collect_with_pinned_buffers():
    stages = design.config_stages("chain4")
    ch, bf, nb, steps = 32, 512, 4, 5
    gpu = abi.Chain(ch, stages, buffer_frames=bf, max_batch=nb)
    cpu = orc.Chain(ch, stages)
    nbytes = bf * nb * ch * 4
    pin_in = [abi.PinnedBuffer(nbytes) for _ in range(2)]
    pin_out = [abi.PinnedBuffer(nbytes) for _ in range(2)]
    xs = [signal_input(bf * nb, ch, seed=s) for s in range(steps)]
    refs = [cpu.process(x) for x in xs]
    got = []
    for s in range(steps + 1):
        if s < steps:
            pin_in[s & 1].array((bf * nb, ch), np.float32)[:] = xs[s]
            gpu.submit(pin_in[s & 1].ptr, [bf] * nb, pin_out[s & 1].ptr, bf * nb)
        if s >= 1:
            counts = gpu.collect(nb)
            got.append(pin_out[(s - 1) & 1].array((bf * nb, ch), np.float32)[:sum(counts)].copy())
    for s in range(steps):
        assert_parity(got[s], refs[s], REL_F32, f"step {s}")


def test_golden_fixtures_from_scipy():
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "dsp_golden.npz"))
    for case, cfg, ch, bf in (("a", "gain_biquad", 8, 256), ("b", "chain4", 4, 640)):
        gpu = abi.Chain(ch, design.config_stages(cfg), buffer_frames=bf)
        x = G[f"{case}_x"]
        y = np.concatenate([gpu.process(x[i:i + bf]) for i in range(0, len(x), bf)])
        ref = G[f"{case}_y"]
        for i in range(0, len(ref), max(1, len(ref) // 3)):
            assert_parity(y[i:i + len(ref) // 3], ref[i:i + len(ref) // 3], REL_F32, f"golden {case}")
    gpu = abi.Chain(2, [design.config_stages("chain4")[1]], buffer_frames=300)
    y = np.concatenate([gpu.process(G["c_x"][i:i + 300]) for i in range(0, 1024, 300)])
    assert_parity(y, G["c_y"], REL_F32, "golden c")


# -------------------------------------------------------------------- errors --

def test_error_behaviour():
    with pytest.raises(abi.PipeB200Error) as e:
        abi.Chain(1, [{"kind": "resample", "up": 3, "down": 2, "taps": np.ones(6)}], buffer_frames=16)
    assert e.value.code == abi.PB_ERR_UNSUPPORTED
    with pytest.raises(abi.PipeB200Error) as e:
        abi.Chain(0, [{"kind": "copy"}], buffer_frames=16)
    assert e.value.code == abi.PB_ERR_INVALID
    gpu = abi.Chain(1, [{"kind": "copy"}], buffer_frames=16)
    with pytest.raises(abi.PipeB200Error) as e:
        gpu.process(np.zeros((17, 1), np.float32))          # more frames than bufferSize*max_batch
    assert e.value.code == abi.PB_ERR_INVALID
    with pytest.raises(abi.PipeB200Error) as e:
        gpu.collect(1)                                       # collect without submit
    assert e.value.code == abi.PB_ERR_STATE
    with pytest.raises(abi.PipeB200Error) as e:
        gpu.set_stage(0, {"kind": "gain", "gain": 2.0})      # kind may not change
    assert e.value.code == abi.PB_ERR_INVALID


# ------------------------------------------------------------ tcgen05 path (K2) --

def _k2_chain(channels, bf, nb, flags=0):
    return abi.Chain(channels, design.config_stages("chain4"), buffer_frames=bf, max_batch=nb, flags=flags)


@pytest.mark.parametrize("channels,bf,nb", [(128, 160, 1), (128, 1600, 1), (256, 4096, 5), (1024, 1600, 2)])
def test_k2_tensor_path_matches_oracle(channels, bf, nb):
    # calls aligned to 160-frame tiles take the tcgen05/TMA kernel (path 2); same 1e-6 bar as K1
    gpu, cpu = _k2_chain(channels, bf, nb), orc.Chain(channels, design.config_stages("chain4"))
    for step in range(3):
        x = signal_input(bf * nb, channels, seed=step)
        refs = [cpu.process(x[i * bf:(i + 1) * bf], threads=os.cpu_count() or 1) for i in range(nb)]
        y = gpu.process(x.astype(np.float32))
        assert gpu.last_path()[0] == 2
        assert len(y) == sum(len(r) for r in refs)
        pos = 0
        for i, r in enumerate(refs):
            assert_parity(y[pos:pos + len(r)], r, REL_F32, f"step {step} buffer {i}")
            pos += len(r)


def test_k2_and_k1_share_carried_state():
    # calls of at least one 160-frame tile run on K2 from their first frame at whatever resampler phase they start (last tile
    # partial), shorter calls on K1 alone; history, biquad state and resampler phase must carry across every hand-over
    ch = 128
    gpu, cpu = _k2_chain(ch, 1600, 1), orc.Chain(ch, design.config_stages("chain4"))
    paths = []
    sizes = [1600, 1600, 100, 60, 1600, 37, 1440 + 123, 1600]   # 160-aligned again after 100+60 and 37+1563
    x = signal_input(sum(sizes), ch, seed=5)
    pos = 0
    run_peak = np.zeros(ch)
    for n in sizes:
        blk = x[pos:pos + n]
        pos += n
        ref = cpu.process(blk)
        run_peak = np.maximum(run_peak, np.abs(ref).max(axis=0))
        y = gpu.process(blk.astype(np.float32))
        paths.append(gpu.last_path()[0])
        assert_parity(y, ref, REL_F32, f"{n} frames", floor=run_peak)
    assert paths == [2, 2, 1, 1, 2, 1, 2, 2]  # 1563 frames from position 4997: 123 on K1, then 1440 on K2


def test_k2_no_tensor_flag_and_meter():
    ch, bf = 128, 1600
    x = signal_input(bf, ch, seed=9)
    ref = orc.Chain(ch, design.config_stages("chain4")).process(x)
    g1 = _k2_chain(ch, bf, 1, flags=abi.CHAIN_NO_TENSOR)
    y1 = g1.process(x.astype(np.float32))
    assert g1.last_path()[0] == 1
    g2 = _k2_chain(ch, bf, 1, flags=abi.CHAIN_METER)
    y2 = g2.process(x.astype(np.float32))
    assert g2.last_path()[0] == 2
    assert_parity(y1, ref, REL_F32)
    assert_parity(y2, ref, REL_F32)
    peak, sumsq, frames = g2.meter_read()
    rp, rs = orc.meter(ref)
    assert frames == len(ref)
    np.testing.assert_allclose(peak, rp, rtol=2e-6)
    np.testing.assert_allclose(sumsq, rs, rtol=2e-6)


def test_k2_mutation_and_reset():
    ch, bf = 128, 1600
    st = design.config_stages("chain4")
    gpu, cpu = abi.Chain(ch, st, buffer_frames=bf), orc.Chain(ch, st)
    x = signal_input(3 * bf, ch, seed=11)
    b2, a2 = design.biquad("lowpass", 3000.0, 48000.0, q=0.8)
    for i in range(3):
        if i == 1:
            for c in (gpu, cpu):
                c.set_stage(0, {"kind": "gain", "gain": 0.3})
                c.set_stage(2, {"kind": "biquad", "b": b2, "a": a2})
        blk = x[i * bf:(i + 1) * bf]
        assert_parity(gpu.process(blk.astype(np.float32)), cpu.process(blk), REL_F32, f"buffer {i}")
        # the 3 kHz low-pass removes 8.6 dB of a broadband signal: the chain moves that run to the exact-order kernel
        # (K2's error is relative to the level inside the chain), carrying every piece of state across
        assert gpu.last_path()[0] == (2 if i == 0 else 1)
    gpu.reset()
    cpu.reset()
    assert_parity(gpu.process(x[:bf].astype(np.float32)), cpu.process(x[:bf]), REL_F32, "after reset")


@pytest.mark.parametrize("kind,f0,q,n_taps,path", [("highpass", 200.0, 0.707, 257, 2), ("lowpass", 8000.0, 0.9, 33, 2),
                                                   ("peaking", 1000.0, 4.0, 129, 2), ("lowpass", 120.0, 0.6, 257, 1),
                                                   # poles at |z| = 0.9963: a memory of thousands of samples stays on the kernel
                                                   # whose biquad state is double
                                                   ("highpass", 40.0, 0.707, 257, 1)])
def test_k2_other_filters(kind, f0, q, n_taps, path):
    # the biquad is folded per 16-row block into the resampler matrix and its state enters as a rank-2 correction:
    # poles close to z = 1 (slow state decay) and short FIRs must hold the same bar.  A biquad that removes most of a
    # broadband signal (the 120 Hz low-pass) is kept on the exact-order path K1, because K2's error is relative to the
    # level inside the chain.
    ch, bf = 128, 1600
    b, a = design.biquad(kind, f0, 48000.0, q=q, gain_db=6.0)
    st = [{"kind": "gain", "gain": 0.95}, {"kind": "fir", "taps": design.lowpass_fir(n_taps, 18000.0 / 48000.0)},
          {"kind": "gain", "gain": 1.0}, {"kind": "biquad", "b": b, "a": a}, {"kind": "gain", "gain": 0.8},
          {"kind": "resample", "up": 147, "down": 160, "taps": design.resampler_prototype(147, 160, 16)},
          {"kind": "gain", "gain": 1.25}]
    gpu, cpu = abi.Chain(ch, st, buffer_frames=bf), orc.Chain(ch, st)
    run_peak = np.zeros(ch)
    for step in range(4):
        x = signal_input(bf, ch, seed=20 + step)
        ref = cpu.process(x)
        run_peak = np.maximum(run_peak, np.abs(ref).max(axis=0))
        y = gpu.process(x.astype(np.float32))
        assert gpu.last_path()[0] == path
        assert_parity(y, ref, REL_F32, f"{kind} step {step}", floor=run_peak)


def test_k2_accuracy_follows_each_channels_own_level():
    # K2 splits every channel on its OWN integer grid (per-channel block exponent: chain_tc.cuh): channels at 0, -20, -40 and
    # -60 dBFS next to each other each meet 1e-6 of their own peak on the default path, on the first call (whose speculated
    # scale is wrong for the quiet channels: pass B redoes it) and on the following ones (no redo).
    ch, bf = 128, 1600
    amp = 10.0 ** (-np.array([0.0, 20.0, 40.0, 60.0])[np.arange(ch) % 4] / 20.0)
    gpu, cpu = _k2_chain(ch, bf, 1), orc.Chain(ch, design.config_stages("chain4"))
    for step in range(3):
        x = signal_input(bf, ch, seed=31 + step) * amp
        ref = cpu.process(x)
        y = gpu.process(x.astype(np.float32))
        assert gpu.last_path()[0] == 2
        assert_parity(y, ref, REL_F32, f"step {step}")   # per channel, against the channel's own peak


def test_k2_serves_input_beyond_full_scale():
    # floating-point pipelines legitimately exceed 1.0 (pipe.go:438-440: an error would end the run): |g x| up to 4
    ch, bf = 128, 1600
    gpu, cpu = _k2_chain(ch, bf, 1), orc.Chain(ch, design.config_stages("chain4"))
    for step in range(2):
        x = 5.0 * signal_input(bf, ch, seed=2 + step)     # gain 0.8: |g x| up to 4
        ref = cpu.process(x)
        y = gpu.process(x.astype(np.float32))
        assert gpu.last_path()[0] == 2
        assert_parity(y, ref, REL_F32, f"step {step}")


def test_k2_level_jumps_between_calls_are_redone_with_the_exact_scale():
    # the scale of a call is speculated from the previous call's peak and verified: a channel that gets 30 dB louder or quieter
    # from one buffer to the next leaves the grid window, and the call is redone (pass B) -- results, carried state and the
    # fused meter must come out as if nothing had happened.  Contract: the first two tiles of a call, whose
    # windows reach into the carried history, have their own scale: the START of the buffer right after a 30 dB drop is held to the
    # previous buffer's peak (it inherits state computed on the louder grid), but behind those tiles the buffer is served on its own
    # grid, and the buffer after it is within 1e-6 of its own peak from its first frame (tools/k2_soak.py found that it was not).
    ch, bf = 128, 1600
    st = design.config_stages("chain4")
    gpu, cpu = abi.Chain(ch, st, buffer_frames=bf, flags=abi.CHAIN_METER), orc.Chain(ch, st)
    levels = [1.0, 0.03, 0.03, 2.5, 1e-4, 1.0]
    refs = []
    prev_peak = np.zeros(ch)
    for step, lv in enumerate(levels):
        amp = np.where(np.arange(ch) % 2 == 0, lv, 1.0)        # odd channels stay put, even channels jump
        x = signal_input(bf, ch, seed=40 + step) * amp
        ref = cpu.process(x)
        refs.append(ref)
        y = gpu.process(x.astype(np.float32))
        assert gpu.last_path()[0] == 2
        peak = np.abs(ref).max(axis=0)
        err = np.abs(y - ref).max(axis=0)
        bar = REL_F32 * np.maximum(peak, prev_peak)
        assert (err <= bar).all(), f"step {step} (level {lv}): worst {np.max(err / np.maximum(peak, prev_peak)):.3e}"
        if step in (2, 5):   # a level that held for two buffers: the bar is the buffer's own peak again, from the first frame
            assert_parity(y, ref, REL_F32, f"step {step} (level {lv})")
        # behind the first two tiles (and the biquad's memory of them) even the buffer that carries the jump is on its own grid
        assert_parity(y[600:], ref[600:], REL_F32, f"step {step} (level {lv}), outputs 600..")
        prev_peak = peak
    peak, sumsq, frames = gpu.meter_read()
    rp, rs = orc.meter(np.concatenate(refs))
    assert frames == sum(len(r) for r in refs)
    np.testing.assert_allclose(peak, rp, rtol=2e-6)
    np.testing.assert_allclose(sumsq, rs, rtol=2e-6)


def test_k2_call_after_a_60_db_drop_starts_clean():
    # What tools/k2_soak.py found: a channel that is loud in one call and 60 dB down in the next two.  The call that carries the
    # drop starts from loud history (its first tiles are served on the loud grid, its error bar is the louder peak); with ONE scale
    # per call the whole call stayed on that grid and the state it left behind put 1e-4 of the quiet channel's peak into the first
    # outputs of the call after it.  Whole and partial last tiles, a batch and single buffers.
    ch = 256
    st = design.config_stages("chain4")
    for bf, nb in ((4096, 1), (4000, 1), (4096, 3)):
        gpu, cpu = abi.Chain(ch, st, buffer_frames=bf, max_batch=nb), orc.Chain(ch, st)
        quiet = np.where(np.arange(ch) % 4 == 1, 1e-3, 1.0)
        prev_peak = np.zeros(ch)
        for call, amp in enumerate((np.ones(ch), quiet, quiet, quiet)):
            total = bf * nb
            x = signal_input(total, ch, seed=70 + call) * amp
            ref = cpu.process(x, threads=os.cpu_count() or 1)
            d_in, d_out = abi.DeviceBuffer(total * ch * 4), abi.DeviceBuffer(total * ch * 4)
            d_in.upload(x.astype(np.float32))
            counts = gpu.process_batch_device(d_in.ptr, [bf] * nb, d_out.ptr, total)
            gpu.sync()
            assert gpu.last_path()[0] == 2 and sum(counts) == len(ref)
            y = d_out.download((len(ref), ch), np.float32)
            peak = np.abs(ref).max(axis=0)
            err = np.abs(y - ref).max(axis=0)
            if call == 1:    # the call that carries the drop: held to the louder of the two peaks
                assert (err <= REL_F32 * np.maximum(peak, prev_peak)).all(), f"bf {bf} x {nb}, call {call}"
            else:            # every other call, the two after the drop included: its own peak, from the first output frame
                assert_parity(y, ref, REL_F32, f"bf {bf} x {nb}, call {call}")
            prev_peak = peak
        gpu.close()


def test_k2_silent_and_tiny_channels():
    ch, bf = 128, 1600
    amp = np.ones(ch)
    amp[0::4] = 0.0          # digital silence
    amp[1::4] = 1e-20        # far below any sensible level (the scale is capped at 2^96: peaks down to ~1e-26 keep full accuracy)
    gpu, cpu = _k2_chain(ch, bf, 1), orc.Chain(ch, design.config_stages("chain4"))
    for step in range(2):
        x = signal_input(bf, ch, seed=50 + step) * amp
        ref = cpu.process(x)
        y = gpu.process(x.astype(np.float32))
        assert gpu.last_path()[0] == 2
        assert not np.any(y[:, 0::4])
        assert_parity(y[:, 1::4], ref[:, 1::4], REL_F32, f"tiny channels, step {step}")
        assert_parity(y[:, 2::4], ref[:, 2::4], REL_F32, f"step {step}")
        assert_parity(y[:, 3::4], ref[:, 3::4], REL_F32, f"step {step}")


def test_k2_full_batch_at_bench_size():
    # the launch bench.py times: 20 buffers x 4096 frames x 1024 ch in ONE process_batch_device call (4096 tiles, ~28 per CTA,
    # look-back chains hundreds of tiles deep), every buffer against the oracle; and the same batch again (carried state,
    # verified scales: no redo)
    ch, bf, nb = 1024, 4096, 20
    st = design.config_stages("chain4")
    x = signal_input(bf * nb, ch, seed=8)
    cpu = orc.Chain(ch, st)
    gpu = abi.Chain(ch, st, buffer_frames=bf, max_batch=nb)
    xf = x.astype(np.float32)
    d_in, d_out = abi.DeviceBuffer(xf.nbytes), abi.DeviceBuffer(xf.nbytes)
    d_in.upload(xf)
    for rep in range(2):
        refs = [cpu.process(x[b * bf:(b + 1) * bf], threads=os.cpu_count() or 1) for b in range(nb)]
        counts = gpu.process_batch_device(d_in.ptr, [bf] * nb, d_out.ptr, len(x))
        gpu.sync()
        assert counts == [len(r) for r in refs] and gpu.last_path()[0] == 2
        y = d_out.download((sum(counts), ch), np.float32)
        pos = 0
        for b, r in enumerate(refs):
            assert_parity(y[pos:pos + len(r)], r, REL_F32, f"batch {rep} buffer {b}")
            pos += len(r)
    d_in.free()
    d_out.free()


def test_k2_chains_of_several_lines_share_one_device():
    # several Lines on one device: three K2 chains launched round-robin on three streams without a sync in between, so their
    # persistent grids compete for the SMs (K2's look-back waits on tiles of CTAs that may not be resident yet: it has to make
    # progress, not run into its bound).  Same stages and input for every chain, so one oracle run checks all of them.
    import torch
    ch, bf, nb, rounds, n_chains = 1024, 4096, 6, 3, 3
    st = design.config_stages("chain4")
    x = signal_input(bf * nb * rounds, ch, seed=11)
    cpu = orc.Chain(ch, st)
    refs = [cpu.process(x[i * bf:(i + 1) * bf], threads=os.cpu_count() or 1) for i in range(nb * rounds)]
    xf = x.astype(np.float32)
    d_in = abi.DeviceBuffer(xf.nbytes)
    d_in.upload(xf)
    per_round = bf * nb * ch * 4
    chains = [abi.Chain(ch, st, buffer_frames=bf, max_batch=nb) for _ in range(n_chains)]
    outs = [abi.DeviceBuffer(xf.nbytes) for _ in range(n_chains)]
    streams = [torch.cuda.Stream(device=0) for _ in range(n_chains)]
    counts = [[] for _ in range(n_chains)]
    for r in range(rounds):
        for k, g in enumerate(chains):
            done = sum(counts[k])
            counts[k] += g.process_batch_device(d_in.ptr + r * per_round, [bf] * nb, outs[k].ptr + done * ch * 4,
                                                bf * nb * rounds - done, stream=streams[k].cuda_stream)
    for k, g in enumerate(chains):
        g.sync(streams[k].cuda_stream)
        assert counts[k] == [len(r) for r in refs] and g.last_path()[0] == 2
        y = outs[k].download((sum(counts[k]), ch), np.float32)
        pos = 0
        for b, r in enumerate(refs):
            assert_parity(y[pos:pos + len(r)], r, REL_F32, f"chain {k} buffer {b}")
            pos += len(r)
    d_in.free()
    for o in outs:
        o.free()


@pytest.mark.parametrize("cfg,channels,dtype", [("chain4", 1024, np.float32), ("chain4", 512, np.float64),
                                                ("gain_biquad", 1024, np.float32)])
def test_large_host_buffers_go_through_in_pieces(cfg, channels, dtype):
    # pb_chain_process cuts one host buffer of >= 8 MiB into four pieces (multiples of 160 frames) that follow each other through
    # H2D, kernels and D2H; to the chain every piece is a call of its own.  All three kernel families, pageable buffers (numpy) and
    # a pinned pair, a short last buffer; frame counts bit-exact, every buffer against the oracle.
    bf = 4096
    st = design.config_stages(cfg)
    gpu, cpu = abi.Chain(channels, st, buffer_frames=bf, dtype=dtype), orc.Chain(channels, st)
    rel = REL_F32 if dtype == np.float32 else REL_F64
    sizes = [bf, bf, bf, 3000]
    x = signal_input(sum(sizes), channels, seed=21)
    item = np.dtype(dtype).itemsize
    pin_in, pin_out = abi.PinnedBuffer(bf * channels * item), abi.PinnedBuffer(bf * channels * item)
    pos, _, l0 = 0, *gpu.last_path()
    for b, n in enumerate(sizes):
        blk = x[pos:pos + n]
        pos += n
        ref = cpu.process(blk, threads=os.cpu_count() or 1)
        if b == 2:
            # a call whose output does not fit is rejected before anything is enqueued: the carried state is untouched and the
            # same buffer, offered again with room, gives the oracle's result
            pin_in.array((n, channels), dtype)[:] = blk
            got = abi._i64()
            rc = abi.lib().pb_chain_process(gpu._h, pin_in.ptr, n, pin_out.ptr, len(ref) - 1, abi.C.byref(got))
            assert rc == abi.PB_ERR_CAPACITY
        if b % 2 == 0:
            y = gpu.process(blk.astype(dtype))                       # pageable in and out
        else:
            pin_in.array((n, channels), dtype)[:] = blk
            got = abi._i64()
            abi.check(abi.lib().pb_chain_process(gpu._h, pin_in.ptr, n, pin_out.ptr, bf, abi.C.byref(got)))
            y = pin_out.array((bf, channels), dtype)[:got.value].copy()
        assert len(y) == len(ref)
        assert_parity(y, ref, rel, f"buffer {b} ({n} frames)")
    path, l1 = gpu.last_path()
    assert l1 - l0 >= 4 * len(sizes)                                   # at least one launch per piece
    pin_in.free()
    pin_out.free()


# ------------------------------------------- K3: streaming kernels (runs without FIR and resampler) --

@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("channels,sizes", [(64, [4096] * 3), (1, [512, 512, 16]), (33, [1, 127, 128, 129, 0, 2047, 5]),
                                            (100, [300, 1000])])
def test_k3_gain_biquad_stream_path(dtype, channels, sizes):
    stages = design.config_stages("gain_biquad") + [{"kind": "gain", "gain": 1.5}]
    bf = max(sizes)
    gpu, cpu = abi.Chain(channels, stages, buffer_frames=bf, dtype=dtype), orc.Chain(channels, stages)
    x = signal_input(sum(sizes), channels, seed=3)
    pos, run_peak = 0, np.zeros(channels)
    for i, n in enumerate(sizes):
        blk = x[pos:pos + n]
        pos += n
        ref = cpu.process(blk)
        if len(ref):
            run_peak = np.maximum(run_peak, np.abs(ref).max(axis=0))
        y = gpu.process(blk.astype(dtype))
        assert_parity(y, ref, REL_F32 if dtype == np.float32 else REL_F64, f"buffer {i} ({n} frames)", floor=run_peak)
        if n:
            assert gpu.last_path()[0] == 3


@pytest.mark.parametrize("kind,f0,q", [("lowpass", 8000.0, 0.9), ("highpass", 20.0, 0.707), ("peaking", 1000.0, 4.0)])
def test_k3_long_batch_slides_the_look_back_window(kind, f0, q):
    # 8 channels are ONE look-back group: a batch of 2048 tiles keeps hundreds of tiles of that group in flight, so tiles
    # resolve their incoming state over several windows of 32 aggregates (the 20 Hz high-pass has poles at |z| = 0.998:
    # its state decays over thousands of frames, nothing may be dropped)
    ch, bf, nb = 8, 4096, 64
    b, a = design.biquad(kind, f0, 48000.0, q=q, gain_db=6.0)
    stages = [{"kind": "gain", "gain": 0.7}, {"kind": "biquad", "b": b, "a": a}]
    x = signal_input(bf * nb, ch, seed=5)
    cpu = orc.Chain(ch, stages)
    ref = cpu.process(x)
    gpu = abi.Chain(ch, stages, buffer_frames=bf, max_batch=nb)
    xf = x.astype(np.float32)
    d_in, d_out = abi.DeviceBuffer(xf.nbytes), abi.DeviceBuffer(xf.nbytes)
    d_in.upload(xf)
    for rep in range(2):  # the second launch reuses the look-back arrays under a new epoch
        counts = gpu.process_batch_device(d_in.ptr, [bf] * nb, d_out.ptr, len(x))
        gpu.sync()
        assert counts == [bf] * nb and gpu.last_path()[0] == 3
        y = d_out.download((len(x), ch), np.float32)
        if rep == 0:
            assert_parity(y, ref, REL_F32, "batch 0")
        else:
            assert_parity(y, cpu.process(x), REL_F32, "batch 1 (carried state)")


def test_k3_and_k1_agree_and_share_state_layout():
    # the same run on the generic tile kernel (PB_CHAIN_NO_STREAM) and on the streaming kernel, both against the oracle
    stages = design.config_stages("gain_biquad")
    x = signal_input(3000, 48, seed=11)
    ref = orc.Chain(48, stages).process(x)
    g1 = abi.Chain(48, stages, buffer_frames=3000, flags=abi.CHAIN_NO_STREAM)
    g3 = abi.Chain(48, stages, buffer_frames=3000)
    y1, y3 = g1.process(x.astype(np.float32)), g3.process(x.astype(np.float32))
    assert g1.last_path()[0] == 1 and g3.last_path()[0] == 3
    assert_parity(y1, ref, REL_F32, "K1")
    assert_parity(y3, ref, REL_F32, "K3")


def test_k3_copy_and_gain_runs():
    # copy is bit-exact (configs[0]); a gain-only run is one rounding (exact against the float32 product)
    for dtype, ch, n in ((np.float64, 2, 512 * 5 + 3), (np.float32, 7, 1001), (np.float32, 1024, 4096)):
        x = signal_input(n, ch, seed=2).astype(dtype)
        g = abi.Chain(ch, [{"kind": "copy"}], buffer_frames=n, dtype=dtype)
        assert np.array_equal(g.process(x), x) and g.last_path()[0] == 3
        g = abi.Chain(ch, [{"kind": "gain", "gain": 0.3}, {"kind": "gain", "gain": 1.7}], buffer_frames=n, dtype=dtype)
        y = g.process(x)
        ref = orc.Chain(ch, [{"kind": "gain", "gain": 0.3}, {"kind": "gain", "gain": 1.7}]).process(x.astype(np.float64))
        assert_parity(y, ref, REL_F32 if dtype == np.float32 else REL_F64, "gain run")


def test_k3_meter_and_mutation():
    b, a = design.biquad("lowpass", 8000.0, 48000.0, q=0.9)
    b2, a2 = design.biquad("highpass", 300.0, 48000.0, q=0.7)
    stages = [{"kind": "gain", "gain": 0.5}, {"kind": "biquad", "b": b, "a": a}]
    gpu, cpu = abi.Chain(40, stages, buffer_frames=1000, flags=abi.CHAIN_METER), orc.Chain(40, stages)
    x = signal_input(3000, 40, seed=4)
    refs = []
    for i in range(3):
        if i == 1:
            gpu.set_stage(1, {"kind": "biquad", "b": b2, "a": a2})
            cpu.set_stage(1, {"kind": "biquad", "b": b2, "a": a2})
        if i == 2:
            gpu.set_stage(0, {"kind": "gain", "gain": 0.9})
            cpu.set_stage(0, {"kind": "gain", "gain": 0.9})
        blk = x[i * 1000:(i + 1) * 1000]
        refs.append(cpu.process(blk))
        assert_parity(gpu.process(blk.astype(np.float32)), refs[-1], REL_F32, f"buffer {i}")
        assert gpu.last_path()[0] == 3
    peak, sumsq, frames = gpu.meter_read()
    rp, rs = orc.meter(np.concatenate(refs))
    assert frames == 3000
    np.testing.assert_allclose(peak, rp, rtol=2e-6)
    np.testing.assert_allclose(sumsq, rs, rtol=2e-6)
    # meter on a gain-only run goes through the channel-structured kernel too
    g = abi.Chain(5, [{"kind": "gain", "gain": 2.0}], buffer_frames=300, flags=abi.CHAIN_METER)
    y = g.process(x[:300, :5].astype(np.float32))
    pk, sq, fr = g.meter_read()
    assert fr == 300 and np.array_equal(pk, np.abs(y.astype(np.float64)).max(axis=0))


@pytest.mark.parametrize("channels,nb", [(64, 320), (1024, 20)])
def test_k3_full_batch_at_baseline_size(channels, nb):
    # configs[1] at the size bench.py / tools/hbm_chains.py time it (320 buffers x 4096 frames x 64 ch in ONE launch: 5120 tiles
    # per channel group, hundreds of them in flight) and the same run at 1024 ch: every buffer against the oracle, and
    # linearity -- an input scaled by a power of two gives the scaled output: no rounding depends on the exponent, only the
    # order in which a tile folds its predecessors may differ between two launches (1e-16 in double, so at most a handful
    # of the 84 M float32 results may land on the other side of a rounding boundary)
    bf = 4096
    stages = design.config_stages("gain_biquad")
    x = signal_input(bf * nb, channels, seed=8)
    ref = orc.Chain(channels, stages).process(x, threads=os.cpu_count() or 1)
    xf = x.astype(np.float32)
    gpu = abi.Chain(channels, stages, buffer_frames=bf, max_batch=nb)
    d_in, d_out = abi.DeviceBuffer(xf.nbytes), abi.DeviceBuffer(xf.nbytes)
    d_in.upload(xf)
    counts = gpu.process_batch_device(d_in.ptr, [bf] * nb, d_out.ptr, len(x))
    gpu.sync()
    assert counts == [bf] * nb and gpu.last_path()[0] == 3
    y = d_out.download((len(x), channels), np.float32)
    for b in range(nb):
        assert_parity(y[b * bf:(b + 1) * bf], ref[b * bf:(b + 1) * bf], REL_F32, f"buffer {b}")
    gpu.reset()
    d_in.upload(xf * np.float32(0.25))
    gpu.process_batch_device(d_in.ptr, [bf] * nb, d_out.ptr, len(x))
    gpu.sync()
    y2 = d_out.download((len(x), channels), np.float32)
    diff = y2 != y * np.float32(0.25)
    assert int(diff.sum()) <= 16, f"{int(diff.sum())} values differ"
    assert float(np.abs(y2 - y * np.float32(0.25)).max()) <= 2.0 ** -24 * float(np.abs(y).max())
    d_in.free()
    d_out.free()


@pytest.mark.parametrize("dtype,channels", [(np.float32, 64), (np.float64, 64), (np.float32, 40), (np.float32, 100)])
def test_k3_two_sweeps_with_few_channel_groups(dtype, channels):
    # At most 4 channel groups (<= 128 ch) and >= 256 tiles per launch: the run takes two sweeps with a scan in between instead
    # of the look-back (chain_stream.cuh, kStOneSweep).  Ragged last buffer (its tile is not chained), three launches in a row
    # (the carried state crosses launches), the fused meter, a channel count that is not a multiple of 32, both dtypes.
    bf = 4096
    sizes = [bf] * 33 + [1234]      # f32: 528 full tiles + a ragged one; f64: 1056 + 1
    b, a = design.biquad("highpass", 20.0, 48000.0, q=0.707)
    stages = [{"kind": "gain", "gain": 0.8}, {"kind": "biquad", "b": b, "a": a}, {"kind": "gain", "gain": 1.25}]
    gpu = abi.Chain(channels, stages, buffer_frames=bf, max_batch=len(sizes), dtype=dtype, flags=abi.CHAIN_METER)
    cpu = orc.Chain(channels, stages)
    total = sum(sizes)
    el = np.dtype(dtype).itemsize
    d_in, d_out = abi.DeviceBuffer(total * channels * el), abi.DeviceBuffer(total * channels * el)
    refs = []
    for rep in range(3):
        x = signal_input(total, channels, seed=20 + rep)
        ref = cpu.process(x, threads=os.cpu_count() or 1)
        refs.append(ref)
        d_in.upload(x.astype(dtype))
        counts = gpu.process_batch_device(d_in.ptr, sizes, d_out.ptr, total)
        gpu.sync()
        assert counts == sizes and gpu.last_path() == (3, 3 + 3 * rep)   # aggregate sweep, scan, apply sweep
        y = d_out.download((total, channels), dtype)
        pos = 0
        for i, n in enumerate(sizes):
            # float64: the 20 Hz high-pass has poles 0.002 from the unit circle; in the TDF-II basis the powers A^k that carry a
            # state across 128 rows have entries of order 1e2 and the states they produce are 1e-3 of the signal, so the
            # time-parallel form loses ~5 digits to cancellation that the sequential oracle does not: 7e-11 measured in two sweeps,
            # 9e-11 in one (tools/k3_f64_cond.py; 1e-13 for a 200 Hz high-pass, 5e-16 for an 8 kHz low-pass).  Well-conditioned filters hold 1e-12 (test_k3_gain_biquad_stream_path).
            assert_parity(y[pos:pos + n], ref[pos:pos + n], REL_F32 if dtype == np.float32 else 2e-10, f"launch {rep} buffer {i}")
            pos += n
    peak, sumsq, frames = gpu.meter_read()
    rp, rs = orc.meter(np.concatenate(refs))
    assert frames == 3 * total
    np.testing.assert_allclose(peak, rp, rtol=2e-6)
    np.testing.assert_allclose(sumsq, rs, rtol=2e-6)
    d_in.free()
    d_out.free()


def test_k2_serves_one_4096_frame_buffer_per_call_in_one_launch():
    # bufferSize 4096 is not a multiple of K2's 160-frame tile and the resampler phase at the start of a buffer returns to 0 only
    # every fifth buffer.  The tiles of a call start at its first frame with the tables of the phase found there (160 frames give
    # 147 outputs from any phase) and the last tile is partial (96 frames: rows behind them zero-filled, outputs masked, carried
    # state taken at the last real row): every call is ONE K2 launch plus its verify launch, no K1 head or tail.  Bit-exact
    # frame counts (3763, 3763, 3763, 3763, 3764, ...) and the 1e-6 bar on every buffer.
    ch, bf = 128, 4096
    st = design.config_stages("chain4")
    gpu, cpu = abi.Chain(ch, st, buffer_frames=bf), orc.Chain(ch, st)
    lens, launches = [], 0
    for b in range(11):
        x = signal_input(bf, ch, seed=30 + b)
        ref = cpu.process(x)
        y = gpu.process(x.astype(np.float32))
        lens.append(len(y))
        path, n_launch = gpu.last_path()
        assert path == 2 and n_launch - launches == 2, (path, n_launch - launches)
        launches = n_launch
        assert_parity(y, ref, REL_F32, f"buffer {b}")
    assert lens == [3763, 3763, 3763, 3763, 3764, 3763, 3763, 3763, 3763, 3764, 3763]


@pytest.mark.parametrize("channels", [128, 384])
def test_k2_partial_last_tile_at_every_kind_of_boundary(channels):
    # Call lengths that leave 1, 15, 16, 17, 159 ... frames in the last tile, end exactly on a tile boundary, are shorter than one
    # tile (K1 serves those, from the state K2 left, and K2 continues from K1's), with the fused meter: every call against
    # the oracle, frame counts bit-exact, and the meter over the whole stream.
    st = design.config_stages("chain4")
    sizes = [161, 175, 176, 177, 319, 320, 4096, 159, 1, 160, 2000, 4095, 33, 1600, 4000, 4001]
    bf = max(sizes)
    gpu, cpu = abi.Chain(channels, st, buffer_frames=bf, flags=abi.CHAIN_METER), orc.Chain(channels, st)
    x = signal_input(sum(sizes), channels, seed=77)
    pos, refs, run_peak = 0, [], np.zeros(channels)
    for i, n in enumerate(sizes):
        blk = x[pos:pos + n]
        pos += n
        ref = cpu.process(blk)
        refs.append(ref)
        if len(ref):
            run_peak = np.maximum(run_peak, np.abs(ref).max(axis=0))
        assert gpu.peek_out_frames(n) == len(ref)
        y = gpu.process(blk.astype(np.float32))
        assert len(y) == len(ref), f"call {i} ({n} frames): {len(y)} vs {len(ref)} frames"
        assert gpu.last_path()[0] == (2 if n >= 160 else 1), (n, gpu.last_path())
        assert_parity(y, ref, REL_F32, f"call {i} ({n} frames)", floor=run_peak)
    peak, sumsq, frames = gpu.meter_read()
    rp, rs = orc.meter(np.concatenate(refs))
    assert frames == sum(len(r) for r in refs)
    np.testing.assert_allclose(peak, rp, rtol=2e-6)
    np.testing.assert_allclose(sumsq, rs, rtol=2e-6)


# ------------------------------------------------ graph edits: InsertProcessor on a fused run (pipe.go:297-333) --

def _edit_case(channels, bf, stages, edits, dtype=np.float32, n_buffers=6, rel=REL_F32):
    """edits: {buffer index: (position, stage)} applied BEFORE that buffer, on the GPU chain and on the oracle's stage list."""
    gpu = abi.Chain(channels, stages, buffer_frames=bf, dtype=dtype)
    cpu = orc.StageList(channels, stages)
    paths, run_peak = [], np.zeros(channels)
    for b in range(n_buffers):
        if b in edits:
            pos, st = edits[b]
            gpu.insert_stage(pos, st)
            cpu.insert(pos, st)
        x = signal_input(bf, channels, seed=60 + b)
        ref = cpu.process(x)
        run_peak = np.maximum(run_peak, np.abs(ref).max(axis=0)) if len(ref) else run_peak
        assert gpu.peek_out_frames(bf) == len(ref)
        y = gpu.process(x.astype(dtype))
        assert_parity(y, ref, rel, f"buffer {b}", floor=run_peak)
        paths.append(gpu.last_path()[0])
    return paths


def test_insert_processor_keeps_the_state_of_existing_stages():
    # a gain, a second biquad and a FIR spliced into a running [gain, biquad] run: the run is re-planned (one segment, then two),
    # the biquad that was there keeps its state across both edits, the new stages start from zero state
    b2, a2 = design.biquad("highpass", 300.0, 48000.0, q=0.7)
    stages = design.config_stages("gain_biquad")
    edits = {2: (2, {"kind": "biquad", "b": b2, "a": a2}),                       # behind the existing biquad: a second segment
             4: (0, {"kind": "fir", "taps": design.lowpass_fir(65, 0.2)})}        # in front of everything
    paths = _edit_case(48, 1000, stages, edits)
    assert paths[0] == 3       # the streaming kernel first; afterwards [FIR, biquad] on the generic kernel + [biquad] streaming


def test_insert_processor_into_the_headline_chain_keeps_it_on_the_tensor_kernel():
    # the 4-stage chain on K2; a gain in front keeps the shape (still K2, state carried through the re-plan); a second FIR in
    # front of the first cuts the run into [gain, gain, FIR] on the generic kernel and [FIR, biquad, resample], which is still
    # K2's shape and continues there with the state it had
    ch, bf = 128, 1600
    stages = design.config_stages("chain4")
    edits = {2: (0, {"kind": "gain", "gain": 1.5}),
             4: (2, {"kind": "fir", "taps": design.lowpass_fir(33, 0.3)})}
    paths = _edit_case(ch, bf, stages, edits)
    assert paths == [2] * 6


def test_insert_processor_validation_and_float64():
    gpu = abi.Chain(4, [{"kind": "gain", "gain": 2.0}], buffer_frames=64, dtype=np.float64)
    with pytest.raises(abi.PipeB200Error) as e:
        gpu.insert_stage(5, {"kind": "gain", "gain": 1.0})
    assert e.value.code == abi.PB_ERR_INVALID
    with pytest.raises(abi.PipeB200Error) as e:
        gpu.insert_stage(0, {"kind": "resample", "up": 3, "down": 2, "taps": np.ones(6)})   # rejected: the run is unchanged
    assert e.value.code == abi.PB_ERR_UNSUPPORTED
    x = signal_input(64, 4)
    assert np.array_equal(gpu.process(x), 2.0 * x)
    b, a = design.biquad("lowpass", 5000.0, 48000.0)
    paths = _edit_case(4, 256, [{"kind": "gain", "gain": 0.5}], {1: (1, {"kind": "biquad", "b": b, "a": a}),
                                                               3: (2, {"kind": "resample", "up": 2, "down": 3,
                                                                       "taps": design.resampler_prototype(2, 3, 8)})},
                       dtype=np.float64, rel=REL_F64)
    assert paths[-1] == 1
