"""Randomised parity soak of the generic fused tile kernel (K1) against the CPU oracle: random stage lists (gains, FIR of 2..257 taps,
biquads, resamplers 147/160, 1/2, 2/3, 3/4, 80/147 -- one or two fused segments), 1..130 channels, f32 and f64, ragged call
lengths, single buffers and batches, levels that jump between calls, the fused meter.  PB_CHAIN_NO_TENSOR | PB_CHAIN_NO_STREAM keep
every run on K1.  Run on a GPU box:    python tools/k1_soak.py [iterations] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle as orc  # noqa: E402
from pipe_b200 import abi, design  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
worst32, worst64, n_calls, n_fail = 0.0, 0.0, 0, 0


def random_segment():
    st = []
    if rng.random() < 0.6:
        st.append({"kind": "gain", "gain": float(rng.uniform(0.3, 2.0))})
    if rng.random() < 0.6:
        st.append({"kind": "fir", "taps": design.lowpass_fir(int(rng.choice([2, 17, 64, 129, 257])), float(rng.uniform(0.2, 0.45)))})
    if rng.random() < 0.6:
        kind = str(rng.choice(["highpass", "lowpass", "peaking"]))
        # (no narrow low-passes: a chain that removes 40+ dB of a broadband signal has its float32 error relative to the level INSIDE
        # the chain -- buffers and histories between the stages are float32 --, not to its output: DESIGN.md section 3)
        f0 = float(rng.choice([8000.0, 15000.0])) if kind == "lowpass" else float(rng.choice([100.0, 1000.0, 8000.0, 15000.0]))
        b, a = design.biquad(kind, f0, 48000.0, q=float(rng.uniform(0.5, 2.0)), gain_db=float(rng.uniform(-6, 6)))
        st.append({"kind": "biquad", "b": b, "a": a})
    if rng.random() < 0.5:
        up, down = [(147, 160), (1, 2), (2, 3), (3, 4), (80, 147)][int(rng.integers(0, 5))]   # (up <= down: a Processor cannot emit more than bufferSize, pipe.go:437-443)
        tpp = int(rng.choice([8, 16]))
        st.append({"kind": "resample", "up": up, "down": down, "taps": design.resampler_prototype(up, down, tpp)})
    if not st or rng.random() < 0.3:
        st.append({"kind": "gain", "gain": float(rng.uniform(0.5, 1.5))})
    return st


for it in range(iters):
    ch = int(rng.choice([1, 3, 16, 32, 33, 64, 100, 130]))
    dtype = np.float32 if rng.random() < 0.6 else np.float64
    two_segments = rng.random() < 0.3
    stages = random_segment() + (random_segment() if two_segments else [])
    bf = int(rng.choice([256, 1000, 4096]))
    nb = int(rng.choice([1, 1, 4]))
    meter = bool(rng.integers(0, 2))
    flags = abi.CHAIN_NO_TENSOR | abi.CHAIN_NO_STREAM | (abi.CHAIN_METER if meter else 0)
    gpu = abi.Chain(ch, stages, buffer_frames=bf, max_batch=nb, dtype=dtype, flags=flags)
    cpu = orc.Chain(ch, stages)
    el = np.dtype(dtype).itemsize
    levels = 10.0 ** (-rng.integers(0, 3, size=ch) * 1.0)
    run_peak = np.zeros(ch)
    for call in range(int(rng.integers(2, 4))):
        if rng.random() < 0.3:
            levels = 10.0 ** (-rng.integers(0, 3, size=ch) * 1.0)
        sizes = [bf] * (nb - 1) + [int(rng.integers(1, bf + 1)) if rng.random() < 0.5 else bf]
        total = sum(sizes)
        x = orc.source_fill(int(rng.integers(0, 1 << 30)), total * ch).reshape(total, ch) * levels
        ref = cpu.process(x, threads=os.cpu_count() or 1)
        cap = max(len(ref), 1) + 8
        d_in, d_out = abi.DeviceBuffer(total * ch * el), abi.DeviceBuffer(4 * bf * nb * ch * el + 64)
        d_in.upload(x.astype(dtype))
        counts = gpu.process_batch_device(d_in.ptr, sizes, d_out.ptr, 4 * bf * nb)
        gpu.sync()
        assert sum(counts) == len(ref), (counts, len(ref), stages)
        n_calls += 1
        if len(ref) == 0:
            continue
        y = d_out.download((len(ref), ch), dtype).astype(np.float64)
        run_peak = np.maximum(run_peak, np.abs(ref).max(axis=0))
        pk = np.abs(ref).max(axis=0) if len(ref) >= 256 else run_peak   # (very short outputs: the peak of the stream so far)
        # (channels whose output has not arrived yet -- a first call shorter than the filters' delay -- are held to 1 % of the input peak)
        pk = np.maximum(pk, 1e-2 * np.abs(x).max(axis=0))
        err = float((np.abs(y - ref).max(axis=0) / np.maximum(pk, 1e-300)).max())
        # two fused segments hand over a buffer of the chain's dtype: two roundings where the oracle has none (measured 1.5e-6 in f32)
        bar = (2e-6 if two_segments else 1e-6) if dtype == np.float32 else 1e-9
        if dtype == np.float32:
            worst32 = max(worst32, err)
        else:
            worst64 = max(worst64, err)
        if os.environ.get("SOAK_VERBOSE_IT") == str(it):
            e = np.abs(y - ref)
            cw = int(np.argmax(e.max(axis=0) / np.maximum(pk, 1e-300)))
            fw = int(np.argmax(e[:, cw]))
            print(f"  it {it} call {call}: sizes {sizes} -> {counts} out, levels {levels}, err/peak {err:.3e}, worst channel {cw} frame {fw}: y {y[fw, cw]:.6e} ref {ref[fw, cw]:.6e}, "
                  f"peak {pk[cw]:.3e}, path {gpu.last_path()}", flush=True)
        if err > bar:
            n_fail += 1
            kinds = [s["kind"] + (f"({len(s['taps'])})" if s["kind"] == "fir" else f"({s['up']}/{s['down']})" if s["kind"] == "resample" else "") for s in stages]
            print(f"FAIL iteration {it} call {call}: {ch} ch {np.dtype(dtype).name}, {kinds}, bf {bf} x {nb}, last {sizes[-1]}, path {gpu.last_path()}: err/peak {err:.3e}", flush=True)
    gpu.close()
print(f"k1 soak: {iters} chains, {n_calls} calls, worst err / own peak f32 {worst32:.3e}, f64 {worst64:.3e}: {'ok' if n_fail == 0 else str(n_fail) + ' FAILED'}", flush=True)
sys.exit(1 if n_fail else 0)
