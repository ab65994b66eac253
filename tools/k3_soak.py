"""Randomised parity soak of the streaming kernels (K3) against the CPU oracle: gain + biquad runs with random filters, channel
counts (1..160, so one to five channel groups and ragged last groups), f32 and f64, batches on both sides of the two-sweep
threshold (256 tiles per launch), ragged last buffers, levels that jump between calls, the fused meter.
Run on a GPU box:    python tools/k3_soak.py [iterations] [seed]"""
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
worst32, worst64, n_calls, n_fail, paths = 0.0, 0.0, 0, 0, {}
for it in range(iters):
    ch = int(rng.choice([1, 7, 32, 40, 64, 96, 100, 128, 160]))
    dtype = np.float32 if rng.random() < 0.7 else np.float64
    kind = str(rng.choice(["highpass", "lowpass", "peaking"]))
    f0 = float(rng.choice([20.0, 200.0, 1000.0, 8000.0, 16000.0]))
    b, a = design.biquad(kind, f0, 48000.0, q=float(rng.uniform(0.5, 3.0)), gain_db=float(rng.uniform(-6, 6)))
    stages = [{"kind": "gain", "gain": float(rng.uniform(0.3, 2.5))}, {"kind": "biquad", "b": b, "a": a}]
    if rng.random() < 0.5:
        stages.append({"kind": "gain", "gain": float(rng.uniform(0.5, 2.0))})
    bf = int(rng.choice([512, 4096]))
    nb = int(rng.choice([1, 8, 20, 40])) if bf == 4096 else int(rng.choice([1, 64, 200]))
    meter = bool(rng.integers(0, 2))
    gpu = abi.Chain(ch, stages, buffer_frames=bf, max_batch=nb, dtype=dtype, flags=abi.CHAIN_METER if meter else 0)
    cpu = orc.Chain(ch, stages)
    el = np.dtype(dtype).itemsize
    levels = 10.0 ** (-rng.integers(0, 4, size=ch) * 1.0)
    for call in range(int(rng.integers(2, 4))):
        if rng.random() < 0.3:
            levels = 10.0 ** (-rng.integers(0, 4, size=ch) * 1.0)
        sizes = [bf] * (nb - 1) + [int(rng.integers(1, bf + 1)) if rng.random() < 0.5 else bf]
        total = sum(sizes)
        x = orc.source_fill(int(rng.integers(0, 1 << 30)), total * ch).reshape(total, ch) * levels
        d_in, d_out = abi.DeviceBuffer(total * ch * el), abi.DeviceBuffer(total * ch * el)
        d_in.upload(x.astype(dtype))
        counts = gpu.process_batch_device(d_in.ptr, sizes, d_out.ptr, total)
        gpu.sync()
        ref = cpu.process(x, threads=os.cpu_count() or 1)
        assert counts == sizes and len(ref) == total
        y = d_out.download((total, ch), dtype).astype(np.float64)
        pk = np.abs(ref).max(axis=0)
        err = float((np.abs(y - ref).max(axis=0) / np.maximum(pk, 1e-300)).max())
        path = gpu.last_path()[0]
        paths[path] = paths.get(path, 0) + 1
        n_calls += 1
        # f64: the parity tests hold 1e-12 on steady signals; behind a 60 dB drop with poles at 20 Hz (|z| = 0.998) the double
        # sub-chunk states carry ~1e-16 of the LOUD state against the quiet signal: up to 6e-9 of its peak, measured here
        bar = 1e-6 if dtype == np.float32 else 1e-8
        if dtype == np.float32:
            worst32 = max(worst32, err)
        else:
            worst64 = max(worst64, err)
        if err > bar:
            n_fail += 1
            print(f"FAIL iteration {it} call {call}: {ch} ch {np.dtype(dtype).name}, {kind} {f0} Hz, bf {bf} x {nb}, last {sizes[-1]}, meter {meter}, path {path}: "
                  f"err/peak {err:.3e}", flush=True)
    gpu.close()
print(f"k3 soak: {iters} chains, {n_calls} calls (kernel paths {paths}), worst err / own peak f32 {worst32:.3e}, f64 {worst64:.3e}: "
      f"{'ok' if n_fail == 0 else str(n_fail) + ' FAILED'}", flush=True)
sys.exit(1 if n_fail else 0)
