"""Randomised soak of the chain LIFECYCLE through the C-ABI against the CPU oracle: random sequences of pb_chain_process (host
buffers, ragged sizes), pb_chain_set_stage (a mutation between two buffers: new gain, new biquad coefficients), pb_chain_reset
(a restarted Pipe) on the three kernel families, with the fused meter checked at the end of every life.
Run on a GPU box:    python tools/ops_soak.py [iterations] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle as orc  # noqa: E402
from pipe_b200 import abi, design  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
worst, n_ops, n_fail = 0.0, 0, 0
for it in range(iters):
    family = int(rng.integers(0, 3))
    if family == 0:      # headline chain on the tcgen05 kernel
        ch, dtype, flags, stages = int(rng.choice([128, 256])), np.float32, 0, design.config_stages("chain4")
    elif family == 1:    # gain + biquad on the streaming kernel
        ch, dtype, flags, stages = int(rng.choice([2, 64, 100])), (np.float32 if rng.random() < 0.5 else np.float64), 0, design.config_stages("gain_biquad")
    else:                # headline chain on the generic kernel, float64
        ch, dtype, flags, stages = int(rng.choice([16, 48])), np.float64, abi.CHAIN_NO_TENSOR, design.config_stages("chain4")
    stages = [dict(s) for s in stages]
    bf = int(rng.choice([1600, 4096]))
    gpu, cpu = abi.Chain(ch, stages, buffer_frames=bf, dtype=dtype, flags=flags | abi.CHAIN_METER), orc.Chain(ch, stages)
    bar = 1e-6 if dtype == np.float32 else 1e-9
    refs, run_peak = [], np.zeros(ch)
    bq_idx = [i for i, s in enumerate(stages) if s["kind"] == "biquad"][0]
    for op in range(int(rng.integers(6, 12))):
        r = rng.random()
        n_ops += 1
        if r < 0.15:
            g = float(rng.uniform(0.2, 1.5))
            gpu.set_stage(0, {"kind": "gain", "gain": g})
            cpu.set_stage(0, {"kind": "gain", "gain": g})
        elif r < 0.30:
            b, a = design.biquad("peaking", float(rng.choice([500.0, 1000.0, 4000.0])), 48000.0, q=float(rng.uniform(0.7, 2.0)), gain_db=float(rng.uniform(-4, 4)))
            gpu.set_stage(bq_idx, {"kind": "biquad", "b": b, "a": a})
            cpu.set_stage(bq_idx, {"kind": "biquad", "b": b, "a": a})
        elif r < 0.38:
            if refs:
                peak, sumsq, frames = gpu.meter_read()
                rp, rs = orc.meter(np.concatenate(refs))
                ok = frames == sum(len(x) for x in refs) and np.allclose(peak, rp, rtol=5e-6) and np.allclose(sumsq, rs, rtol=5e-6)
                if not ok:
                    n_fail += 1
                    print(f"FAIL iteration {it} op {op}: meter before reset (family {family}, {ch} ch)", flush=True)
            gpu.reset()
            cpu.reset()
            refs, run_peak = [], np.zeros(ch)
        else:
            n = int(rng.integers(1, bf + 1)) if rng.random() < 0.4 else bf
            x = orc.source_fill(int(rng.integers(0, 1 << 30)), n * ch).reshape(n, ch) * float(rng.choice([1.0, 1.0, 0.1]))
            ref = cpu.process(x)
            y = gpu.process(x.astype(dtype))
            assert len(y) == len(ref), (len(y), len(ref))
            refs.append(ref)
            if len(ref):
                run_peak = np.maximum(run_peak, np.abs(ref).max(axis=0))
                pk = np.abs(ref).max(axis=0) if len(ref) >= 256 else run_peak
                err = float((np.abs(y.astype(np.float64) - ref).max(axis=0) / np.maximum(pk, 1e-300)).max())
                worst = max(worst, err) if dtype == np.float32 else worst
                if err > bar:
                    n_fail += 1
                    print(f"FAIL iteration {it} op {op}: family {family}, {ch} ch {np.dtype(dtype).name}, {n} frames, path {gpu.last_path()}: err/peak {err:.3e}", flush=True)
    gpu.close()
print(f"ops soak: {iters} chains, {n_ops} operations, worst f32 err / own peak {worst:.3e}: {'ok' if n_fail == 0 else str(n_fail) + ' FAILED'}", flush=True)
sys.exit(1 if n_fail else 0)
