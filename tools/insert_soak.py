"""Randomised soak of InsertProcessor on a fused run (pb_chain_insert_stage, reference pipe.go:297) against the oracle's stage list:
chains start as [gain, biquad] or the 4-Processor headline chain, random gains / biquads / FIRs are spliced in at random positions
between two buffers, every stage that was there must keep its carried state, the new ones start from zero.
Run on a GPU box:    python tools/insert_soak.py [iterations] [seed]"""
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
worst, n_edits, n_fail, paths = 0.0, 0, 0, {}
for it in range(iters):
    headline = rng.random() < 0.5
    ch = int(rng.choice([128, 256])) if headline else int(rng.choice([3, 48, 64]))
    dtype = np.float32 if (headline or rng.random() < 0.6) else np.float64
    stages = [dict(s) for s in design.config_stages("chain4" if headline else "gain_biquad")]
    bf = int(rng.choice([1000, 1600]))
    gpu, cpu = abi.Chain(ch, stages, buffer_frames=bf, dtype=dtype), orc.StageList(ch, stages)
    bar = 2e-6 if dtype == np.float32 else 1e-9     # (an edit can cut the run into two fused segments: one more float32 hand-over)
    n_stages, run_peak = len(stages), np.zeros(ch)
    for step in range(int(rng.integers(5, 9))):
        if step > 0 and rng.random() < 0.4 and n_stages < 8:
            r = rng.random()
            if r < 0.4:
                new = {"kind": "gain", "gain": float(rng.uniform(0.5, 1.5))}
            elif r < 0.75:
                b, a = design.biquad(str(rng.choice(["highpass", "peaking"])), float(rng.choice([300.0, 2000.0])), 48000.0, q=float(rng.uniform(0.7, 1.5)))
                new = {"kind": "biquad", "b": b, "a": a}
            else:
                new = {"kind": "fir", "taps": design.lowpass_fir(int(rng.choice([17, 65])), float(rng.uniform(0.25, 0.45)))}
            # (in front of a resampler only: behind it the sample rate -- and the oracle's filter design -- would be another)
            rs = [i for i, s in enumerate(stages) if s["kind"] == "resample"]
            pos = int(rng.integers(0, (rs[0] if rs else n_stages) + 1))
            try:
                gpu.insert_stage(pos, new)
            except abi.PipeB200Error as e:
                print(f"  (iteration {it}: insert of {new['kind']} at {pos} refused: {e})", flush=True)
                continue
            cpu.insert(pos, new)
            stages.insert(pos, new)
            n_stages += 1
            n_edits += 1
        n = int(rng.integers(200, bf + 1)) if rng.random() < 0.3 else bf
        x = orc.source_fill(int(rng.integers(0, 1 << 30)), n * ch).reshape(n, ch)
        ref = cpu.process(x)
        y = gpu.process(x.astype(dtype))
        assert len(y) == len(ref), (len(y), len(ref))
        run_peak = np.maximum(run_peak, np.abs(ref).max(axis=0))
        pk = np.abs(ref).max(axis=0) if len(ref) >= 256 else run_peak
        err = float((np.abs(y.astype(np.float64) - ref).max(axis=0) / np.maximum(pk, 1e-300)).max())
        p0 = gpu.last_path()[0]
        paths[p0] = paths.get(p0, 0) + 1
        if dtype == np.float32:
            worst = max(worst, err)
        if err > bar:
            n_fail += 1
            print(f"FAIL iteration {it} step {step}: {ch} ch {np.dtype(dtype).name}, stages {[s['kind'] for s in stages]}, path {gpu.last_path()}: err/peak {err:.3e}", flush=True)
    gpu.close()
print(f"insert soak: {iters} chains, {n_edits} edits (kernel paths of the last segment {paths}), worst f32 err / own peak {worst:.3e}: "
      f"{'ok' if n_fail == 0 else str(n_fail) + ' FAILED'}", flush=True)
sys.exit(1 if n_fail else 0)
