"""Randomised parity soak of the tcgen05 chain kernel (K2) against the CPU oracle: random biquads (RBJ high-pass, low-pass, peaking), FIR lengths, channel counts
(multiples of 128), call lengths (whole and partial last tiles, calls shorter than a tile fall to K1), per-channel levels from
0 to -60 dBFS with jumps between calls, batches and single buffers, the fused meter.  Run on a GPU box:
    python tools/k2_soak.py [iterations] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle as orc  # noqa: E402
from pipe_b200 import abi, design  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
worst, n_k2, n_calls, n_fail = 0.0, 0, 0, 0
for it in range(iters):
    ch = int(rng.choice([128, 256, 384, 512]))
    kind = str(rng.choice(["highpass", "lowpass", "peaking"]))
    f0 = float(rng.choice([40.0, 200.0, 1000.0, 6000.0, 12000.0])) if kind != "lowpass" else float(rng.choice([6000.0, 12000.0, 18000.0]))
    b, a = design.biquad(kind, f0, 48000.0, q=float(rng.uniform(0.5, 2.0)), gain_db=float(rng.uniform(-6, 6)))
    taps = int(rng.choice([33, 129, 257]))
    stages = [{"kind": "gain", "gain": float(rng.uniform(0.3, 2.5))},
              {"kind": "fir", "taps": design.lowpass_fir(taps, float(rng.uniform(0.3, 0.45)))},
              {"kind": "biquad", "b": b, "a": a},
              {"kind": "resample", "up": 147, "down": 160, "taps": design.resampler_prototype(147, 160, 16)}]
    bf = int(rng.choice([1600, 4096, 4000, 1000]))
    nb = int(rng.choice([1, 1, 3, 5]))
    meter = bool(rng.integers(0, 2))
    gpu = abi.Chain(ch, stages, buffer_frames=bf, max_batch=nb, dtype=np.float32,
                    flags=(abi.CHAIN_METER if meter else 0) | (abi.CHAIN_NO_TENSOR if os.environ.get("SOAK_K1") else 0))
    cpu = orc.Chain(ch, stages)
    levels = 10.0 ** (-rng.integers(0, 4, size=ch) * 1.0)           # 0, -20, -40, -60 dBFS per channel
    prev_levels = levels
    jumps = []
    for call in range(int(rng.integers(2, 5))):
        prev_levels = levels
        jumps.append(False)
        if rng.random() < 0.3:
            jumps[-1] = True
            levels = 10.0 ** (-rng.integers(0, 4, size=ch) * 1.0)   # level jump: the speculated scales are wrong, the call is redone
        sizes = [bf] * (nb - 1) + [int(rng.integers(1, bf + 1)) if rng.random() < 0.5 else bf]
        total = sum(sizes)
        x = orc.source_fill(int(rng.integers(0, 1 << 30)), total * ch).reshape(total, ch) * levels
        d_in, d_out = abi.DeviceBuffer(total * ch * 4), abi.DeviceBuffer(total * ch * 4)
        d_in.upload(x.astype(np.float32))
        counts = gpu.process_batch_device(d_in.ptr, sizes, d_out.ptr, total)
        gpu.sync()
        ref = cpu.process(x, threads=os.cpu_count() or 1)
        assert sum(counts) == len(ref), (counts, len(ref))
        n_calls += 1
        if len(ref) == 0:      # (a call too short to trigger an output)
            continue
        y = d_out.download((sum(counts), ch), np.float32)
        # (channels whose output has not arrived yet -- a first call shorter than the filters' delay -- are held to 1 % of the input peak)
        pk = np.maximum(np.abs(ref).max(axis=0), 1e-2 * np.abs(x).max(axis=0))
        err = float((np.abs(y.astype(np.float64) - ref).max(axis=0) / np.maximum(pk, 1e-300)).max()) if len(ref) else 0.0
        path = gpu.last_path()[0]
        n_k2 += path == 2
        worst = max(worst, err)
        if err > 1e-6:
            e = np.abs(y.astype(np.float64) - ref)
            cw = int(np.argmax(e.max(axis=0) / np.maximum(pk, 1e-300)))
            fw = int(np.argmax(e[:, cw]))
            print(f"FAIL iteration {it} call {call}: {ch} ch, {kind} {f0} Hz, {taps} taps, bf {bf}, sizes {sizes}, meter {meter}, path {path}: err/peak {err:.3e}; "
                  f"worst channel {cw} (level now {levels[cw]:g}, before {prev_levels[cw]:g}), peak {pk[cw]:.3e}, worst output frame {fw} of {len(ref)}: "
                  f"y {y[fw, cw]:.6e} ref {ref[fw, cw]:.6e}; channels over 1e-6: {int((e.max(axis=0) > 1e-6 * pk).sum())}; level jumps per call so far {jumps}; "
                  f"err/peak over output frames 64.. : {float((e[64:].max(axis=0) / np.maximum(pk, 1e-300)).max()):.2e}", flush=True)
            n_fail += 1
    gpu.close()
print(f"k2 soak: {iters} chains, {n_calls} calls ({n_k2} on the tcgen05 kernel), worst err / own peak {worst:.3e}: {'ok' if n_fail == 0 else str(n_fail) + ' FAILED'}", flush=True)
sys.exit(1 if n_fail else 0)
