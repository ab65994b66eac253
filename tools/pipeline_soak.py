"""Randomised soak of the pipelined host path (pb_chain_submit / pb_chain_collect, two batches in flight: the analogue of the cap-1
async fitting, reference internal/fitting/fitting.go:56-60) against the CPU oracle: random batch sizes, ragged last buffers, pinned
AND pageable host memory, the three kernel families, interleaved with synchronous pb_chain_process calls on the same handle.
Run on a GPU box:    python tools/pipeline_soak.py [iterations] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle as orc  # noqa: E402
from pipe_b200 import abi, design  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
worst, n_batches, n_fail = 0.0, 0, 0
for it in range(iters):
    family = int(rng.integers(0, 3))
    if family == 0:
        ch, flags, stages = int(rng.choice([128, 256])), 0, design.config_stages("chain4")
    elif family == 1:
        ch, flags, stages = int(rng.choice([2, 64, 100])), 0, design.config_stages("gain_biquad")
    else:
        ch, flags, stages = int(rng.choice([16, 48])), abi.CHAIN_NO_TENSOR, design.config_stages("chain4")
    bf, nb = int(rng.choice([512, 1600, 4096])), int(rng.choice([1, 3, 5]))
    gpu, cpu = abi.Chain(ch, stages, buffer_frames=bf, max_batch=nb, flags=flags), orc.Chain(ch, stages)
    cap = bf * nb
    pinned = rng.random() < 0.6
    if pinned:
        bufs_in = [abi.PinnedBuffer(cap * ch * 4) for _ in range(2)]
        bufs_out = [abi.PinnedBuffer(cap * ch * 4) for _ in range(2)]
        arr_in = [b.array((cap, ch), np.float32) for b in bufs_in]
        arr_out = [b.array((cap, ch), np.float32) for b in bufs_out]
    else:
        arr_in = [np.zeros((cap, ch), np.float32) for _ in range(2)]
        arr_out = [np.zeros((cap, ch), np.float32) for _ in range(2)]
    pending = []   # (slot, sizes, ref)

    def check_out(y, ref, what):
        global worst, n_fail
        if len(ref) == 0:
            return
        pk = np.maximum(np.abs(ref).max(axis=0), 1e-2)
        err = float((np.abs(y.astype(np.float64) - ref).max(axis=0) / pk).max())
        worst = max(worst, err)
        if err > 1e-6:
            n_fail += 1
            print(f"FAIL iteration {it} {what}: family {family}, {ch} ch, bf {bf} x {nb}, pinned {pinned}: err/peak {err:.3e}", flush=True)

    def collect_one():
        slot, sizes, ref = pending.pop(0)
        counts = gpu.collect(len(sizes))
        assert sum(counts) == len(ref), (counts, len(ref))
        check_out(arr_out[slot][:len(ref)].copy(), ref, f"batch of {len(sizes)}")

    steps = int(rng.integers(4, 9))
    for step in range(steps):
        if rng.random() < 0.2 and not pending:   # a synchronous call in between (nothing may be in flight on the handle)
            n = int(rng.integers(1, bf + 1))
            x = orc.source_fill(int(rng.integers(0, 1 << 30)), n * ch).reshape(n, ch)
            check_out(gpu.process(x.astype(np.float32)), cpu.process(x), "synchronous call")
            continue
        k = int(rng.integers(1, nb + 1))
        sizes = [bf] * (k - 1) + [int(rng.integers(1, bf + 1)) if rng.random() < 0.4 else bf]
        total = sum(sizes)
        x = orc.source_fill(int(rng.integers(0, 1 << 30)), total * ch).reshape(total, ch)
        ref = cpu.process(x)
        if len(pending) == 2:
            collect_one()
        slot = 0 if not pending else 1 - pending[-1][0]
        arr_in[slot][:total] = x
        gpu.submit(arr_in[slot].ctypes.data, sizes, arr_out[slot].ctypes.data, cap)
        pending.append((slot, sizes, ref))
        n_batches += 1
    while pending:
        collect_one()
    gpu.close()
print(f"pipeline soak: {iters} chains, {n_batches} batches, worst err / peak {worst:.3e}: {'ok' if n_fail == 0 else str(n_fail) + ' FAILED'}", flush=True)
sys.exit(1 if n_fail else 0)
