"""Experiment: the kernels read the host buffer in place (pinned memory is device-accessible under UVA) and store the result
straight into pinned host memory, instead of H2D copy -> kernels -> D2H copy.  Run on a GPU box."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from pipe_b200 import abi, design  # noqa: E402

ch, bf = 1024, 4096
for cfg, dt in (("chain4", np.float32), ("gain_biquad", np.float32), ("chain4", np.float64)):
    item = np.dtype(dt).itemsize
    pin_in, pin_out = abi.PinnedBuffer(bf * ch * item), abi.PinnedBuffer(bf * ch * item)
    x = np.random.default_rng(1).uniform(-0.8, 0.8, (bf, ch)).astype(dt)
    pin_in.array((bf, ch), dt)[:] = x
    ref_chain = abi.Chain(ch, design.config_stages(cfg), buffer_frames=bf, dtype=dt)
    for mode in ("in+out", "in only", "out only"):
        chain = abi.Chain(ch, design.config_stages(cfg), buffer_frames=bf, dtype=dt)
        d_in, d_out = abi.DeviceBuffer(bf * ch * item), abi.DeviceBuffer(bf * ch * item)
        d_in.upload(x)
        src = pin_in.ptr if mode != "out only" else d_in.ptr
        dst = pin_out.ptr if mode != "in only" else d_out.ptr
        try:
            for _ in range(3):
                n = chain.process_batch_device(src, [bf], dst, bf)
                chain.sync()
            t0 = time.perf_counter()
            reps = 20
            for _ in range(reps):
                n = chain.process_batch_device(src, [bf], dst, bf)
                chain.sync()
            ms = 1e3 * (time.perf_counter() - t0) / reps
            print(f"{cfg:12s} {np.dtype(dt).name:8s} zero-copy {mode:8s}: {ms:.3f} ms per call, path {chain.last_path()}, out frames {n}")
        except Exception as e:  # noqa: BLE001
            print(f"{cfg} {np.dtype(dt).name} zero-copy {mode}: FAILED {e}")
        chain.close()
        d_in.free()
        d_out.free()
    # check the in+out result against the ordinary host path on a fresh chain
    chain = abi.Chain(ch, design.config_stages(cfg), buffer_frames=bf, dtype=dt)
    n = chain.process_batch_device(pin_in.ptr, [bf], pin_out.ptr, bf)
    chain.sync()
    y0 = pin_out.array((bf, ch), dt)[:n[0]].copy()
    y1 = ref_chain.process(x)
    print("   max |zero-copy - copy path| / peak:", float(np.abs(y0 - y1).max() / np.abs(y1).max()))
    chain.close()
    ref_chain.close()
    pin_in.free()
    pin_out.free()
