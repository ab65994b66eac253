"""Development check of the tcgen05 path (K2) against the oracle.  Run on a GPU box."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle as orc  # noqa: E402
from pipe_b200 import abi, design  # noqa: E402


def check(channels, frames, buffers, amp=1.0, flags=0):
    st = design.config_stages("chain4")
    gpu, cpu = abi.Chain(channels, st, buffer_frames=frames, flags=flags), orc.Chain(channels, st)
    worst = 0.0
    for b in range(buffers):
        x = amp * orc.source_fill(b * frames * channels, frames * channels).reshape(frames, channels)
        ref = cpu.process(x, threads=os.cpu_count())
        y = gpu.process(x.astype(np.float32))
        assert y.shape == ref.shape, (y.shape, ref.shape)
        err = np.abs(y - ref).max(axis=0) / np.abs(ref).max(axis=0)
        worst = max(worst, float(err.max()))
        if err.max() > 1e-4:
            bad = np.argwhere(np.abs(y - ref) > 1e-4 * np.abs(ref).max())
            print(f"  buffer {b}: worst {err.max():.3e}; {len(bad)} bad entries; first {bad[:6].tolist()}")
            print("   y  ", y[bad[0][0], :4], "\n   ref", ref[bad[0][0], :4])
    print(f"chain4 {channels:5d} ch x {frames} x {buffers} amp={amp} flags={flags} path={gpu.last_path()} worst err/peak = {worst:.3e}")


if __name__ == "__main__":
    check(128, 160, 1)
    check(128, 160, 3)
    check(128, 1600, 3)
    check(1024, 1600, 3)
    check(1024, 4000, 2, amp=0.5)
    check(256, 1600, 2, flags=abi.CHAIN_NO_TENSOR)
