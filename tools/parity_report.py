"""Print the parity margin (worst error / channel peak) of the CUDA path against the oracle
for the BASELINE.json configs.  Run on a GPU box:  python tools/parity_report.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle as orc  # noqa: E402
from pipe_b200 import abi, design  # noqa: E402


def margin(cfg, channels, frames, buffers, flags=0):
    st = design.config_stages(cfg)
    gpu, cpu = abi.Chain(channels, st, buffer_frames=frames, flags=flags), orc.Chain(channels, st)
    worst = 0.0
    for b in range(buffers):
        x = orc.source_fill(b * frames * channels, frames * channels).reshape(frames, channels)
        ref = cpu.process(x, threads=os.cpu_count())
        y = gpu.process(x.astype(np.float32))
        worst = max(worst, float(np.max(np.abs(y - ref).max(axis=0) / np.abs(ref).max(axis=0))))
    path = gpu.last_path()[0]
    return worst, path


if __name__ == "__main__":
    for cfg, ch, fr, nb in (("gain_biquad", 64, 4096, 4), ("chain4", 64, 4096, 3), ("chain4", 1024, 4096, 2)):
        for flags in (0, abi.CHAIN_NO_TENSOR, abi.CHAIN_NO_STREAM):  # default paths, then K1 instead of K2, K1 instead of K3
            w, path = margin(cfg, ch, fr, nb, flags)
            print(f"{cfg:12s} {ch:5d} ch x {fr} x {nb}  flags={flags} path={path}  worst err/peak = {w:.3e}  (bar 1e-6)")
