"""compute-sanitizer target for the streaming kernels (K3): gain + biquad with the fused meter, several tiles per
channel group, a ragged last tile and an odd channel count; then a gain-only run and a copy.  Run on a GPU box:
  compute-sanitizer --tool memcheck  python tools/san_check_k3.py
  compute-sanitizer --tool racecheck python tools/san_check_k3.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle as orc  # noqa: E402
from pipe_b200 import abi, design  # noqa: E402

st = design.config_stages("gain_biquad")
# (64 ch x 70000 frames in one buffer: 274 tiles in 2 channel groups -> the two-sweep path: aggregate sweep, scan, apply sweep)
for ch, bf, dtype in ((64, 3000, np.float32), (33, 1500, np.float32), (40, 700, np.float64), (64, 70000, np.float32)):
    gpu, cpu = abi.Chain(ch, st, buffer_frames=bf, dtype=dtype, flags=abi.CHAIN_METER), orc.Chain(ch, st)
    for b in range(2):
        x = orc.source_fill(b * bf * ch, bf * ch).reshape(bf, ch)
        ref = cpu.process(x)
        y = gpu.process(x.astype(dtype))
        print(ch, bf, np.dtype(dtype).name, gpu.last_path(), float((np.abs(y - ref).max(axis=0) / np.abs(ref).max(axis=0)).max()))
x = orc.source_fill(0, 1001 * 7).reshape(1001, 7).astype(np.float32)
g = abi.Chain(7, [{"kind": "gain", "gain": 0.5}], buffer_frames=1001)
print("gain", g.last_path(), bool(np.array_equal(g.process(x), x * np.float32(0.5))))
g = abi.Chain(7, [{"kind": "copy"}], buffer_frames=1001)
print("copy", bool(np.array_equal(g.process(x), x)))
