"""How fast does a read-only kernel stream HBM on this part?  Times the stand-alone meter Sink (pb_meter_device: one read per
sample, no writes) over a 320 MiB buffer and torch's own sum().  Run on a GPU box."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pipe_b200 import abi  # noqa: E402

for ch in (1024, 64):
    frames = 320 * 1024 * 1024 // 4 // ch
    x = torch.empty((frames, ch), dtype=torch.float32, device="cuda:0")
    abi.source_fill(x.data_ptr(), abi.PB_F32, 0, x.numel(), seed=7)
    m = torch.zeros((2, ch), dtype=torch.float64, device="cuda:0")
    st = torch.cuda.current_stream()
    for name, fn in (("pb_meter_device", lambda: abi.meter_device(x.data_ptr(), abi.PB_F32, frames, ch, m.data_ptr(), m.data_ptr() + 8 * ch,
                                                                  stream=st.cuda_stream)),
                     ("torch.sum", lambda: x.sum())):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(10):
            fn()
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name:16s} {ch:5d} ch: {ms * 1e3:7.1f} us per 320 MiB read = {x.numel() * 4 / ms / 1e6:6.0f} GB/s", flush=True)
