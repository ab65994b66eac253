"""Throughput of the generic fused kernel (K1) on the HBM-bound Processor runs, device-resident, against the
measured HBM peak.  Run on a GPU box:  python tools/hbm_chains.py [--json out.json]

Cases (a step is one batch launch; every batch is larger than the 126 MB L2):
  configs[1]  gain + biquad, 64 ch x 4096-frame buffers, f32      (8 B per sample)
  the same run at 1024 ch, f32 and f64                           (8 / 16 B per sample)
  gain only, 1024 ch f32                                         (8 B per sample)
  configs[0] at scale: mock.Processor copy, 2 ch f64             (16 B per sample)
  copy, 1024 ch f64                                              (16 B per sample)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from pipe_b200 import abi, design  # noqa: E402


def peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6650.0


def run(name, stages, channels, frames, nb, dtype, steps=10, warmup=3):
    dev = torch.device("cuda", 0)
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    total = frames * nb
    chain = abi.Chain(channels, stages, buffer_frames=frames, max_batch=nb, dtype=dtype)
    x = torch.empty((total, channels), dtype=tdt, device=dev)
    y = torch.empty((total, channels), dtype=tdt, device=dev)
    abi.source_fill(x.data_ptr(), abi.PB_F32 if dtype == np.float32 else abi.PB_F64, 0, total * channels, seed=1234)
    torch.cuda.synchronize()
    st = torch.cuda.current_stream()
    sizes = [frames] * nb
    for _ in range(warmup):
        chain.process_batch_device(x.data_ptr(), sizes, y.data_ptr(), total, stream=st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        chain.process_batch_device(x.data_ptr(), sizes, y.data_ptr(), total, stream=st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    el = 4 if dtype == np.float32 else 8
    nbytes = 2.0 * total * channels * el
    gbs = nbytes / (ms * 1e-3) / 1e9
    row = {"case": name, "channels": channels, "buffer_frames": frames, "batch_buffers": nb, "dtype": np.dtype(dtype).name,
           "ms_per_launch": ms, "Msamples_per_s": total * channels / (ms * 1e-3) / 1e6, "GBps": gbs, "hbm_frac": gbs / peak(),
           "kernel_path": chain.last_path()[0], "MB_per_launch": nbytes / 1e6}
    print(json.dumps(row), flush=True)
    chain.close()
    del x, y
    return row


if __name__ == "__main__":
    gb = design.config_stages("gain_biquad")
    cases = [
        ("configs[1] gain+biquad 64ch f32", gb, 64, 4096, 320, np.float32),
        ("gain+biquad 1024ch f32", gb, 1024, 4096, 20, np.float32),
        ("gain+biquad 1024ch f64", gb, 1024, 4096, 10, np.float64),
        ("gain 1024ch f32", [{"kind": "gain", "gain": 0.5}], 1024, 4096, 20, np.float32),
        # configs[0]'s 512-frame buffers are 8 KiB each: the batch is given as 65536-frame buffers so that the host-side
        # per-buffer bookkeeping does not dominate a 40 us launch
        ("configs[0] copy 2ch f64", design.config_stages("passthrough"), 2, 65536, 128, np.float64),
        ("copy 1024ch f64", design.config_stages("passthrough"), 1024, 4096, 10, np.float64),
    ]
    if "--case" in sys.argv:
        cases = [cases[int(sys.argv[sys.argv.index("--case") + 1])]]
    rows = [run(*c) for c in cases]
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump({"hbm_peak_gbs": peak(), "rows": rows}, f, indent=1)
