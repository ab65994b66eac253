"""Device-resident throughput of configs[2] when the caller hands over ONE 4096-frame buffer per call (what the reference's
ProcessFunc does, pipe.go:438) instead of a batch.  4096 is not a multiple of K2's 160-frame tile: the tiles of a call start at
its first frame (per-phase tables) and the last tile is partial -- one K2 launch plus its verify launch per call (PB_TC_TAIL_K1=1: whole tiles on K2, the rest on K1, as before).  Run on a GPU box:  python tools/per_buffer.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pipe_b200 import abi, design  # noqa: E402

ch, bf, nbuf = 1024, 4096, 40
for flags, name in ((0, "K2, one launch per call"), (abi.CHAIN_NO_TENSOR, "K1 only (PB_CHAIN_NO_TENSOR)")):
    chain = abi.Chain(ch, design.config_stages("chain4"), buffer_frames=bf, flags=flags)
    x = torch.empty((bf * nbuf, ch), dtype=torch.float32, device="cuda:0")
    y = torch.empty((bf, ch), dtype=torch.float32, device="cuda:0")
    abi.source_fill(x.data_ptr(), abi.PB_F32, 0, x.numel(), seed=1234)
    st = torch.cuda.current_stream()
    for b in range(5):
        chain.process_batch_device(x.data_ptr() + b * bf * ch * 4, [bf], y.data_ptr(), bf, stream=st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for b in range(5, nbuf):
        chain.process_batch_device(x.data_ptr() + b * bf * ch * 4, [bf], y.data_ptr(), bf, stream=st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (nbuf - 5)
    print(f"{name:32s} {ms * 1e3:7.1f} us per 4096 x 1024 buffer = {bf * ch / ms / 1e6:7.1f} Gsamples/s, path {chain.last_path()}", flush=True)
    chain.close()
