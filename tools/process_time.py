"""Development timing of pb_chain_process on ONE 4096 x 1024 host buffer per call (pinned staging, as the cgo shim has it):
the library's share of the drop-in call.  PB_PROCESS_PIECES=1 switches the pieces off.  Run on a GPU box."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pipe_b200 import abi, design  # noqa: E402

ch, bf = 1024, 4096
for cfg, dt in (("chain4", np.float32), ("chain4", np.float64), ("gain_biquad", np.float32)):
    chain = abi.Chain(ch, design.config_stages(cfg), buffer_frames=bf, dtype=dt)
    item = np.dtype(dt).itemsize
    pin_in, pin_out = abi.PinnedBuffer(bf * ch * item), abi.PinnedBuffer(bf * ch * item)
    pin_in.array((bf, ch), dt)[:] = np.random.default_rng(1).uniform(-0.8, 0.8, (bf, ch))
    got = abi._i64()
    for _ in range(5):
        abi.check(abi.lib().pb_chain_process(chain._h, pin_in.ptr, bf, pin_out.ptr, bf, abi.C.byref(got)))
    t0 = time.perf_counter()
    n = 30
    for _ in range(n):
        abi.check(abi.lib().pb_chain_process(chain._h, pin_in.ptr, bf, pin_out.ptr, bf, abi.C.byref(got)))
    ms = 1e3 * (time.perf_counter() - t0) / n
    print(f"{cfg:12s} {np.dtype(dt).name:8s} pieces={os.environ.get('PB_PROCESS_PIECES', '4')}: {ms:.3f} ms per pb_chain_process "
          f"({bf * ch / ms / 1e3:.0f} Msamples/s), path {chain.last_path()}")
    chain.close()
    pin_in.free()
    pin_out.free()
