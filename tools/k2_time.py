"""Development timing of K2 on the bench workload (configs[2], 20 buffers of 4096 x 1024 f32 per launch, device-resident),
without the rest of bench.py.  Run on a GPU box:  python tools/k2_time.py [steps]
PB_TC_PROF=1 / 2 print the per-role counters / the event timeline of CTA 0 (pipe through tools/trace_fmt.py)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pipe_b200 import abi, design  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ch, bf, nb = 1024, 4096, int(os.environ.get("K2_NB", "20"))
chain = abi.Chain(ch, design.config_stages("chain4"), buffer_frames=bf, max_batch=nb)
x = torch.empty((bf * nb, ch), dtype=torch.float32, device="cuda:0")
y = torch.empty((bf * nb, ch), dtype=torch.float32, device="cuda:0")
abi.source_fill(x.data_ptr(), abi.PB_F32, 0, x.numel(), seed=1234)
st = torch.cuda.current_stream()
for _ in range(3):
    chain.process_batch_device(x.data_ptr(), [bf] * nb, y.data_ptr(), bf * nb, stream=st.cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(steps):
    chain.process_batch_device(x.data_ptr(), [bf] * nb, y.data_ptr(), bf * nb, stream=st.cuda_stream)
e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
units = bf * nb * ch
gbs = units * 7.675 / ms / 1e6
print(f"K2 batch of {nb} x {bf} x {ch}: {ms:.4f} ms per launch = {units / ms / 1e6:.1f} Gsamples/s, {gbs:.0f} GB/s algorithmic "
      f"({gbs / 6543.4:.3f} of 6543 GB/s), path {chain.last_path()}", flush=True)
chain.close()
