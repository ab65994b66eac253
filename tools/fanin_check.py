"""configs[4] on hardware: N Lines (one per GPU, 256 ch each) -> fan-in sum on rank 0, two ways
(NCCL reduce, and one peer-memory mixer kernel), both checked against the oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 tools/fanin_check.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import _oracle as orc  # noqa: E402
from pipe_b200 import abi, design, shard  # noqa: E402

CH, BF, NB = 256, 4000, 4   # 4000 = 25 tiles of 160 frames: the tcgen05 kernel


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    stages = design.config_stages("chain4")
    frames = BF * NB
    chain = abi.Chain(CH, stages, buffer_frames=BF, max_batch=NB, device=local)
    x = torch.empty((frames, CH), dtype=torch.float32, device=dev)
    y = torch.zeros((frames, CH), dtype=torch.float32, device=dev)
    abi.source_fill(x.data_ptr(), abi.PB_F32, 0, frames * CH, seed=1234, line=rank, device=local)
    torch.cuda.synchronize()
    sptr = torch.cuda.current_stream().cuda_stream
    counts = chain.process_batch_device(x.data_ptr(), [BF] * NB, y.data_ptr(), frames, stream=sptr)
    chain.sync(sptr)
    n_out = sum(counts)
    path = chain.last_path()[0]
    mine = y[:n_out].clone()

    # (a) NCCL reduce over NVLink
    red = mine.clone()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    shard.fan_in_reduce(red, dst=0)
    torch.cuda.synchronize()
    t_nccl = time.perf_counter() - t0

    # (b) one mixer kernel pulling the peers' buffers over NVLink while summing
    out = torch.zeros_like(mine)
    pf = shard.PeerFanIn(mine.data_ptr(), mine.numel(), abi.PB_F32, local, dst=0)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    pf.sum_into(out.data_ptr(), stream=sptr)
    torch.cuda.synchronize()
    dist.barrier()
    t_peer = time.perf_counter() - t0
    pf.close()

    if rank == 0:
        ref = None
        for line in range(world):
            cpu = orc.Chain(CH, stages)
            xs = orc.source_fill(0, frames * CH, line=line).reshape(frames, CH)
            r = np.concatenate([cpu.process(xs[i * BF:(i + 1) * BF], threads=os.cpu_count() or 1) for i in range(NB)])
            ref = r if ref is None else ref + r
        assert ref.shape[0] == n_out, (ref.shape, n_out)
        pk = np.abs(ref).max(axis=0)
        e_nccl = float((np.abs(red.cpu().numpy() - ref).max(axis=0) / pk).max())
        e_peer = float((np.abs(out.cpu().numpy() - ref).max(axis=0) / pk).max())
        mb = mine.numel() * 4 / 1e6
        print(f"fan-in of {world} Lines x {CH} ch x {n_out} frames (kernel path {path}): NCCL reduce err/peak {e_nccl:.2e} in "
              f"{1e3 * t_nccl:.3f} ms, peer-memory mixer err/peak {e_peer:.2e} in {1e3 * t_peer:.3f} ms ({mb:.1f} MB per Line)")
        assert e_nccl < 2e-6 and e_peer < 2e-6
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
