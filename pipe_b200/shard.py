"""Multi-GPU partitioning of the Processor path.

Lines share nothing in the reference -- each has its own source, fittings, pools and
sink (pipe.go:134-141,172-184; line.go:62-104) -- so the partition is one Line per GPU
with NO data-path collective (BASELINE.json configs[3]).  The only exchange step is the
build-defined fan-in mixer of configs[4] (the reference's merger.go merges error
channels, not signals: SURVEY.md D2), offered two ways:

  * fan_in_reduce():  torch.distributed reduce(SUM) -- NCCL over NVLink on GPUs
    (gloo on CPU for the host-logic tests);
  * PeerFanIn:        every rank exports its output buffer with pb_ipc_export, the
    destination rank opens the peers' buffers and ONE pb_mix_sum_device kernel pulls
    them over NVLink while summing (transfer and sum fused, no staging copy).
"""
from __future__ import annotations

import numpy as np


def assign_lines(n_lines: int, world: int) -> list[list[int]]:
    """Line i runs on rank i % world (independent Lines, round-robin)."""
    if n_lines < 0 or world < 1:
        raise ValueError("assign_lines: need n_lines >= 0 and world >= 1")
    return [[i for i in range(n_lines) if i % world == r] for r in range(world)]


def fan_in_reduce(tensor, dst: int = 0, group=None):
    """In-place sum of every rank's buffer onto rank `dst` (configs[4])."""
    import torch.distributed as dist
    dist.reduce(tensor, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return tensor


class PeerFanIn:
    """Fan-in sum over peer memory: rank `dst` reads the other ranks' device buffers directly."""

    def __init__(self, local_ptr: int, n_values: int, dtype: int, device: int, dst: int = 0, group=None):
        import torch.distributed as dist

        from . import abi
        self.abi, self.dist, self.group = abi, dist, group
        self.rank, self.world, self.dst = dist.get_rank(group), dist.get_world_size(group), dst
        self.n_values, self.dtype, self.device, self.local_ptr = n_values, dtype, device, local_ptr
        handle = (abi.C.c_uint8 * 64)()
        abi.check(abi.lib().pb_ipc_export(device, local_ptr, handle))
        # the handle names the ALLOCATION the buffer lies in (a caching allocator hands out pieces of larger blocks): the
        # buffer's offset inside it travels with the handle
        off = abi._i64()
        abi.check(abi.lib().pb_ipc_offset(device, local_ptr, abi.C.byref(off)))
        handles = [None] * self.world
        dist.all_gather_object(handles, (bytes(handle), off.value), group=group)
        self.peer_ptrs: list[int] = []
        self._bases: list[int] = []
        if self.rank == dst:
            for r, (h, o) in enumerate(handles):
                if r == dst:
                    self.peer_ptrs.append(local_ptr)
                    self._bases.append(0)
                    continue
                buf = (abi.C.c_uint8 * 64).from_buffer_copy(h)
                p = abi._vp()
                abi.check(abi.lib().pb_ipc_open(device, buf, abi.C.byref(p)))
                self._bases.append(p.value)
                self.peer_ptrs.append(p.value + o)

    def sum_into(self, out_ptr: int, stream: int = 0):
        """Call on every rank after its producer kernel finished; only `dst` launches the sum."""
        self.dist.barrier(group=self.group)      # peers' buffers are complete
        if self.rank == self.dst:
            self.abi.mix_sum(self.peer_ptrs, self.dtype, self.n_values, out_ptr, device=self.device, stream=stream)
        # the caller must not overwrite its buffer before dst's kernel has read it:
        # synchronise dst's stream, then barrier (see bench.py fan_in_run)

    def close(self):
        if self.rank == self.dst:
            for r, p in enumerate(self._bases):
                if r != self.dst:
                    self.abi.lib().pb_ipc_close(self.device, p)
        self.peer_ptrs = []
