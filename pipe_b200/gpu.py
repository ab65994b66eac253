"""GPU Processors: ProcessorAllocatorFunc factories that put one fused CUDA chain
behind pipe.Processor (the reference's plugin boundary, line.go:30 / pipe.go:52-64).

    line = pipe.Line(source=src.source(),
                     processors=pipe.processors(gpu.chain([gain(0.8), fir(taps), ...])),
                     sink=snk.sink())

A contiguous run of GPU stages is ONE Processor to the pipe and ONE pb_chain to
the C-ABI, so the per-buffer walk of Processor.execute (pipe.go:425-451) makes a
single call across the boundary for the whole run.  There is no CPU fallback:
allocation raises when libpipe_b200.so or the device is missing, which aborts
binding exactly like an allocator error in the reference (line.go:72-74).
"""
from __future__ import annotations

import numpy as np

from . import abi
from .pipe import Processor, ProcessorAllocatorFunc, SignalProperties


def copy() -> dict:
    return {"kind": "copy"}


def gain(g: float) -> dict:
    return {"kind": "gain", "gain": float(g)}


def biquad(b, a) -> dict:
    return {"kind": "biquad", "b": tuple(b), "a": tuple(a)}


def fir(taps) -> dict:
    return {"kind": "fir", "taps": np.asarray(taps, dtype=np.float64)}


def resample(up: int, down: int, prototype) -> dict:
    return {"kind": "resample", "up": int(up), "down": int(down), "taps": np.asarray(prototype, dtype=np.float64)}


class ChainProcessor:
    """One fused pb_chain behind pipe.Processor.  Keeps the handle so tests and mutations can reach it (set_stage == a
    mutation, pipe.go:433).

    Lifecycle: the reference binds components once in pipe.New and a Pipe may be started again after Wait
    (TestReset, pipe_test.go:107-130).  So the handle is created by the allocator and lives as long as this object:
    StartFunc zeroes the carried state (pb_chain_reset), FlushFunc only drains the device -- it does NOT destroy the chain --
    and close() / the finalizer release it.

    dtype is the type of the buffers the pipe hands over (float64 in the reference: pipe.go:394,437); compute_dtype is the
    type the chain computes in.  With compute_dtype=float32 the marshalling copy converts on the way in and out, which is
    what reaches the tcgen05 kernel (K2) and the float32 streaming kernels from a float64 pipe.
    """

    def __init__(self, stages: list[dict], dtype=np.float64, device: int = 0, flags: int = 0, compute_dtype=None):
        self.stages, self.dtype, self.device, self.flags = stages, dtype, device, flags
        self.compute_dtype = compute_dtype if compute_dtype is not None else dtype
        self.chain: abi.Chain | None = None
        self.messages = 0
        self.samples = 0
        self.starts = 0

    def close(self) -> None:
        if self.chain is not None:
            self.chain.close()
            self.chain = None

    def processor(self) -> ProcessorAllocatorFunc:
        def alloc(buffer_size: int, props: SignalProperties) -> Processor:
            # allocate every buffer and table up front (line.go:27-30); a second binding replaces the first
            self.close()
            self.chain = abi.Chain(props.channels, self.stages, buffer_frames=buffer_size, dtype=self.compute_dtype,
                                   sample_rate=props.sample_rate, device=self.device, flags=self.flags)
            ch, sr = self.chain.out_properties()
            lib = abi.lib()
            got = abi._i64()
            np_dtype = self.chain.np_dtype

            def handle():
                if self.chain is None or not self.chain._h:
                    raise abi.PipeB200Error(abi.PB_ERR_STATE, "gpu.chain: the chain has been closed")
                return self.chain._h

            def start_func() -> None:  # StartFunc (pipe.go:84): a (re)started Pipe begins from zero state
                if self.starts:
                    abi.check(lib.pb_chain_reset(handle()))
                self.starts += 1

            def process_func(inp: np.ndarray, out: np.ndarray) -> int:
                x = inp if (inp.dtype == np_dtype and inp.flags.c_contiguous) else np.ascontiguousarray(inp, dtype=np_dtype)
                direct = out.dtype == np_dtype and out.flags.c_contiguous
                dst = out if direct else np.empty(out.shape, dtype=np_dtype)
                abi.check(lib.pb_chain_process(handle(), x.ctypes.data, len(x), dst.ctypes.data, len(dst), abi.C.byref(got)))
                n = got.value
                if not direct:
                    out[:n] = dst[:n]
                self.messages += 1
                self.samples += n
                return n

            def flush_func() -> None:  # FlushFunc (run.go:181-185): everything enqueued has completed; the chain stays bound
                if self.chain is not None and self.chain._h:
                    self.chain.sync()

            return Processor(process_func, start_func, flush_func, SignalProperties(sr, ch))
        return alloc

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def chain(stages: list[dict], dtype=np.float64, device: int = 0, flags: int = 0, compute_dtype=None) -> ProcessorAllocatorFunc:
    """ProcessorAllocatorFunc for a run of GPU stages.  compute_dtype=np.float32 selects the float32 kernels (K2 / K3) while the
    pipe keeps handing over float64 buffers, as the reference does."""
    return ChainProcessor(stages, dtype=dtype, device=device, flags=flags, compute_dtype=compute_dtype).processor()
