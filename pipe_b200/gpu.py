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
    """Keeps the handle so tests and mutations can reach it (set_stage == a mutation, pipe.go:433)."""

    def __init__(self, stages: list[dict], dtype=np.float64, device: int = 0, flags: int = 0):
        self.stages, self.dtype, self.device, self.flags = stages, dtype, device, flags
        self.chain: abi.Chain | None = None
        self.messages = 0
        self.samples = 0

    def processor(self) -> ProcessorAllocatorFunc:
        def alloc(buffer_size: int, props: SignalProperties) -> Processor:
            # allocate every buffer and table up front (line.go:27-30)
            self.chain = abi.Chain(props.channels, self.stages, buffer_frames=buffer_size, dtype=self.dtype,
                                   sample_rate=props.sample_rate, device=self.device, flags=self.flags)
            ch, sr = self.chain.out_properties()
            lib, h = abi.lib(), self.chain._h
            got = abi._i64()
            np_dtype = self.chain.np_dtype

            def process_func(inp: np.ndarray, out: np.ndarray) -> int:
                x = inp if (inp.dtype == np_dtype and inp.flags.c_contiguous) else np.ascontiguousarray(inp, dtype=np_dtype)
                direct = out.dtype == np_dtype and out.flags.c_contiguous
                dst = out if direct else np.empty(out.shape, dtype=np_dtype)
                abi.check(lib.pb_chain_process(h, x.ctypes.data, len(x), dst.ctypes.data, len(dst), abi.C.byref(got)))
                n = got.value
                if not direct:
                    out[:n] = dst[:n]
                self.messages += 1
                self.samples += n
                return n

            def flush_func() -> None:  # FlushFunc is the guaranteed teardown point (run.go:181-185)
                if self.chain is not None:
                    self.chain.close()

            return Processor(process_func, None, flush_func, SignalProperties(sr, ch))
        return alloc


def chain(stages: list[dict], dtype=np.float64, device: int = 0, flags: int = 0) -> ProcessorAllocatorFunc:
    return ChainProcessor(stages, dtype=dtype, device=device, flags=flags).processor()
