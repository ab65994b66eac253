"""pipe_b200 -- B200-native Processor hot path behind pipelined/pipe's plugin boundary.

Layout:
  pipe_b200.abi     ctypes binding of the C-ABI (include/pipe_b200.h); raises if
                    libpipe_b200.so is missing -- there is no CPU fallback.
  pipe_b200.pipe    host-side mirror of the reference API for this path
                    (Line / Source / Processor / Sink / run, mock components).
  pipe_b200.gpu     GPU Processors: ProcessorAllocatorFunc factories that put a
                    fused CUDA chain behind pipe.Processor.
  pipe_b200.design  coefficient helpers (numpy).
"""
__version__ = "0.1.0"
