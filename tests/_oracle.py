"""ctypes front-end of the CPU oracle (oracle/pipe_oracle.{h,c}).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Nothing under pipe_b200/
imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")

STAGE_COPY, STAGE_GAIN, STAGE_BIQUAD, STAGE_FIR, STAGE_RESAMPLE = range(5)
RUN_OK, RUN_ERR_BIND, RUN_ERR_START, RUN_ERR_EXEC, RUN_ERR_FLUSH = 0, 1, 2, 4, 8
MAX_PROCS = 8


class _Stage(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("n_taps", C.c_int32), ("up", C.c_int32), ("down", C.c_int32),
        ("_pad", C.c_int32), ("gain", C.c_double), ("b", C.c_double * 3), ("a", C.c_double * 2),
        ("taps", C.POINTER(C.c_double)),
    ]


class MockComponent(C.Structure):
    _fields_ = [
        ("error_on_call", C.c_int32), ("error_on_make", C.c_int32),
        ("error_on_start", C.c_int32), ("error_on_flush", C.c_int32),
        ("started", C.c_int32), ("flushed", C.c_int32),
        ("messages", C.c_int64), ("samples", C.c_int64),
    ]


class MockLine(C.Structure):
    _fields_ = [
        ("limit", C.c_int64), ("channels", C.c_int32), ("n_procs", C.c_int32),
        ("value", C.c_double), ("sink_discard", C.c_int32), ("_pad", C.c_int32),
        ("source", MockComponent), ("procs", MockComponent * MAX_PROCS), ("sink", MockComponent),
        ("sink_values", C.POINTER(C.c_double)), ("sink_values_capacity", C.c_int64),
        ("sink_values_len", C.c_int64),
    ]


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so with the committed Makefile."""
    src = os.path.join(ORACLE_DIR, "pipe_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", ORACLE_DIR, "-B" if force else "-s", "liboracle.so"],
                       check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    dp = C.POINTER(C.c_double)
    L.orc_chain_new.restype = C.c_void_p
    L.orc_chain_new.argtypes = [C.c_int32, C.c_int32, C.POINTER(_Stage)]
    L.orc_chain_free.argtypes = [C.c_void_p]
    L.orc_chain_reset.argtypes = [C.c_void_p]
    L.orc_chain_peek_out_frames.restype = C.c_int64
    L.orc_chain_peek_out_frames.argtypes = [C.c_void_p, C.c_int64]
    L.orc_chain_process.restype = C.c_int64
    L.orc_chain_process.argtypes = [C.c_void_p, dp, C.c_int64, dp, C.c_int64]
    L.orc_chain_process_mt.restype = C.c_int64
    L.orc_chain_process_mt.argtypes = [C.c_void_p, dp, C.c_int64, dp, C.c_int64, C.c_int32]
    L.orc_chain_set_stage.restype = C.c_int32
    L.orc_chain_set_stage.argtypes = [C.c_void_p, C.c_int32, C.POINTER(_Stage)]
    L.orc_mix_sum.argtypes = [C.POINTER(dp), C.c_int32, C.c_int64, dp]
    L.orc_meter.argtypes = [dp, C.c_int64, C.c_int32, dp, dp]
    L.orc_source_fill.argtypes = [dp, C.c_int64, C.c_int64, C.c_uint64, C.c_uint64]
    L.orc_pipe_run.restype = C.c_int32
    L.orc_pipe_run.argtypes = [C.c_int64, C.c_int32, C.POINTER(MockLine)]
    L.orc_mock_source_drain.restype = C.c_int32
    L.orc_mock_source_drain.argtypes = [C.c_int64, C.POINTER(MockLine)]
    _lib = L
    return L


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _mk_stage(d: dict, keep: list) -> _Stage:
    s = _Stage()
    kind = d["kind"]
    s.kind = {"copy": 0, "gain": 1, "biquad": 2, "fir": 3, "resample": 4}[kind] if isinstance(kind, str) else kind
    s.gain = float(d.get("gain", 1.0))
    b = d.get("b", (1.0, 0.0, 0.0))
    a = d.get("a", (0.0, 0.0))
    for i in range(3):
        s.b[i] = float(b[i])
    for i in range(2):
        s.a[i] = float(a[i])
    taps = d.get("taps")
    if taps is not None:
        t = np.ascontiguousarray(np.asarray(taps, dtype=np.float64))
        keep.append(t)
        s.taps = _dp(t)
        s.n_taps = t.size
    s.up = int(d.get("up", 0))
    s.down = int(d.get("down", 0))
    return s


class Chain:
    """float64 reference chain with carried state (the DSP specification)."""

    def __init__(self, channels: int, stages: list[dict]):
        self._keep: list = []
        arr = (_Stage * max(1, len(stages)))()
        for i, d in enumerate(stages):
            arr[i] = _mk_stage(d, self._keep)
        self._h = lib().orc_chain_new(channels, len(stages), arr)
        if not self._h:
            raise ValueError("oracle rejected the chain description")
        self.channels = channels
        self.stages = stages

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_chain_free(self._h)
            self._h = None

    def reset(self):
        lib().orc_chain_reset(self._h)

    def peek_out_frames(self, n: int) -> int:
        return int(lib().orc_chain_peek_out_frames(self._h, n))

    def set_stage(self, idx: int, d: dict):
        keep: list = []
        s = _mk_stage(d, keep)
        if lib().orc_chain_set_stage(self._h, idx, C.byref(s)) != 0:
            raise ValueError("oracle rejected set_stage")

    def process(self, x: np.ndarray, threads: int = 1) -> np.ndarray:
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64)).reshape(-1, self.channels)
        n = x.shape[0]
        out = np.empty((max(n, 1), self.channels), dtype=np.float64)
        if threads > 1:
            got = lib().orc_chain_process_mt(self._h, _dp(x), n, _dp(out), out.shape[0], threads)
        else:
            got = lib().orc_chain_process(self._h, _dp(x), n, _dp(out), out.shape[0])
        if got < 0:
            raise RuntimeError("oracle chain failed")
        return out[:got]


def mix_sum(inputs: list[np.ndarray]) -> np.ndarray:
    ins = [np.ascontiguousarray(np.asarray(a, dtype=np.float64)) for a in inputs]
    out = np.empty_like(ins[0])
    ptrs = (C.POINTER(C.c_double) * len(ins))(*[_dp(a) for a in ins])
    lib().orc_mix_sum(ptrs, len(ins), ins[0].size, _dp(out))
    return out


def meter(x: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    frames, ch = x.shape
    peak = np.empty(ch)
    sumsq = np.empty(ch)
    lib().orc_meter(_dp(x), frames, ch, _dp(peak), _dp(sumsq))
    return peak, sumsq


def source_fill(first_index: int, n_values: int, seed: int = 1234, line: int = 0) -> np.ndarray:
    out = np.empty(n_values, dtype=np.float64)
    lib().orc_source_fill(_dp(out), first_index, n_values, seed, line)
    return out


def mock_line(limit=0, channels=1, value=0.0, n_procs=1, discard=True, capture=0, **knobs) -> MockLine:
    """knobs: e.g. source_error_on_flush=1, proc0_error_on_start=1, sink_error_on_make=1."""
    l = MockLine()
    l.limit, l.channels, l.value, l.n_procs = limit, channels, value, n_procs
    l.sink_discard = 1 if discard else 0
    if capture:
        buf = np.zeros(capture, dtype=np.float64)
        l._buf = buf  # keep alive
        l.sink_values = _dp(buf)
        l.sink_values_capacity = capture
    for k, v in knobs.items():
        comp, field = k.split("_", 1)
        target = l.source if comp == "source" else l.sink if comp == "sink" else l.procs[int(comp[4:])]
        setattr(target, field, int(v))
    return l


def pipe_run(buffer_size: int, lines: list[MockLine]) -> tuple[int, list[MockLine]]:
    arr = (MockLine * len(lines))(*lines)
    ret = lib().orc_pipe_run(buffer_size, len(lines), arr)
    out = list(arr)
    for src, dst in zip(lines, out):
        if hasattr(src, "_buf"):
            dst._buf = src._buf
    return int(ret), out


def mock_source_drain(buffer_size: int, line: MockLine) -> int:
    return int(lib().orc_mock_source_drain(buffer_size, C.byref(line)))


class StageList:
    """The oracle's chain as separate per-stage chains run one after the other (the same arithmetic: every Processor keeps its own
    carried state), which makes pipe.InsertProcessor (pipe.go:297) trivial to restate: a new stage, with zero state, is put into
    the list; everything that was there keeps its state."""

    def __init__(self, channels: int, stages: list[dict]):
        self.channels = channels
        self.chains = [Chain(channels, [st]) for st in stages]

    def insert(self, pos: int, stage: dict):
        self.chains.insert(pos, Chain(self.channels, [stage]))

    def process(self, x: np.ndarray, threads: int = 1) -> np.ndarray:
        for c in self.chains:
            x = c.process(x, threads=threads)
        return x
