"""ctypes binding of the C-ABI declared in include/pipe_b200.h.

This is the only way Python reaches the CUDA path, and it is the same ABI a Go
host would bind through cgo.  There is NO CPU fallback: if
libpipe_b200.so is missing or fails to load, importing symbols raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PB_LIB") or os.path.join(PKG_DIR, "libpipe_b200.so")  # PB_LIB: a tuning build of the same library

ABI_VERSION = 1
PB_OK = 0
PB_ERR_INVALID, PB_ERR_CUDA, PB_ERR_NO_DEVICE, PB_ERR_NOMEM = -1, -2, -3, -4
PB_ERR_UNSUPPORTED, PB_ERR_CAPACITY, PB_ERR_STATE = -5, -6, -7
PB_F32, PB_F64 = 0, 1
STAGE_COPY, STAGE_GAIN, STAGE_BIQUAD, STAGE_FIR, STAGE_RESAMPLE = range(5)
CHAIN_METER, CHAIN_NO_TENSOR, CHAIN_NO_STREAM = 1, 2, 4

_KINDS = {"copy": 0, "gain": 1, "biquad": 2, "fir": 3, "resample": 4}
_ERR_NAMES = {-1: "PB_ERR_INVALID", -2: "PB_ERR_CUDA", -3: "PB_ERR_NO_DEVICE", -4: "PB_ERR_NOMEM",
              -5: "PB_ERR_UNSUPPORTED", -6: "PB_ERR_CAPACITY", -7: "PB_ERR_STATE"}


class PipeB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{_ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class StageDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("n_taps", C.c_int32), ("up", C.c_int32), ("down", C.c_int32),
        ("_pad", C.c_int32), ("gain", C.c_double), ("b", C.c_double * 3), ("a", C.c_double * 2),
        ("taps", C.POINTER(C.c_double)),
    ]


class ChainDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("dtype", C.c_int32), ("channels", C.c_int32),
        ("sample_rate", C.c_double), ("buffer_frames", C.c_int32), ("max_batch", C.c_int32),
        ("n_stages", C.c_int32), ("flags", C.c_int32), ("stages", C.POINTER(StageDesc)),
    ]


_i32, _i64, _vp, _dbl = C.c_int32, C.c_int64, C.c_void_p, C.c_double
_pi64, _pi32, _pd = C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_double)

# name -> (restype, argtypes); every symbol include/pipe_b200.h declares
SIGNATURES = {
    "pb_chain_create": (_i32, [C.POINTER(ChainDesc), C.POINTER(_vp)]),
    "pb_chain_destroy": (_i32, [_vp]),
    "pb_chain_reset": (_i32, [_vp]),
    "pb_chain_out_properties": (_i32, [_vp, _pi32, _pd]),
    "pb_chain_peek_out_frames": (_i32, [_vp, _i64, _pi64]),
    "pb_chain_process": (_i32, [_vp, _vp, _i64, _vp, _i64, _pi64]),
    "pb_chain_process_batch_device": (_i32, [_vp, _vp, _pi64, _i32, _vp, _i64, _pi64, _vp]),
    "pb_chain_sync": (_i32, [_vp, _vp]),
    "pb_chain_pipeline_depth": (_i32, [_vp]),
    "pb_chain_submit": (_i32, [_vp, _vp, _pi64, _i32, _vp, _i64]),
    "pb_chain_collect": (_i32, [_vp, _pi64, _i32]),
    "pb_chain_set_stage": (_i32, [_vp, _i32, C.POINTER(StageDesc)]),
    "pb_chain_insert_stage": (_i32, [_vp, _i32, C.POINTER(StageDesc)]),
    "pb_chain_meter_read": (_i32, [_vp, _pd, _pd, _pi64]),
    "pb_chain_last_path": (_i32, [_vp, _pi32, _pi64]),
    "pb_source_fill_device": (_i32, [_i32, _i32, _vp, _i64, _i64, C.c_uint64, C.c_uint64, _vp]),
    "pb_meter_device": (_i32, [_i32, _i32, _vp, _i64, _i32, _vp, _vp, _vp]),
    "pb_mix_sum_device": (_i32, [_i32, _i32, C.POINTER(_vp), _i32, _i64, _vp, _vp]),
    "pb_device_count": (_i32, [_pi32]),
    "pb_device_alloc": (_i32, [_i32, _i64, C.POINTER(_vp)]),
    "pb_device_free": (_i32, [_i32, _vp]),
    "pb_host_alloc_pinned": (_i32, [_i64, C.POINTER(_vp)]),
    "pb_host_free_pinned": (_i32, [_vp]),
    "pb_memcpy_h2d": (_i32, [_i32, _vp, _vp, _i64]),
    "pb_memcpy_d2h": (_i32, [_i32, _vp, _vp, _i64]),
    "pb_device_synchronize": (_i32, [_i32]),
    "pb_ipc_export": (_i32, [_i32, _vp, C.POINTER(C.c_uint8)]),
    "pb_ipc_offset": (_i32, [_i32, _vp, _pi64]),
    "pb_ipc_open": (_i32, [_i32, C.POINTER(C.c_uint8), C.POINTER(_vp)]),
    "pb_ipc_close": (_i32, [_i32, _vp]),
    "pb_abi_version": (_i32, []),
    "pb_last_error": (C.c_char_p, []),
}

_lib = None


def lib() -> C.CDLL:
    """Load libpipe_b200.so; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m pipe_b200.build` (needs nvcc). "
            "pipe_b200 has no CPU fallback for the Processor path.")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if L.pb_abi_version() != ABI_VERSION:
        raise ImportError(f"libpipe_b200.so ABI {L.pb_abi_version()} != binding {ABI_VERSION}")
    _lib = L
    return L


def check(code: int) -> None:
    if code != PB_OK:
        msg = lib().pb_last_error()
        raise PipeB200Error(code, msg.decode() if msg else "")


def make_stage(d: dict, keep: list) -> StageDesc:
    s = StageDesc()
    kind = d["kind"]
    s.kind = _KINDS[kind] if isinstance(kind, str) else int(kind)
    s.gain = float(d.get("gain", 1.0))
    b, a = d.get("b", (1.0, 0.0, 0.0)), d.get("a", (0.0, 0.0))
    for i in range(3):
        s.b[i] = float(b[i])
    for i in range(2):
        s.a[i] = float(a[i])
    taps = d.get("taps")
    if taps is not None:
        t = np.ascontiguousarray(np.asarray(taps, dtype=np.float64))
        keep.append(t)
        s.taps = t.ctypes.data_as(_pd)
        s.n_taps = t.size
    s.up, s.down = int(d.get("up", 0)), int(d.get("down", 0))
    return s


def _np_dtype(dtype: int):
    return np.float32 if dtype == PB_F32 else np.float64


class Chain:
    """Owner of one pb_chain handle (one fused run of GPU Processors)."""

    def __init__(self, channels: int, stages: list[dict], *, buffer_frames: int, dtype=np.float32,
                 sample_rate: float = 48000.0, max_batch: int = 1, device: int = 0, flags: int = 0):
        self.dtype = PB_F32 if np.dtype(dtype) == np.float32 else PB_F64
        self.np_dtype = _np_dtype(self.dtype)
        self.channels, self.buffer_frames, self.max_batch, self.device = channels, buffer_frames, max_batch, device
        keep: list = []
        arr = (StageDesc * max(1, len(stages)))()
        for i, d in enumerate(stages):
            arr[i] = make_stage(d, keep)
        desc = ChainDesc(ABI_VERSION, device, self.dtype, channels, float(sample_rate), buffer_frames, max_batch,
                         len(stages), flags, arr)
        h = _vp()
        check(lib().pb_chain_create(C.byref(desc), C.byref(h)))
        self._h = h

    # -- lifecycle -----------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            lib().pb_chain_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self):
        check(lib().pb_chain_reset(self._h))

    def out_properties(self) -> tuple[int, float]:
        ch, sr = _i32(), _dbl()
        check(lib().pb_chain_out_properties(self._h, C.byref(ch), C.byref(sr)))
        return ch.value, sr.value

    def peek_out_frames(self, n: int) -> int:
        out = _i64()
        check(lib().pb_chain_peek_out_frames(self._h, n, C.byref(out)))
        return out.value

    def set_stage(self, idx: int, d: dict):
        keep: list = []
        s = make_stage(d, keep)
        check(lib().pb_chain_set_stage(self._h, idx, C.byref(s)))

    def insert_stage(self, pos: int, d: dict):
        """InsertProcessor on the fused run (pipe.go:297): re-plans the run, existing stages keep their carried state."""
        keep: list = []
        s = make_stage(d, keep)
        check(lib().pb_chain_insert_stage(self._h, pos, C.byref(s)))

    # -- ProcessFunc with host buffers ----------------------------------------
    def process(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(np.asarray(x, dtype=self.np_dtype)).reshape(-1, self.channels)
        n = x.shape[0]
        out = np.empty((max(n, 1), self.channels), dtype=self.np_dtype)
        got = _i64()
        check(lib().pb_chain_process(self._h, x.ctypes.data, n, out.ctypes.data, out.shape[0], C.byref(got)))
        return out[:got.value]

    # -- device-resident batch --------------------------------------------------
    def process_batch_device(self, in_ptr: int, buf_frames: list[int], out_ptr: int, out_capacity_frames: int,
                             stream: int = 0) -> list[int]:
        n = len(buf_frames)
        bf = (C.c_int64 * n)(*buf_frames)
        bo = (C.c_int64 * n)()
        check(lib().pb_chain_process_batch_device(self._h, in_ptr, bf, n, out_ptr, out_capacity_frames, bo, stream))
        return list(bo)

    def sync(self, stream: int = 0):
        check(lib().pb_chain_sync(self._h, stream))

    # -- pipelined host path ------------------------------------------------------
    def submit(self, in_ptr: int, buf_frames: list[int], out_ptr: int, out_capacity_frames: int):
        n = len(buf_frames)
        bf = (C.c_int64 * n)(*buf_frames)
        check(lib().pb_chain_submit(self._h, in_ptr, bf, n, out_ptr, out_capacity_frames))

    def collect(self, n_buffers: int) -> list[int]:
        bo = (C.c_int64 * n_buffers)()
        check(lib().pb_chain_collect(self._h, bo, n_buffers))
        return list(bo)

    def meter_read(self) -> tuple[np.ndarray, np.ndarray, int]:
        peak, sumsq, fr = np.empty(self.channels), np.empty(self.channels), _i64()
        check(lib().pb_chain_meter_read(self._h, peak.ctypes.data_as(_pd), sumsq.ctypes.data_as(_pd), C.byref(fr)))
        return peak, sumsq, fr.value

    def last_path(self) -> tuple[int, int]:
        p, k = _i32(), _i64()
        check(lib().pb_chain_last_path(self._h, C.byref(p), C.byref(k)))
        return p.value, k.value


class DeviceBuffer:
    """cudaMalloc'd buffer through the ABI (what a torch-free host would use)."""

    def __init__(self, nbytes: int, device: int = 0):
        p = _vp()
        check(lib().pb_device_alloc(device, nbytes, C.byref(p)))
        self.ptr, self.nbytes, self.device = p.value, nbytes, device

    def upload(self, a: np.ndarray):
        a = np.ascontiguousarray(a)
        check(lib().pb_memcpy_h2d(self.device, self.ptr, a.ctypes.data, a.nbytes))

    def download(self, shape, dtype) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        check(lib().pb_memcpy_d2h(self.device, out.ctypes.data, self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            lib().pb_device_free(self.device, self.ptr)
            self.ptr = None

    __del__ = free


class PinnedBuffer:
    def __init__(self, nbytes: int):
        p = _vp()
        check(lib().pb_host_alloc_pinned(nbytes, C.byref(p)))
        self.ptr, self.nbytes = p.value, nbytes

    def array(self, shape, dtype) -> np.ndarray:
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        assert n <= self.nbytes
        buf = (C.c_char * n).from_address(self.ptr)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self.ptr:
            lib().pb_host_free_pinned(self.ptr)
            self.ptr = None

    __del__ = free


def source_fill(ptr: int, dtype: int, first_index: int, n_values: int, seed: int = 1234, line: int = 0,
                device: int = 0, stream: int = 0):
    check(lib().pb_source_fill_device(device, dtype, ptr, first_index, n_values, seed, line, stream))


def meter_device(ptr: int, dtype: int, frames: int, channels: int, peak_ptr: int, sumsq_ptr: int,
                 device: int = 0, stream: int = 0):
    check(lib().pb_meter_device(device, dtype, ptr, frames, channels, peak_ptr, sumsq_ptr, stream))


def mix_sum(ptrs: list[int], dtype: int, n_values: int, out_ptr: int, device: int = 0, stream: int = 0):
    arr = (_vp * len(ptrs))(*ptrs)
    check(lib().pb_mix_sum_device(device, dtype, arr, len(ptrs), n_values, out_ptr, stream))


def device_count() -> int:
    n = _i32()
    code = lib().pb_device_count(C.byref(n))
    return n.value if code == PB_OK else 0
