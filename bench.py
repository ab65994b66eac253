#!/usr/bin/env python
"""bench.py -- headline benchmark of the Processor hot path.

    python bench.py --gpus N --steps K --warmup W            # own arm (CUDA)
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm

Metric (BASELINE.json): Msamples/s through the 4-stage Processor chain
[gain, 257-tap FIR, biquad, resample 48k->44.1k] at 48 kHz x 1024 ch float32,
plus the achieved fraction of the HBM roofline.  A "sample" is one input
channel-sample at the Source output; frames/s = samples/s / channels.

A step is one pass of the hot path over one batch of `batch_buffers` buffers of
`buffer_frames` frames (one fused kernel launch; per-buffer `processed` counts
are still produced).  The batch (320 MiB in, 294 MiB out) is larger than the
126 MB L2, so consecutive steps cannot hit in L2.

N > 1 (torchrun, one rank per GPU): each rank runs an independent Line of the
same shape on its own GPU -- configs[3], no data-path collective -- and the
value is the total over ranks divided by the slowest rank's device time.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

CHANNELS = 1024
BUFFER_FRAMES = 4096
BATCH_BUFFERS = 20            # 20 * 4096 frames = 512 tiles of 160 frames; acc returns to 0 every 5 buffers
SAMPLE_RATE = 48000.0
BYTES_PER_SAMPLE = 4.0 + 4.0 * 147.0 / 160.0   # read 4 B, write 4*147/160 B
METRIC = "Msamples/s through 4-stage Processor chain @48kHz x1024ch"
WORKLOAD = "configs[2]: Source->[gain, 257-tap FIR, biquad, resample 48k->44.1k]->Sink, 48 kHz x 1024 ch float32"


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_path: int):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return json.load(f).get(str(kernel_path), {}).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 8 and parts[0].isdigit() and int(parts[0]) == self.idx:
                self.rows.append(parts)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                power.append(float(r[3]))
            except ValueError:
                continue
            for name, val in zip(names, r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


# ---------------------------------------------------------------- CPU arms --

def cpu_chain_throughput(n_buffers: int, threads: int, repeats: int = 1):
    """Oracle (C restatement, float64 like the reference) on host cores; returns Msamples/s."""
    import _oracle as orc
    from pipe_b200 import design
    chain = orc.Chain(CHANNELS, design.config_stages("chain4"))
    x = orc.source_fill(0, BUFFER_FRAMES * CHANNELS).reshape(BUFFER_FRAMES, CHANNELS)
    chain.process(x, threads=threads)  # warm
    t0 = time.perf_counter()
    for _ in range(repeats):
        for _ in range(n_buffers):
            chain.process(x, threads=threads)
    dt = time.perf_counter() - t0
    return repeats * n_buffers * BUFFER_FRAMES * CHANNELS / dt / 1e6, dt


def cpu_thread_per_stage(n_buffers: int):
    """The reference's own execution shape (run.go:173-196 + fitting.Async, fitting.go:56-60): one thread per component,
    cap-1 queues between them, every Processor a single-threaded float64 loop (the C restatement, one stage per chain).
    The oracle calls release the GIL, so the stage threads really overlap.  Returns Msamples/s."""
    import queue

    import _oracle as orc
    from pipe_b200 import design
    stages = design.config_stages("chain4")
    procs = [orc.Chain(CHANNELS, [st]) for st in stages]
    x = orc.source_fill(0, BUFFER_FRAMES * CHANNELS).reshape(BUFFER_FRAMES, CHANNELS)
    qs = [queue.Queue(maxsize=1) for _ in range(len(procs) + 1)]

    def source():
        for _ in range(n_buffers):
            qs[0].put(x)
        qs[0].put(None)

    def processor(i):
        while True:
            buf = qs[i].get()
            if buf is None:
                qs[i + 1].put(None)
                return
            qs[i + 1].put(procs[i].process(buf))

    threads = [threading.Thread(target=source)] + [threading.Thread(target=processor, args=(i,)) for i in range(len(procs))]
    t0 = time.perf_counter()
    for t in threads:
        t.start()
    frames_out = 0
    while True:  # the Sink
        buf = qs[-1].get()
        if buf is None:
            break
        frames_out += len(buf)
    dt = time.perf_counter() - t0
    for t in threads:
        t.join()
    return n_buffers * BUFFER_FRAMES * CHANNELS / dt / 1e6, dt, len(threads) + 1


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step = 4  # buffers per step: a bounded sample of the batch
    times = []
    import _oracle as orc
    from pipe_b200 import design
    chain = orc.Chain(CHANNELS, design.config_stages("chain4"))
    x = orc.source_fill(0, BUFFER_FRAMES * CHANNELS).reshape(BUFFER_FRAMES, CHANNELS)
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        for _ in range(per_step):
            chain.process(x, threads=threads)
        if s >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    samples = args.steps * per_step * BUFFER_FRAMES * CHANNELS
    value = samples / total / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "channels": CHANNELS, "buffer_frames": BUFFER_FRAMES,
                   "note": "the reference is Go and cannot be built here (no Go toolchain): this arm times the C "
                           "restatement of its float64 loop (oracle/pipe_oracle.c), NOT the Go code"},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": threads, "kind": "port",
                         "sample": f"{per_step} buffers of {BUFFER_FRAMES}x{CHANNELS} float64 per step, "
                                   f"channels split over {threads} host threads"},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------- CUDA arm --

def hbm_bound_runs(local, stream):
    """[gain, biquad] (configs[1]: 64 ch x 4096-frame buffers; and at 1024 ch) through the streaming kernels (K3):
    8 B of HBM traffic per sample, batches larger than L2, CUDA events on the launching stream."""
    import torch

    from pipe_b200 import abi, design
    peak, _ = measured_peak()
    out = []
    for name, ch, nb in (("configs[1]: Source->[gain, biquad-IIR]->Sink, 48 kHz x 64 ch float32, 4096-sample buffer", 64, 320),
                         ("[gain, biquad-IIR] at 48 kHz x 1024 ch float32, 4096-sample buffer", 1024, 20)):
        frames = BUFFER_FRAMES * nb
        chain = abi.Chain(ch, design.config_stages("gain_biquad"), buffer_frames=BUFFER_FRAMES, max_batch=nb, device=local)
        x = torch.empty((frames, ch), dtype=torch.float32, device=f"cuda:{local}")
        y = torch.empty_like(x)
        abi.source_fill(x.data_ptr(), abi.PB_F32, 0, frames * ch, seed=1234, line=0, device=local)
        sizes = [BUFFER_FRAMES] * nb
        for _ in range(3):
            chain.process_batch_device(x.data_ptr(), sizes, y.data_ptr(), frames, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        steps = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            chain.process_batch_device(x.data_ptr(), sizes, y.data_ptr(), frames, stream=stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        chain.sync(stream.cuda_stream)
        ms = e0.elapsed_time(e1) / steps
        gbs = 8.0 * frames * ch / (ms * 1e-3) / 1e9
        out.append({"workload": name, "value": frames * ch / (ms * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms,
                    "batch_buffers": nb, "kernel_path": {1: "generic fused tile kernel", 3: "streaming kernel"}.get(chain.last_path()[0]),
                    "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                 "algorithmic_bytes_per_launch": 8.0 * frames * ch}})
        chain.close()
        del x, y
    return out


def pin_to_gpu_numa_node(local: int) -> str:
    """Pin this rank's host threads (and therefore the pages of the pinned staging it allocates next) to the CPUs NVML
    reports as local to its GPU.  Eight ranks staging 640 MB per step each through one NUMA node was the e2e scaling limit."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"NVML-local CPUs ({len(cpus)} of {os.cpu_count()})"
        return "NVML reported no local CPUs inside this cgroup: unchanged"
    except Exception as e:  # no NVML, no permission: report and carry on
        return f"unchanged ({type(e).__name__})"


def pcie_ceiling(nbytes: int, barrier):
    """Pinned H2D and D2H copies of the e2e batch size running at the same time on two streams: the box's ceiling for the
    pipelined host path, measured on every rank at once.  Returns GB/s per direction."""
    import torch
    n = nbytes // 4
    h_in, h_out = torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory()
    d_in, d_out = torch.empty(n, dtype=torch.float32, device="cuda"), torch.empty(n, dtype=torch.float32, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        return reps * n * 4 / (time.perf_counter() - t0) / 1e9
    run(1)
    barrier()
    return run(4)


def dropin_e2e(local: int, n_buffers: int = 12):
    """The drop-in case, shaped like the cgo shim (go/pipeb200.go): the pipe hands over ONE pageable float64 buffer per
    ProcessFunc call (pipe.go:394,437,438); the shim marshals it into its pinned staging (ReadFloat64, converting to float32 in
    float32 mode), calls pb_chain_process (H2D, kernels, D2H, synchronous) and marshals the result back (WriteFloat64).
    Both compute modes, and for each the marshalling on one thread and cut over 8 threads by frame ranges (the shim's
    Options.Workers: goroutines over signal.Floating.Slice views), timed on the host clock around the whole call sequence."""
    from concurrent.futures import ThreadPoolExecutor
    from pipe_b200 import abi, design
    import _oracle as orc  # input generation only (outside the timed region)
    x64 = orc.source_fill(0, BUFFER_FRAMES * CHANNELS).reshape(BUFFER_FRAMES, CHANNELS)   # pageable float64, as signal.Floating
    out64 = np.empty_like(x64)
    res = {}
    pool = ThreadPoolExecutor(max_workers=8)

    def par_copy(dst, src, rows, workers):
        if workers <= 1:
            np.copyto(dst[:rows], src[:rows], casting="same_kind")
            return
        per = (rows + workers - 1) // workers
        list(pool.map(lambda lo: np.copyto(dst[lo:min(rows, lo + per)], src[lo:min(rows, lo + per)], casting="same_kind"),
                      range(0, rows, per)))   # numpy releases the GIL inside the copy
    for mode, dt in (("float64 compute (go: Chain)", np.float64), ("float32 compute (go: ChainWith{Float32: true})", np.float32)):
        chain = abi.Chain(CHANNELS, design.config_stages("chain4"), buffer_frames=BUFFER_FRAMES, dtype=dt, device=local)
        pin_in = abi.PinnedBuffer(x64.size * np.dtype(dt).itemsize)
        pin_out = abi.PinnedBuffer(x64.size * np.dtype(dt).itemsize)
        a_in, a_out = pin_in.array(x64.shape, dt), pin_out.array(x64.shape, dt)
        got = abi._i64()
        lib = abi.lib()
        for workers in (1, 8):
            def one():
                par_copy(a_in, x64, BUFFER_FRAMES, workers)              # ReadFloat64 (+ narrowing in float32 mode)
                abi.check(lib.pb_chain_process(chain._h, pin_in.ptr, BUFFER_FRAMES, pin_out.ptr, BUFFER_FRAMES, abi.C.byref(got)))
                n = got.value
                par_copy(out64, a_out, n, workers)                       # WriteFloat64 (+ widening)
                return n
            chain.reset()
            for _ in range(3):
                one()
            t0 = time.perf_counter()
            for _ in range(n_buffers):
                n = one()
            dt_s = time.perf_counter() - t0
            t1 = time.perf_counter()
            for _ in range(n_buffers):   # the library's share: the same calls without the marshalling copies
                abi.check(lib.pb_chain_process(chain._h, pin_in.ptr, BUFFER_FRAMES, pin_out.ptr, BUFFER_FRAMES, abi.C.byref(got)))
            lib_s = time.perf_counter() - t1
            res[f"{mode}, marshalling on {workers} thread{'s' if workers > 1 else ''}"] = {
                "value": n_buffers * BUFFER_FRAMES * CHANNELS / dt_s / 1e6, "unit": "Msamples/s",
                "ms_per_buffer": 1e3 * dt_s / n_buffers, "ms_per_buffer_in_pb_chain_process": 1e3 * lib_s / n_buffers,
                "kernel_path": chain.last_path()[0], "out_frames_last": int(n)}
        chain.close()
        pin_in.free()
        pin_out.free()
    pool.shutdown()
    return res


def per_buffer_run(local: int, stream, n_buffers: int = 40):
    """One 4096 x 1024 buffer per pb_chain_process_batch_device call, device-resident: the reference's ProcessFunc granularity
    (pipe.go:438) without the host copies."""
    import torch
    from pipe_b200 import abi, design
    peak, _ = measured_peak()
    chain = abi.Chain(CHANNELS, design.config_stages("chain4"), buffer_frames=BUFFER_FRAMES, device=local)
    x = torch.empty((BUFFER_FRAMES * n_buffers, CHANNELS), dtype=torch.float32, device=f"cuda:{local}")
    y = torch.empty((BUFFER_FRAMES, CHANNELS), dtype=torch.float32, device=f"cuda:{local}")
    abi.source_fill(x.data_ptr(), abi.PB_F32, 0, x.numel(), seed=1234, device=local)
    step = BUFFER_FRAMES * CHANNELS * 4
    _, l0 = chain.last_path()
    for b in range(5):
        chain.process_batch_device(x.data_ptr() + b * step, [BUFFER_FRAMES], y.data_ptr(), BUFFER_FRAMES, stream=stream.cuda_stream)
    torch.cuda.synchronize()
    _, l1 = chain.last_path()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for b in range(5, n_buffers):
        chain.process_batch_device(x.data_ptr() + b * step, [BUFFER_FRAMES], y.data_ptr(), BUFFER_FRAMES, stream=stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    chain.sync(stream.cuda_stream)
    ms = e0.elapsed_time(e1) / (n_buffers - 5)
    gbs = BUFFER_FRAMES * CHANNELS * BYTES_PER_SAMPLE / (ms * 1e-3) / 1e9
    out = {"workload": "configs[2], ONE 4096 x 1024 buffer per call (device-resident; the reference's ProcessFunc granularity)",
           "value": BUFFER_FRAMES * CHANNELS / (ms * 1e-3) / 1e6, "unit": "Msamples/s", "us_per_buffer": 1e3 * ms,
           "launches_per_call": (l1 - l0) / 5.0, "kernel_path": chain.last_path()[0],
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak}}
    chain.close()
    return out


def fan_in_run(rank, local, world, stream):
    """configs[4]: `world` Lines x 256 ch, each through the 4-stage chain on its own GPU, then the fan-in sum onto rank 0 --
    once as an NCCL reduce, once as ONE mixer kernel on rank 0 that pulls the peers' buffers over NVLink while it sums
    (pb_mix_sum_device on pb_ipc_open'ed pointers).  Device-timed on rank 0; checked against the oracle's sum."""
    import torch
    import torch.distributed as dist
    from pipe_b200 import abi, design, shard
    ch, bf, nb = 256, 4000, 4        # 4000 = 25 tiles of 160 frames
    frames = bf * nb
    dev = torch.device("cuda", local)
    stages = design.config_stages("chain4")
    chain = abi.Chain(ch, stages, buffer_frames=bf, max_batch=nb, device=local)
    x = torch.empty((frames, ch), dtype=torch.float32, device=dev)
    y = torch.zeros((frames, ch), dtype=torch.float32, device=dev)
    abi.source_fill(x.data_ptr(), abi.PB_F32, 0, frames * ch, seed=1234, line=rank, device=local)
    sptr = stream.cuda_stream
    counts = chain.process_batch_device(x.data_ptr(), [bf] * nb, y.data_ptr(), frames, stream=sptr)
    chain.sync(sptr)
    n_out = sum(counts)
    mine = y[:n_out].clone()
    reps = 10

    def timed(fn):
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        dist.barrier()
        return e0.elapsed_time(e1) / reps
    red = mine.clone()
    shard.fan_in_reduce(red, dst=0)                    # result kept for the parity check (and warm-up)
    scratch = mine.clone()
    ms_nccl = timed(lambda: dist.reduce(scratch, dst=0))   # accumulates on rank 0; only the time matters here
    out = torch.zeros_like(mine)
    pf = shard.PeerFanIn(mine.data_ptr(), mine.numel(), abi.PB_F32, local, dst=0)
    pf.sum_into(out.data_ptr(), stream=sptr)
    torch.cuda.synchronize()
    dist.barrier()
    ms_peer = None
    if rank == 0:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            abi.mix_sum(pf.peer_ptrs, abi.PB_F32, mine.numel(), out.data_ptr(), device=local, stream=sptr)
        e1.record(stream)
        torch.cuda.synchronize()
        ms_peer = e0.elapsed_time(e1) / reps
    dist.barrier()
    pf.close()
    res = None
    if rank == 0:
        import _oracle as orc
        ref = None
        for line in range(world):
            cpu = orc.Chain(ch, stages)
            xs = orc.source_fill(0, frames * ch, line=line).reshape(frames, ch)
            r = np.concatenate([cpu.process(xs[i * bf:(i + 1) * bf], threads=os.cpu_count() or 1) for i in range(nb)])
            ref = r if ref is None else ref + r
        pk = np.abs(ref).max(axis=0)
        line_mb = mine.numel() * 4 / 1e6
        res = {"workload": f"configs[4]: {world} Lines x {ch} ch -> [4-stage chain] -> fan-in sum on rank 0",
               "frames_out": int(n_out), "bytes_per_line": mine.numel() * 4,
               "nvlink_bytes_algorithmic": (world - 1) * mine.numel() * 4,
               "nccl_reduce": {"ms": ms_nccl, "err_over_peak": float((np.abs(red.cpu().numpy() - ref).max(axis=0) / pk).max()),
                               "gbs_into_rank0": (world - 1) * line_mb / ms_nccl},
               "peer_mixer_kernel": {"ms": ms_peer, "err_over_peak": float((np.abs(out.cpu().numpy() - ref).max(axis=0) / pk).max()),
                                     "gbs_into_rank0": (world - 1) * line_mb / ms_peer,
                                     "note": "one pb_mix_sum_device launch on rank 0: loads of the 3 peers' buffers over NVLink + sum + store"}}
    chain.close()
    return res


def run_cuda(args):
    import torch
    import torch.distributed as dist

    from pipe_b200 import abi, design

    rank, local, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; pipe_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = pin_to_gpu_numa_node(local)      # before any pinned staging is allocated
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    stages = design.config_stages("chain4")
    nb = args.batch_buffers
    frames = BUFFER_FRAMES * nb
    sizes = [BUFFER_FRAMES] * nb
    chain = abi.Chain(CHANNELS, stages, buffer_frames=BUFFER_FRAMES, max_batch=nb, device=local,
                      sample_rate=SAMPLE_RATE, flags=abi.CHAIN_METER if args.meter else 0)
    # Source output resident in HBM before the timed region (one Line per rank, its own seed stream)
    x = torch.empty((frames, CHANNELS), dtype=torch.float32, device=dev)
    y = torch.empty((frames, CHANNELS), dtype=torch.float32, device=dev)
    abi.source_fill(x.data_ptr(), abi.PB_F32, 0, frames * CHANNELS, seed=1234, line=rank, device=local)
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream

    def step():
        return chain.process_batch_device(x.data_ptr(), sizes, y.data_ptr(), frames, stream=sptr)

    for _ in range(max(args.warmup, 3)):
        counts = step()
    chain.sync(sptr)
    _, launches0 = chain.last_path()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record(stream)
    for s in range(args.steps):
        counts = step()
        evs[s + 1].record(stream)
    barrier()
    chain.sync(sptr)
    clk = clocks.stop() if rank == 0 else {}
    step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    total_ms = evs[0].elapsed_time(evs[-1])
    path, launches1 = chain.last_path()
    out_frames = sum(counts)

    # ---- end to end through the C-ABI with HOST buffers (pinned), copies inside the timed region ----
    e2e_steps = max(2, min(args.steps, 6))
    pin_in = [torch.empty((frames, CHANNELS), dtype=torch.float32).pin_memory() for _ in range(2)]
    pin_out = [torch.empty((frames, CHANNELS), dtype=torch.float32).pin_memory() for _ in range(2)]
    xin = x.cpu()
    for b in pin_in:
        b.copy_(xin)
    del xin
    chain.reset()

    def e2e_run(n):
        for s in range(n + 1):
            if s < n:
                chain.submit(pin_in[s & 1].data_ptr(), sizes, pin_out[s & 1].data_ptr(), frames)
            if s >= 1:
                oc = chain.collect(nb)
        return oc

    e2e_run(2)  # warm: allocates the slots
    pcie_gbs = pcie_ceiling(frames * CHANNELS * 4, barrier)   # every rank at once: the box's ceiling at this N
    barrier()
    t0 = time.perf_counter()
    oc = e2e_run(e2e_steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    result_probe = float(pin_out[(e2e_steps - 1) & 1][0, 0])  # device->host result actually read
    if world > 1:
        t = torch.tensor([total_ms, e2e_s * 1e3, -pcie_gbs], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms, pcie_min = float(t[0]), float(t[1]), -float(t[2])
        t = torch.tensor([pcie_gbs], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        pcie_sum = float(t[0])
    else:
        e2e_ms, pcie_min, pcie_sum = e2e_s * 1e3, pcie_gbs, pcie_gbs
    # a step moves 4 B per sample in and 4*147/160 B out at the same time: the input direction is the longer one
    e2e_ceiling = world * pcie_min * 1e9 / 4.0 / 1e6
    fan_in = fan_in_run(rank, local, world, stream) if (world == 4 and not args.no_secondary) else None

    # ---- the HBM-bound runs (configs[1] and the same run at 1024 ch) on the streaming kernels, device-resident; reported
    #      beside the headline so that every BASELINE.json config that fits one GPU has a measured roofline fraction
    secondary = hbm_bound_runs(local, stream) if (rank == 0 and world == 1 and not args.no_secondary) else None
    if secondary is not None:
        secondary.append(per_buffer_run(local, stream))
    dropin = dropin_e2e(local) if (rank == 0 and world == 1 and not args.no_secondary) else None

    samples_step = frames * CHANNELS               # per rank
    value = world * samples_step * args.steps / (total_ms * 1e-3) / 1e6
    e2e_value = world * samples_step * e2e_steps / (e2e_ms * 1e-3) / 1e6

    if rank == 0:
        peak, peak_src = measured_peak()
        kern_ms = statistics.mean(step_ms)       # one fused launch per step on this stream
        achieved = samples_step * BYTES_PER_SAMPLE / (kern_ms * 1e-3) / 1e9
        cpu_threads = os.cpu_count() or 1
        # the CPU baseline is timed at N = 1 only (the driver's scaling runs reuse that line)
        cpu_val, cpu_dt = cpu_chain_throughput(args.cpu_buffers, cpu_threads) if world == 1 else (None, 0.0)
        tps = cpu_thread_per_stage(max(2, min(8, args.cpu_buffers // 8))) if (world == 1 and args.cpu_buffers >= 16) else None
        line = {
            "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": WORKLOAD, "channels": CHANNELS, "buffer_frames": BUFFER_FRAMES, "batch_buffers": nb,
                "frames_per_step": frames, "out_frames_per_step": out_frames,
                "frames_per_s": value * 1e6 / CHANNELS, "lines": world, "sharding": "one independent Line per GPU",
                "l2": f"inputs larger than L2 ({frames * CHANNELS * 4 / 2**20:.0f} MiB in per step)",
                "timing": "CUDA events on the launching stream; max over ranks",
                "kernel_path": {1: "generic fused tile kernel", 2: "tcgen05/TMA chain kernel"}.get(path, str(path)),
                "fused_meter_sink": bool(args.meter),
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(path) if nb == BATCH_BUFFERS else None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": samples_step * BYTES_PER_SAMPLE,
                         "kernel_ms": kern_ms},
            "cpu_baseline": {"value": cpu_val, "unit": "Msamples/s", "cores": cpu_threads, "kind": "port",
                             "sample": (f"{args.cpu_buffers} buffers of {BUFFER_FRAMES}x{CHANNELS} float64 through the "
                                        f"C restatement (oracle/pipe_oracle.c), {cpu_threads} threads, {cpu_dt:.1f} s")
                             if world == 1 else "not timed at N > 1: the CPU baseline is measured by the N = 1 run",
                             "thread_per_stage": ({"value": tps[0], "unit": "Msamples/s", "threads": tps[2], "seconds": tps[1],
                                                   "note": "the reference's own shape: one thread per component, cap-1 queues "
                                                           "(run.go:173-196, fitting.go:56-60), each Processor a single-threaded "
                                                           "float64 loop; the channel-split number above is the faster, conservative baseline"}
                                                  if tps else None)},
            "e2e": {"value": e2e_value, "unit": "Msamples/s",
                    "h2d_bytes_per_step": frames * CHANNELS * 4, "d2h_bytes_per_step": out_frames * CHANNELS * 4,
                    "steps": e2e_steps, "path": "pb_chain_submit/collect, PINNED float32 host buffers, batches of "
                                                f"{nb} buffers, 2 batches in flight (the best case for a host path, not the drop-in case: see dropin)",
                    "result_probe": result_probe, "cpu_affinity": affinity,
                    "pcie_ceiling": {"gbs_per_direction_per_rank_min": pcie_min, "gbs_per_direction_all_ranks": pcie_sum,
                                     "msamples_per_s_all_ranks": e2e_ceiling, "frac_of_ceiling": e2e_value / e2e_ceiling,
                                     "how": "pinned H2D + D2H copies of the batch size on two streams, all ranks at once, "
                                            "inside this run"},
                    "dropin": dropin},
            "gpu_launches": int(launches1 - launches0),
            "clocks": clk,
        }
        if secondary is not None:
            line["secondary"] = secondary
        if fan_in is not None:
            line["fan_in"] = fan_in
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--batch-buffers", type=int, default=BATCH_BUFFERS)
    ap.add_argument("--cpu-buffers", type=int, default=300, help="bounded CPU-baseline sample (buffers): ~10 s on 16 threads")
    ap.add_argument("--meter", action="store_true", help="fuse the peak/RMS meter sink into the chain kernel")
    ap.add_argument("--no-secondary", action="store_true", help="skip the HBM-bound secondary runs (configs[1])")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
