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


def run_cuda(args):
    import torch
    import torch.distributed as dist

    from pipe_b200 import abi, design

    rank, local, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; pipe_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
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
    barrier()
    t0 = time.perf_counter()
    oc = e2e_run(e2e_steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    result_probe = float(pin_out[(e2e_steps - 1) & 1][0, 0])  # device->host result actually read
    if world > 1:
        t = torch.tensor([total_ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = float(t[0]), float(t[1])
    else:
        e2e_ms = e2e_s * 1e3

    # ---- the HBM-bound runs (configs[1] and the same run at 1024 ch) on the streaming kernels, device-resident; reported
    #      beside the headline so that every BASELINE.json config that fits one GPU has a measured roofline fraction
    secondary = hbm_bound_runs(local, stream) if (rank == 0 and world == 1 and not args.no_secondary) else None

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
                             if world == 1 else "not timed at N > 1: the CPU baseline is measured by the N = 1 run"},
            "e2e": {"value": e2e_value, "unit": "Msamples/s",
                    "h2d_bytes_per_step": frames * CHANNELS * 4, "d2h_bytes_per_step": out_frames * CHANNELS * 4,
                    "steps": e2e_steps, "path": "pb_chain_submit/collect, pinned host buffers, 2 batches in flight",
                    "result_probe": result_probe},
            "gpu_launches": int(launches1 - launches0),
            "clocks": clk,
        }
        if secondary is not None:
            line["secondary"] = secondary
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
