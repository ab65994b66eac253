"""The C++ host mirror of the reference API (pipe_b200/host/pipe.hpp): compiled here with g++ against
libpipe_b200.so, checked against the reference's own plumbing goldens (no GPU needed: mock components
only), and -- on a GPU -- run as Line{Source, [gpu::Chain], Sink} through the C-ABI against the oracle.
Golden sources (paths relative to /root/reference): pipe_test.go:84-105,337-404,437-457,
mock/mock_test.go:69-92,133-146, run.go:78-132."""
import os
import subprocess

import numpy as np
import pytest

import _oracle as orc
from pipe_b200 import build as pb_build
from pipe_b200 import design

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_check.cpp")


@pytest.fixture(scope="module")
def host_check(tmp_path_factory):
    lib = pb_build.build_lib()
    exe = str(tmp_path_factory.mktemp("cpp") / "host_check")
    libdir = os.path.dirname(lib)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-pthread", SRC, "-o", exe, f"-L{libdir}", "-lpipe_b200",
           f"-Wl,-rpath,{libdir}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def _kv(text):
    return dict(line.split("=", 1) for line in text.splitlines() if "=" in line)


def test_cpp_host_matches_the_reference_plumbing_goldens(host_check):
    res = subprocess.run([host_check, "plumbing"], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr
    kv = _kv(res.stdout)
    # pipe_test.go:104-105
    assert (kv["golden862.err"], kv["golden862.messages"], kv["golden862.samples"]) == ("0", "862", "441344")
    assert kv["golden862.proc_messages"] == "862" and kv["golden862.hooks"] == "111111"
    # pipe_test.go:337,363,394,399,404: messages per Limit at bufferSize 512, sink frame totals equal Limit
    for limit, msgs in ((1040, 3), (1640, 4), (3048, 6), (4096, 8)):
        assert (kv[f"limit{limit}.err"], kv[f"limit{limit}.messages"], kv[f"limit{limit}.samples"]) == ("0", str(msgs), str(limit))
    # mock_test.go:69-92 (11 / 5 -> 3 calls) and the pass-through values (mock_test.go:133-146)
    assert (kv["values.err"], kv["values.messages"], kv["values.all_ones"]) == ("0", "3", "1")
    # run.go:113-132: one buffer per line per iteration, a line is flushed and removed at EOF
    assert kv["two.messages"] == "3,6" and kv["two.samples"] == "1040,3048" and kv["two.flushed"] == "11"
    # pipe_test.go:437-457 / run.go:192,221: "error running: %w"; everything started is flushed
    assert kv["procerr.exec"] == "error running: mock error" and kv["procerr.flushed"] == "111"
    # line.go:72-74: allocator errors abort binding, nothing starts
    assert kv["makeerr.msg"] == "processor: mock error" and kv["makeerr.started"] == "0"
    # run.go:78-99: a start error flushes what was started (the flag is set before the error, mock.go:49-58)
    assert kv["starterr.msg"] == "error starting lines: mock error" and kv["starterr.flags"] == "111000"
    # pipe.New + Start + Wait (async, run.go:173-196)
    assert (kv["async.err"], kv["async.messages"], kv["async.samples"], kv["async.flushed"]) == ("0", "862", "441344", "111")
    assert kv["asyncerr.msg"] == "error running: mock error"


def _stages_file(path, stages):
    with open(path, "w") as f:
        for s in stages:
            k = s["kind"]
            if k == "copy":
                f.write("copy\n")
            elif k == "gain":
                f.write(f"gain {s['gain']!r}\n")
            elif k == "biquad":
                f.write("biquad " + " ".join(repr(float(v)) for v in list(s["b"]) + list(s["a"])) + "\n")
            elif k == "fir":
                f.write(f"fir {len(s['taps'])} " + " ".join(repr(float(v)) for v in s["taps"]) + "\n")
            elif k == "resample":
                f.write(f"resample {s['up']} {s['down']} {len(s['taps'])} " + " ".join(repr(float(v)) for v in s["taps"]) + "\n")


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,channels,buffer,frames", [("gain_biquad", 64, 4096, 3 * 4096 + 100), ("chain4", 128, 1600, 4 * 1600),
                                                       ("chain4", 8, 4096, 5 * 4096 + 17)])
def test_cpp_host_line_through_the_gpu_chain(host_check, tmp_path, cfg, channels, buffer, frames):
    stages = design.config_stages(cfg)
    x = orc.source_fill(0, frames * channels).reshape(frames, channels)
    cpu = orc.Chain(channels, stages)
    refs = [cpu.process(x[i:i + buffer]) for i in range(0, frames, buffer)]
    sf, fin, fout = str(tmp_path / "stages.txt"), str(tmp_path / "in.f32"), str(tmp_path / "out.f32")
    _stages_file(sf, stages)
    x.astype(np.float32).tofile(fin)
    res = subprocess.run([host_check, "gpu", sf, fin, fout, str(channels), str(frames), str(buffer), "48000"],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    kv = _kv(res.stdout)
    assert kv["gpu.err"] == "0", kv
    # integer bookkeeping, bit-exact: messages, frames per message (3763, 3763, ... for chain4), output properties
    lens = [int(v) for v in kv["gpu.lens"].split(",") if v]
    assert lens == [len(r) for r in refs]
    assert int(kv["gpu.messages"]) == len(refs) and int(kv["gpu.samples"]) == sum(len(r) for r in refs)
    assert int(kv["out.channels"]) == channels
    rate = 48000.0 * (147.0 / 160.0 if cfg == "chain4" else 1.0)
    assert abs(float(kv["out.sample_rate"]) - rate) < 1e-6
    y = np.fromfile(fout, dtype=np.float32).reshape(-1, channels)
    ref = np.concatenate(refs)
    assert y.shape == ref.shape
    pos = 0
    for i, r in enumerate(refs):  # 1e-6 of the per-channel peak per buffer
        if len(r) >= 256:
            err = np.abs(y[pos:pos + len(r)] - r).max(axis=0)
            assert (err <= 1e-6 * np.abs(r).max(axis=0)).all(), f"buffer {i}: {np.max(err / np.abs(r).max(axis=0)):.3e}"
        pos += len(r)
