"""Randomised parity soak of the tcgen05 chain kernel against the CPU oracle (tools/k2_soak.py): random filters, channel counts,
call lengths (whole and partial last tiles, batches and single buffers), per-channel levels from 0 to -60 dBFS that jump between
calls.  The bar is the one of the parity tests: 1e-6 of every channel's own peak per call.  Two seeds here; the tool takes any."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [3, 8])
def test_k2_randomised_soak(seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "k2_soak.py"), "12", str(seed)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ok" in r.stdout.splitlines()[-1]


def test_k3_randomised_soak():
    # the streaming kernels: gain + biquad, 1..160 channels, f32 and f64, both sides of the two-sweep threshold, ragged buffers
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "k3_soak.py"), "16", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ok" in r.stdout.splitlines()[-1]


def test_k1_randomised_soak():
    # the generic fused tile kernel: random stage lists in one or two fused segments, resampler ratios, 1..130 channels, f32 and f64
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "k1_soak.py"), "24", "4"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ok" in r.stdout.splitlines()[-1]


def test_lifecycle_randomised_soak():
    # random sequences of process / set_stage (mutations) / reset through the C-ABI on the three kernel families, meter included
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ops_soak.py"), "16", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ok" in r.stdout.splitlines()[-1]


def test_insert_processor_randomised_soak():
    # random gains / biquads / FIRs spliced into running fused runs (pb_chain_insert_stage) against the oracle's stage list
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "insert_soak.py"), "16", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ok" in r.stdout.splitlines()[-1]


def test_pipelined_host_path_randomised_soak():
    # pb_chain_submit / pb_chain_collect with two batches in flight, pinned and pageable host memory, synchronous calls in between
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "pipeline_soak.py"), "12", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ok" in r.stdout.splitlines()[-1]
