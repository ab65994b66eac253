"""float64 K3 against the oracle for a filter with poles near the unit circle (20 Hz high-pass at 48 kHz), one sweep and two
sweeps: the error of the time-parallel form comes from the conditioning of the TDF-II basis, not from the scan.
Run on a GPU box:  python tools/k3_f64_cond.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1:
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import _oracle as orc
    from pipe_b200 import abi, design
    ch, bf, nb = 64, 4096, 34
    for kind, f0 in (("highpass", 20.0), ("highpass", 200.0), ("lowpass", 8000.0)):
        b, a = design.biquad(kind, f0, 48000.0, q=0.707)
        stages = [{"kind": "gain", "gain": 0.8}, {"kind": "biquad", "b": b, "a": a}]
        x = orc.source_fill(0, bf * nb * ch).reshape(bf * nb, ch)
        ref = orc.Chain(ch, stages).process(x, threads=os.cpu_count() or 1)
        gpu = abi.Chain(ch, stages, buffer_frames=bf, max_batch=nb, dtype=np.float64)
        d_in, d_out = abi.DeviceBuffer(x.nbytes), abi.DeviceBuffer(x.nbytes)
        d_in.upload(x)
        gpu.process_batch_device(d_in.ptr, [bf] * nb, d_out.ptr, len(x))
        gpu.sync()
        y = d_out.download(x.shape, np.float64)
        err = (np.abs(y - ref).max(axis=0) / np.abs(ref).max(axis=0)).max()
        print(f"PB_ST_TWO_SWEEPS={os.environ.get('PB_ST_TWO_SWEEPS'):3s} {kind} {f0:6.0f} Hz: launches {gpu.last_path()[1]}, worst err / peak = {err:.2e}", flush=True)
else:
    for ts in ("0", "4"):
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, PB_ST_TWO_SWEEPS=ts))
