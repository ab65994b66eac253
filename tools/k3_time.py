"""Development timing of K3 (gain + biquad) at a given channel count: one sweep with the decoupled look-back against two sweeps
with a scan in between (PB_ST_TWO_SWEEPS=<max channel groups>, 0 = never), and fewer resident CTAs (PB_ST_CTAS_PER_SM).
Run on a GPU box:    python tools/k3_time.py 64 320     # channels, buffers of 4096 frames per launch"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 3:   # child: one measurement
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import numpy as np
    import hbm_chains as h
    from pipe_b200 import design
    ch, nb = int(sys.argv[1]), int(sys.argv[2])
    r = h.run(f"TWO_SWEEPS={os.environ.get('PB_ST_TWO_SWEEPS')} CTAS_PER_SM={os.environ.get('PB_ST_CTAS_PER_SM')}",
              design.config_stages("gain_biquad"), ch, 4096, nb, np.float32)
else:
    ch, nb = sys.argv[1], sys.argv[2]
    import json
    for ts, cps in (("0", "5"), ("0", "3"), ("0", "2"), ("64", "5"), ("64", "3")):
        out = subprocess.run([sys.executable, __file__, ch, nb, "child"], env=dict(os.environ, PB_ST_TWO_SWEEPS=ts, PB_ST_CTAS_PER_SM=cps),
                             stdout=subprocess.PIPE, check=False).stdout.decode().strip().splitlines()
        try:
            r = json.loads(out[-1])
            print(f"{ch} ch x {nb} buffers  {r['case']:32s} {r['ms_per_launch']:.4f} ms  {r['GBps']:.0f} GB/s  frac {r['hbm_frac']:.3f}", flush=True)
        except Exception:
            print("no result:", out[-3:], flush=True)
