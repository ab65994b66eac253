"""Drive tools/tc_probe.cu on a B200 and compare with numpy (development probe for K2)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from pipe_b200 import design  # noqa: E402

SO = os.path.join(HERE, "libtc_probe.so")


def build():
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-shared",
                    "-Xcompiler", "-fPIC", "-o", SO, os.path.join(HERE, "tc_probe.cu")], check=True)


def split16(v, fixed=False):
    hi = np.rint(v).astype(np.float16) if fixed else v.astype(np.float16)
    lo = (v - hi.astype(np.float64)).astype(np.float16)
    return hi, lo


def toeplitz_cores(h_scaled):
    """T_e[n_i][k_i] = gtap(8e + n_i - k_i + 1), e = -21..53 (75 cores of 8x8)."""
    def gtap(t):
        return h_scaled[t] if 0 <= t <= 256 else 0.0
    out = np.zeros((75, 8, 8))
    for e in range(-21, 54):
        for n_i in range(8):
            for k_i in range(8):
                out[e + 21, n_i, k_i] = gtap(8 * e + n_i - k_i + 1)
    return out


def v_cores(V_scaled):
    """V core (q, kb', nb)[n_i][k_i] = V[8nb+n_i][16q + 8(1-kb') + k_i]; layout [q][kb'][nb][n_i][k_i]."""
    out = np.zeros((27, 2, 2, 8, 8))
    for q in range(27):
        for kbp in range(2):
            for nb in range(2):
                for n_i in range(8):
                    v = 8 * nb + n_i
                    if v < V_scaled.shape[0]:
                        out[q, kbp, nb, n_i, :] = V_scaled[v, 16 * q + 8 * (1 - kbp): 16 * q + 8 * (1 - kbp) + 8]
    return out


def main():
    if "--build-only" in sys.argv:
        build()
        return
    if not os.path.exists(SO):
        build()
    rng = np.random.default_rng(3)
    h = design.lowpass_fir(257, 20000 / 48000)
    x = rng.uniform(-1, 1, (432, 128)).astype(np.float32)
    V = rng.uniform(-1, 1, (2, 432)) * 0.1
    lib = C.CDLL(SO)
    lib.tc_probe_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int]
    for split, fixed, amp in ((1, 2, 1.0), (1, 2, 0.01), (1, 2, 3.0), (0, 2, 1.0)):
        sx, sh = (2.0 ** 10, 2.0 ** 11) if fixed else (2.0 ** 8, 2.0 ** 10)
        xfixed = fixed == 1
        t_hi, t_lo = split16(toeplitz_cores(h * sh), fixed)
        v_hi, v_lo = split16(v_cores(V * sh), fixed)
        tbl = np.concatenate([t_hi.ravel(), t_lo.ravel(), v_hi.ravel(), v_lo.ravel()]).astype(np.float16)
        xa = (x * amp).astype(np.float32)
        out = np.zeros((128, 384), dtype=np.float32)
        rc = lib.tc_probe_run(xa.ctypes.data, tbl.ctypes.data, out.ctypes.data, C.c_float(sx), split, fixed)
        print("==== split_acc", split, "fixed_hi", fixed, "amp", amp, "rc", rc)
        o64 = out.astype(np.float64)
        D = (o64[:, :192] + o64[:, 192:]) / (sx * sh)
        report(D, xa, h, V, sx, sh, o64, fixed, xfixed)


def report(D, x, h, V, sx, sh, o64, fixed, xfixed):
    xd = x.astype(np.float64)
    xd = x.astype(np.float64)
    # reference: y[n][c] = sum_t h[t] * x[n + 257 - t][c], n = 0..174
    ref = np.zeros((175, 128))
    for n in range(175):
        idx = n + 257 - np.arange(257)
        ref[n] = h @ xd[idx]
    got = D[:, :175].T
    err = np.abs(got - ref)
    print("FIR  : max|err| %.3e  max|ref| %.3e  ratio %.3e  (bar 1e-6)" % (err.max(), np.abs(ref).max(), err.max() / np.abs(ref).max()))
    print("       per-channel worst err/peak %.3e ; rms err %.3e" % ((err.max(axis=0) / np.abs(ref).max(axis=0)).max(), np.sqrt((err ** 2).mean())))
    # decompose: exact value of the three fp16-piece products (float64) vs the true result vs the GPU
    xs = xd * sx
    x_hi = (np.rint(xs) if xfixed else xs).astype(np.float16).astype(np.float64)
    x_lo = (xs - x_hi).astype(np.float16).astype(np.float64)
    hs = h * sh
    h_hi = (np.rint(hs) if fixed else hs).astype(np.float16).astype(np.float64)
    h_lo = (hs - h_hi).astype(np.float16).astype(np.float64)
    ref3 = np.zeros((175, 128))
    refhh = np.zeros((175, 128))
    for n in range(175):
        idx = n + 257 - np.arange(257)
        refhh[n] = h_hi @ x_hi[idx]
        ref3[n] = refhh[n] + h_hi @ x_lo[idx] + h_lo @ x_hi[idx]
    ref3 /= sx * sh
    pk = np.abs(ref).max()
    print("split error  (ref3 - true)/peak : max %.3e rms %.3e" % (np.abs(ref3 - ref).max() / pk, np.sqrt(((ref3 - ref) ** 2).mean()) / pk))
    d = got - ref3
    print("accum error  (gpu - ref3)/peak  : max %.3e rms %.3e mean %.3e" % (np.abs(d).max() / pk, np.sqrt((d ** 2).mean()) / pk, d.mean() / pk))
    dm = o64[:, :175].T / (sx * sh) - refhh / (sx * sh)
    print("main acc only (gpu_main - exact hi*hi)/peak : max %.3e rms %.3e" % (np.abs(dm).max() / pk, np.sqrt((dm ** 2).mean()) / pk))
    print("accum error sign vs value sign  : mean(d*sign(ref3))/peak %.3e  (negative => truncation toward zero)" % ((d * np.sign(ref3)).mean() / pk))
    zref = V @ xd            # [2][128]
    zgot = D[:, 176:178].T
    print("V    : max|err| %.3e  max|ref| %.3e" % (np.abs(zgot - zref).max(), np.abs(zref).max()))
    print("pad  : col175 max %.3e, cols 178..191 max %.3e" % (np.abs(D[:, 175]).max(), np.abs(D[:, 178:]).max()))
    if err.max() > 1e-3:
        # help debugging: which (n) columns are off
        bad = np.argwhere(err > 1e-3)
        print("bad entries:", len(bad), "first", bad[:10].tolist())
        print("got[0,:8]", got[0, :8], "ref[0,:8]", ref[0, :8])


if __name__ == "__main__":
    main()
