import os, sys
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import _oracle as orc
from pipe_b200 import abi, design
st = design.config_stages("chain4")
ch, bf = 256, 1600
gpu, cpu = abi.Chain(ch, st, buffer_frames=bf, flags=abi.CHAIN_METER), orc.Chain(ch, st)
for b in range(2):
    x = orc.source_fill(b * bf * ch, bf * ch).reshape(bf, ch)
    ref = cpu.process(x, threads=8)
    y = gpu.process(x.astype(np.float32))
    print(b, gpu.last_path(), float((np.abs(y - ref).max(axis=0) / np.abs(ref).max(axis=0)).max()))
