#!/usr/bin/env python
"""Fused kernel (G = 4, config-3 shape) vs how many gates are open and which gains they carry.
gpurun -- 'python profiles/tools/gate_density_sweep.py'"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import igate4xsoftphonedsp_b200 as ig                      # noqa: E402
from igate4xsoftphonedsp_b200 import synth                 # noqa: E402

dev = torch.device("cuda", 0)
vp = ig.VoicePath(0)
vp.use_torch_stream()
F, B, G = 1640, 1024, 4
Cn = B * G
g = torch.Generator(device=dev).manual_seed(1)
codes = torch.randint(0, 256, (F, Cn, 160), dtype=torch.uint8, device=dev, generator=g)
law = torch.from_numpy(synth.laws(Cn)).to(dev)
out_law = torch.from_numpy(synth.out_laws(B)).to(dev)
out = vp.alloc_outputs(F, B, G)
nbytes = F * B * 1184


def run(name, gain):
    gain = gain.reshape(F, Cn).contiguous()
    f = lambda: vp.process_batch(codes, law, gain, out_law, G, out=out)
    f(); f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(f"{name:58s} {best:.3f} ms  {nbytes / best / 1e6:6.0f} GB/s  {nbytes / best / 1e6 / 6552.6:5.1%}")


z = lambda: torch.zeros((F, B, G), dtype=torch.int16, device=dev)
run("all gates shut (metering only)", z())
t = z(); t[:, :, 0] = 256
run("1 of 4 open at 2.0 (one talker per bridge)", t)
run("2 of 4 open at 2.0, 0.5 s cadence (bench.py)", torch.from_numpy(synth.gains(F, B, G).view(np.int16)).to(dev).reshape(F, B, G))
t = z(); t[:] = 256
run("4 of 4 open at 2.0", t)
t = z(); t[:, :, 0] = 256; t[:, ::2, 1] = 256
run("per-bridge patterns differ inside a warp (1 or 2 open)", t)
t = z(); t[:, :, 0] = 13; t[:, :, 1] = 64
run("2 of 4 open at sidetone 0.1 / 0.5 (general path)", t)
