#!/usr/bin/env python
"""Real-time shape: ONE 20 ms tick per call (F = 1) on device-resident buffers.  Prints the device time per
tick (CUDA events) and the host wall time per call through the C ABI, for the RX front-end + arbitration +
fused path chain.  gpurun -- 'python profiles/tools/tick_latency.py'"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import igate4xsoftphonedsp_b200 as ig                      # noqa: E402
from igate4xsoftphonedsp_b200 import _native as N          # noqa: E402
from igate4xsoftphonedsp_b200 import synth                 # noqa: E402

dev = torch.device("cuda", 0)
vp = ig.VoicePath(0)
vp.use_torch_stream()
G = 4
for B in (1024, 16384):
    Cn = B * G
    g = torch.Generator(device=dev).manual_seed(1)
    pk = torch.randint(0, 256, (Cn, 180), dtype=torch.uint8, device=dev, generator=g)
    pk[:, 0] = 0x90
    pk[:, 1] = 8
    law = torch.from_numpy(synth.laws(Cn)).to(dev)
    out_law = torch.from_numpy(synth.out_laws(B)).to(dev)
    st = torch.zeros((Cn, 4), dtype=torch.int32, device=dev)
    legs = torch.zeros((Cn, 2), dtype=torch.int32, device=dev)
    br = torch.zeros((B, 4), dtype=torch.int32, device=dev)
    out = vp.alloc_outputs(1, B, G)

    def tick():
        fields, payload = vp.ed137_parse(pk)
        ev = vp.rx_track(fields.reshape(1, Cn, 4), st)
        gain = vp.gate_arbitrate(ev, legs, br, G, N.ARB_CLIENT_PTT)
        vp.process_batch(payload.reshape(1, Cn, 160), law, gain, out_law, G, out=out)

    for _ in range(20):
        tick()
    torch.cuda.synchronize()
    n = 200
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(n):
        tick()
    b.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e6
    print(f"{Cn} channels, one 20 ms tick: parse + rx_track + gate_arbitrate + fused = "
          f"{a.elapsed_time(b) / n * 1e3:.1f} us on the device, {wall:.1f} us wall per tick "
          f"({20000 / wall:.0f}x faster than real time)")
    # the same chain captured once in a CUDA graph (the library launches on the capturing stream)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        tick()
    for _ in range(5):
        gr.replay()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a.record()
    for _ in range(n):
        gr.replay()
    b.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e6
    print(f"    as one CUDA graph: {a.elapsed_time(b) / n * 1e3:.1f} us on the device, {wall:.1f} us wall per tick")
