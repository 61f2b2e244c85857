#!/usr/bin/env python
"""Real-time shape of the gateway call: ONE 20 ms tick per igd_gateway_process call (packets of every leg in, one finished
packet per bridge out) on device-resident buffers -- device time per tick eager and replayed as one CUDA graph, with the
default walks (one thread per bridge: liveness + arbitration in one kernel) and with IGD_F_WALK_SERIAL (header view /
liveness walk / arbitration as separate kernels).  gpurun -- 'python profiles/tools/gateway_tick_latency.py'"""
import json
import os
import sys

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
res = {}
for B in (1024, 16384):
    Cn = B * G
    g = torch.Generator(device=dev).manual_seed(1)
    pk = torch.randint(0, 256, (1, Cn, 180), dtype=torch.uint8, device=dev, generator=g)
    pk[..., 0] = 0x90
    pk[..., 1] = 8
    law = torch.from_numpy(synth.laws(Cn)).to(dev)
    out_law = torch.from_numpy(synth.out_laws(B)).to(dev)
    rtp12 = torch.from_numpy(synth.rtp12(1, B, [8] * B)).to(dev)
    st = (torch.zeros((Cn, 4), dtype=torch.int32, device=dev), torch.zeros((Cn, 2), dtype=torch.int32, device=dev),
          torch.zeros((B, 4), dtype=torch.int32, device=dev),
          torch.from_numpy(ig.make_state(B, now_ms=0).view(np.int32).reshape(B, 10)).to(dev))
    for name, flags in (("default", 0), ("serial_walks", N.F_WALK_SERIAL)):
        outs = None

        def tick():
            global outs
            outs = vp.gateway_process(pk, law, out_law, st[0], st[1], st[2], rtp12, st[3], mode=N.ARB_CLIENT_PTT, now_ms0=1000,
                                      want=("bmeter",), out=outs, flags=flags)

        for _ in range(10):
            tick()
        torch.cuda.synchronize()
        n = 200
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = vp.launch_count()
        a.record()
        for _ in range(n):
            tick()
        b.record()
        torch.cuda.synchronize()
        eager = a.elapsed_time(b) / n * 1e3
        launches = (vp.launch_count() - l0) / n
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            vp.use_torch_stream()
            tick()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            vp.use_torch_stream()
            tick()
        vp.use_torch_stream()
        for _ in range(10):
            gr.replay()
        torch.cuda.synchronize()
        a.record()
        for _ in range(n):
            gr.replay()
        b.record()
        torch.cuda.synchronize()
        res[f"{Cn}ch_{name}"] = {"eager_us_per_tick": round(eager, 2), "graph_us_per_tick": round(a.elapsed_time(b) / n * 1e3, 2),
                                 "launches_per_tick": launches}
print(json.dumps(res, indent=1))
