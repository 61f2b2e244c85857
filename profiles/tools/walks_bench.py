#!/usr/bin/env python
"""Device times (CUDA events, steady state) of the three per-channel walks -- k_rx_track, k_gate_arbitrate,
k_ed137_plan (+ assembly) -- on the chain workload of bench.py --chain.  IGD_LIB_PATH selects an experimental build.
gpurun -- 'python profiles/tools/walks_bench.py [bridges] [frames]'"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import igate4xsoftphonedsp_b200 as ig                      # noqa: E402
from igate4xsoftphonedsp_b200 import _native as N          # noqa: E402
from igate4xsoftphonedsp_b200 import synth                 # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
F = int(sys.argv[2]) if len(sys.argv) > 2 else 1640
G = 4
C = B * G
dev = torch.device("cuda", 0)
vp = ig.VoicePath(0)
vp.use_torch_stream()
law = torch.from_numpy(synth.laws(C)).to(dev)
gate = torch.from_numpy(synth.gates(F, B, G).reshape(F, C).astype(np.uint8)).to(dev)
pk = torch.zeros((F, C, 180), dtype=torch.uint8, device=dev)
pk[..., 0] = 0x90
pk[..., 1] = torch.where(law == 0, 8, 0).to(torch.uint8).view(1, C)
prio = (1 + torch.arange(C, device=dev) % G % 4).view(1, C)
pk[..., 16] = (gate.to(torch.int64) * (prio << 5)).to(torch.uint8)
fields, _ = vp.ed137_parse(pk.view(F * C, 180), want_payload=False)
fields = fields.view(F, C, 4)
rx_state = torch.zeros((C, 4), dtype=torch.int32, device=dev)
legs = torch.zeros((C, 2), dtype=torch.int32, device=dev)
bridges = torch.zeros((B, 4), dtype=torch.int32, device=dev)
rtp12 = torch.from_numpy(synth.rtp12(F, B, [8] * B)).to(dev)
tx_state = torch.from_numpy(ig.make_state(B, now_ms=0).view(np.int32).reshape(B, 10)).to(dev)
ctl_np = np.zeros((F, B), dtype=N.CTL_DT)
ctl_np["pttstatus"] = 1
ctl_np["pttpriority"] = 1
ctl = torch.from_numpy(ctl_np.view(np.int32).reshape(F, B, 2)).to(dev)
enc = torch.zeros((F, B, 160), dtype=torch.uint8, device=dev)


def timed(fn, reps=6):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


ev = vp.rx_track(fields, rx_state)
res = {"lib": os.environ.get("IGD_LIB_PATH", "default"), "shape": f"{B} bridges x {G} legs x {F} ticks",
       "rx_track_ms": timed(lambda: vp.rx_track(fields, rx_state)),
       "gate_arbitrate_ms": timed(lambda: vp.gate_arbitrate(ev, legs, bridges, G, N.ARB_CLIENT_PTT)),
       "pack_ms (plan + assemble)": timed(lambda: vp.ed137_pack(rtp12, enc, tx_state, ctl))}
print(json.dumps(res))
