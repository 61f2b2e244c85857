#!/usr/bin/env python
"""A/B device time of the fused kernel variants at the bench shape (codes resident in HBM, CUDA events, L2 flushed by
the working set itself: 1.09 GB in, 0.9 GB out).  gpurun -- 'python profiles/tools/ab_fused.py [G] [bridges] [frames]'"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import igate4xsoftphonedsp_b200 as ig                      # noqa: E402
from igate4xsoftphonedsp_b200 import _native as N          # noqa: E402
from igate4xsoftphonedsp_b200 import synth                 # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096 // G
F = int(sys.argv[3]) if len(sys.argv) > 3 else 1640
C = B * G
dev = torch.device("cuda", 0)
vp = ig.VoicePath(0)
vp.use_torch_stream()
codes = torch.randint(0, 256, (F, C, 160), dtype=torch.uint8, device=dev)
law = torch.from_numpy(synth.laws(C)).to(dev)
gain = torch.from_numpy(synth.gains(F, B, G)).to(dev)
out_law = torch.from_numpy(synth.out_laws(B)).to(dev)
out = vp.alloc_outputs(F, B, G)


def timed(flags, reps=20):
    for _ in range(3):
        vp.process_batch(codes, law, gain, out_law, G, flags=flags, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        vp.process_batch(codes, law, gain, out_law, G, flags=flags, out=out)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(min(ts))


res = {"lib": os.environ.get("IGD_LIB_PATH", "default"), "shape": f"{B} bridges x {G} legs x {F} frames"}
ref = {k: v.clone() for k, v in vp.process_batch(codes, law, gain, out_law, G, flags=N.F_KERNEL_W, out=out).items()}
for name, fl in (("kernel_w", N.F_KERNEL_W), ("kernel_q", 0), ("kernel_w_again", N.F_KERNEL_W), ("kernel_q_again", 0)):
    res[name + "_ms (median, min)"] = timed(fl)
got = vp.process_batch(codes, law, gain, out_law, G, flags=0, out=out)
res["q_equals_w"] = all(torch.equal(ref[k].view(torch.uint8), got[k].view(torch.uint8)) for k in ref)
bytes_per_bf = G * 160 + 320 + 160 + G * 16      # bench.py BYTES_PER_BF (1184 at G = 4)
res["GBps_q"] = F * B * bytes_per_bf / (res["kernel_q_ms (median, min)"][0] * 1e-3) / 1e9
res["GBps_w"] = F * B * bytes_per_bf / (res["kernel_w_ms (median, min)"][0] * 1e-3) / 1e9
print(json.dumps(res))
