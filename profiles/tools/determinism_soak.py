#!/usr/bin/env python
"""Soak: the fused path re-run many times on the same device-resident inputs must reproduce its outputs bit for
bit (a race in the per-warp slot refill / partial scratch hand-over would show up as a rare mismatch).
gpurun -- 'python profiles/tools/determinism_soak.py'"""
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
bad = 0
for G, B, F, reps in ((4, 1024, 400, 300), (4, 37, 11, 2000), (32, 64, 200, 200), (3, 333, 100, 300), (1, 4096, 50, 300), (2, 999, 77, 300)):
    Cn = B * G
    g = torch.Generator(device=dev).manual_seed(G * 1000 + B)
    codes = torch.randint(0, 256, (F, Cn, 160), dtype=torch.uint8, device=dev, generator=g)
    law = torch.from_numpy(synth.laws(Cn)).to(dev)
    out_law = torch.from_numpy(synth.out_laws(B)).to(dev)
    gain = (torch.randint(0, 3, (F, Cn), device=dev, generator=g) == 0).to(torch.int16) * 256
    ref = None
    for r in range(reps):
        out = vp.process_batch(codes, law, gain, out_law, G)
        h = tuple(int(out[k].view(torch.uint8).to(torch.int64).sum()) ^ int((out[k].view(torch.uint8).to(torch.int64) *
                  torch.arange(1, out[k].numel() * out[k].element_size() + 1, device=dev).reshape(out[k].view(torch.uint8).shape) % 1000003).sum())
                  for k in ("mix", "enc", "meter", "bmeter"))
        if ref is None:
            ref = h
        elif h != ref:
            bad += 1
            print(f"MISMATCH G={G} B={B} F={F} rep {r}")
    print(f"G={G} B={B} F={F}: {reps} runs identical" if bad == 0 else f"G={G}: mismatches so far {bad}")
print("SOAK", "OK" if bad == 0 else f"FAILED ({bad})")
