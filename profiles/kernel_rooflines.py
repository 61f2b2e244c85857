#!/usr/bin/env python
"""Roofline table of the stand-alone kernels (everything except the fused kernel, which bench.py
measures): device-resident inputs larger than L2, CUDA events on the launching stream, algorithmic
bytes / best-of-N launch time vs the measured HBM copy peak.
    gpurun -- 'python profiles/kernel_rooflines.py > gpurun_out/kernel_rooflines.json'
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import igate4xsoftphonedsp_b200 as ig                      # noqa: E402
from igate4xsoftphonedsp_b200 import _native as N          # noqa: E402
from igate4xsoftphonedsp_b200 import synth                 # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
dev = torch.device("cuda", 0)
vp = ig.VoicePath(0)
vp.use_torch_stream()
rows = []


def timed(name, fn, nbytes, unit_count, unit, reps=8):
    fn(); fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    gbs = nbytes / best / 1e6
    rows.append({"kernel": name, "ms": round(best, 4), "algorithmic_bytes": int(nbytes), "GB/s": round(gbs, 1),
                 "frac_of_hbm_peak": round(gbs / PEAK, 3), f"{unit}/s": unit_count / best * 1e3})
    print(f"{name:24s} {best:8.3f} ms  {gbs:8.1f} GB/s  {gbs / PEAK:6.1%}", file=sys.stderr)


F, Cn = 1640, 4096                      # 1.07 GB of codes (same shape as the fused bench)
n = F * Cn * 160
g = torch.Generator(device=dev).manual_seed(1)
codes = torch.randint(0, 256, (F, Cn, 160), dtype=torch.uint8, device=dev, generator=g)
law = torch.from_numpy(synth.laws(Cn)).to(dev)
pcm = vp.g711_decode(codes, law)
timed("k_g711_decode", lambda: vp.g711_decode(codes, law), 3 * n, n, "samples")
timed("k_g711_encode", lambda: vp.g711_encode(pcm, law), 3 * n, n, "samples")
timed("k_frame_meter", lambda: vp.frame_meter(pcm), 2 * n + F * Cn * 16, n, "samples")
gain = torch.from_numpy(synth.gains(F, Cn // 4, 4).view(np.int16)).to(dev)
timed("k_mix(G=4)", lambda: vp.mix(pcm, gain, 4), (2 * n) + n // 2, n, "samples")
meter = vp.frame_meter(pcm)
timed("k_event_summary", lambda: vp.event_summary(meter.reshape(F, Cn, 4), gain), F * Cn * 18, F * Cn, "channel-frames")
del pcm, meter
torch.cuda.empty_cache()

# ---- packet path: 180-byte packets, F x C of them
Fp, Cp = 820, 4096
npk = Fp * Cp
pk = torch.randint(0, 256, (npk, 180), dtype=torch.uint8, device=dev, generator=g)
pk[:, 0] = 0x90
pk[:, 1] = 8
timed("k_ed137_parse", lambda: vp.ed137_parse(pk), npk * (180 + 16 + 160), npk, "packets")
timed("k_ed137_fields (header only)", lambda: vp.ed137_parse(pk, want_payload=False), npk * (20 + 16), npk, "packets")
fields, payload = vp.ed137_parse(pk)
st = torch.zeros((Cp, 4), dtype=torch.int32, device=dev)
timed("k_rx_track", lambda: vp.rx_track(fields.reshape(Fp, Cp, 4), st), npk * 24, npk, "packets")
ev = vp.rx_track(fields.reshape(Fp, Cp, 4), st)
legs = torch.zeros((Cp, 2), dtype=torch.int32, device=dev)
br = torch.zeros((Cp // 4, 4), dtype=torch.int32, device=dev)
timed("k_gate_arbitrate", lambda: vp.gate_arbitrate(ev, legs, br, 4, N.ARB_CLIENT_PTT), npk * 6, npk, "leg-frames")
rtp12 = torch.randint(0, 256, (Fp, Cp, 12), dtype=torch.uint8, device=dev, generator=g)
state = torch.from_numpy(ig.make_state(Cp).view(np.uint8).reshape(Cp, 40)).to(dev)
ctl = torch.zeros((Fp, Cp, 8), dtype=torch.uint8, device=dev)
ctl[..., 0] = 1
ctl[..., 2] = 1
timed("ed137_pack (plan+assemble)", lambda: vp.ed137_pack(rtp12, payload.reshape(Fp, Cp, 160), state, ctl=ctl),
      npk * (12 + 160 + 8 + 180 + 4 + 1), npk, "packets", reps=4)
# ---- the same packet count shaped like BASELINE config 5 (65 536 channels): the sequential walks have 16x
# more independent channels to hide their per-tick latency behind
del pk, fields, payload, ev, rtp12, state, ctl
torch.cuda.empty_cache()
Fq, Cq = 52, 65536
nq = Fq * Cq
pk = torch.randint(0, 256, (nq, 180), dtype=torch.uint8, device=dev, generator=g)
pk[:, 0] = 0x90
pk[:, 1] = 8
fields, payload = vp.ed137_parse(pk)
st = torch.zeros((Cq, 4), dtype=torch.int32, device=dev)
timed("k_rx_track @65536ch", lambda: vp.rx_track(fields.reshape(Fq, Cq, 4), st), nq * 24, nq, "packets")
ev = vp.rx_track(fields.reshape(Fq, Cq, 4), st)
legs = torch.zeros((Cq, 2), dtype=torch.int32, device=dev)
br = torch.zeros((Cq // 4, 4), dtype=torch.int32, device=dev)
timed("k_gate_arbitrate @65536ch", lambda: vp.gate_arbitrate(ev, legs, br, 4, N.ARB_CLIENT_PTT), nq * 6, nq, "leg-frames")
rtp12 = torch.randint(0, 256, (Fq, Cq, 12), dtype=torch.uint8, device=dev, generator=g)
state = torch.from_numpy(ig.make_state(Cq).view(np.uint8).reshape(Cq, 40)).to(dev)
ctl = torch.zeros((Fq, Cq, 8), dtype=torch.uint8, device=dev)
ctl[..., 0] = 1
ctl[..., 2] = 1
timed("ed137_pack @65536ch", lambda: vp.ed137_pack(rtp12, payload.reshape(Fq, Cq, 160), state, ctl=ctl),
      nq * (12 + 160 + 8 + 180 + 4 + 1), nq, "packets", reps=4)
print(json.dumps({"hbm_peak_gbs": PEAK, "rows": rows}, indent=1))
