#!/usr/bin/env python
"""SURVEY.md 8(d) "with RTP" accounting: the whole path on device-resident buffers, raw ED-137 packets in ->
ED-137 packets out, at the bench shape (4096 channels x 1640 frames):

    k_ed137_fields -> k_rx_track -> k_gate_arbitrate -> k_fused_w<packets> -> ed137_pack (plan + assemble)
    (IGD_CHAIN_TWO_CALL=1: k_ed137_parse_tile with payload -> ... -> k_fused_w on the payload array)

Prints one JSON object: channel-samples/s of the chain, per-stage device times (CUDA events), the
algorithmic figure (1284 B per bridge-frame = 1184 core + 20 (G + 1) header bytes) and the bytes the chain
actually moves (it materialises the payload / event / gain arrays between the stages).
gpurun -- 'python profiles/tools/rtp_chain.py'"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import igate4xsoftphonedsp_b200 as ig                      # noqa: E402
from igate4xsoftphonedsp_b200 import _native as N          # noqa: E402
from igate4xsoftphonedsp_b200 import synth                 # noqa: E402

B, G, F = int(os.environ.get("IGD_CHAIN_BRIDGES", "1024")), 4, int(os.environ.get("IGD_CHAIN_FRAMES", "1640"))
C = B * G
dev = torch.device("cuda", 0)
vp = ig.VoicePath(0)
vp.use_torch_stream()

# ---- inbound packets [F][C][180]: PJSIP-style header, ED-137 extension, the leg's G.711 payload
law_np = synth.laws(C)
law = torch.from_numpy(law_np).to(dev)
out_law = torch.from_numpy(synth.out_laws(B)).to(dev)
codes = vp.g711_encode(synth.pcm_noise_tone_torch(F, C, dev), law)
gate = torch.from_numpy(synth.gates(F, B, G).reshape(F, C).astype(np.uint8)).to(dev)
pk = torch.zeros((F, C, 180), dtype=torch.uint8, device=dev)
pk[..., 0] = 0x90
pk[..., 1] = torch.where(law == 0, 8, 0).to(torch.uint8).view(1, C)           # PT 8 = PCMA, 0 = PCMU
f_idx = torch.arange(F, device=dev).view(F, 1)
pk[..., 2] = ((f_idx >> 8) & 255).to(torch.uint8)
pk[..., 3] = (f_idx & 255).to(torch.uint8)
pk[..., 12], pk[..., 13], pk[..., 15] = 0x01, 0x67, 0x01
prio = (1 + torch.arange(C, device=dev) % G % 4).view(1, C)                    # PTT type = 1 + g % 4 while the gate is open
pk[..., 16] = (gate.to(torch.int64) * (prio << 5)).to(torch.uint8)             # bits 31-29 of the big-endian word
pk[..., 20:] = codes
del codes
pk = pk.view(F * C, 180)

rx_state = torch.zeros((C, 4), dtype=torch.int32, device=dev)
legs = torch.zeros((C, 2), dtype=torch.int32, device=dev)
bridges = torch.zeros((B, 4), dtype=torch.int32, device=dev)
out = vp.alloc_outputs(F, B, G)
rtp12 = torch.from_numpy(synth.rtp12(F, B, np.where(synth.out_laws(B) == 0, 8, 0).astype(np.uint8))).to(dev)
tx_state = torch.zeros((B, N.STATE_DT.itemsize), dtype=torch.uint8, device=dev)
ctl_np = np.zeros((F, B), dtype=N.CTL_DT)
ctl_np["pttstatus"] = 1
ctl_np["pttpriority"] = 1
ctl = torch.from_numpy(ctl_np.view(np.uint8).reshape(F, B, N.CTL_DT.itemsize)).to(dev)

names = ("parse", "rx_track", "gate_arbitrate", "fused", "pack")
res = {}


FUSED_PKT = os.environ.get("IGD_CHAIN_TWO_CALL", "") == ""      # default: the fused kernel reads the packets itself
pk3 = pk.view(F, C, 180)


def chain(ev=None):
    def mark(i):
        if ev is not None:
            ev[i].record()
    mark(0)
    fields, payload = vp.ed137_parse(pk, want_payload=not FUSED_PKT)
    mark(1)
    events = vp.rx_track(fields.view(F, C, 4), rx_state)
    mark(2)
    gain = vp.gate_arbitrate(events, legs, bridges, G, N.ARB_CLIENT_PTT)
    mark(3)
    if FUSED_PKT:
        vp.process_packets(pk3, fields, law, gain, out_law, G, out=out)
    else:
        vp.process_batch(payload.view(F, C, 160), law, gain, out_law, G, out=out)
    mark(4)
    pkts, sizes, bm = vp.ed137_pack(rtp12, out["enc"], tx_state, ctl)
    mark(5)
    res["last"] = (fields, events, gain, pkts, sizes)


for _ in range(3):
    chain()
torch.cuda.synchronize()
steps = 10
acc = np.zeros(5)
tot = 0.0
for _ in range(steps):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    chain(ev)
    torch.cuda.synchronize()
    acc += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(5)])
    tot += ev[0].elapsed_time(ev[5])
acc /= steps
tot /= steps
fields, events, gain, pkts, sizes = res["last"]
open_frac = float((gain != 0).float().mean())
moved = F * ((C * (64 + 16) + B * (1184 + 80 + 64) if FUSED_PKT       # header sectors in, fields out; fused reads whole packets + fields
              else C * (180 + 160 + 16) + B * 1184)                     # parse: packets in, payload + fields out; fused on the payload
             + C * (16 + 8)                   # rx_track: fields in, events out
             + C * (8 + 2)                    # gate_arbitrate: events in, gains out
             + B * (12 + 160 + 8 + 180 + 5))  # pack: header + payload + control in, packet + size + level out
alg = F * B * (1184 + 20 * (G + 1))
peak = 6552.6
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
print(json.dumps({
    "form": "fields-only parse + igd_process_packets" if FUSED_PKT else "igd_ed137_parse (payload) + igd_process_batch",
    "workload": f"{C} channels ({B} bridges x {G} legs) x {F} frames, raw 180-byte ED-137 packets in, {B} x {F} packets out",
    "ms_per_step": tot, "channel_samples_per_s": C * F * 160 / (tot * 1e-3),
    "stage_ms": {n: round(float(t), 4) for n, t in zip(names, acc)},
    "algorithmic_bytes": alg, "algorithmic_GBps": alg / tot / 1e6, "frac_of_hbm_peak_algorithmic": alg / tot / 1e6 / peak,
    "bytes_moved_by_the_chain": moved, "moved_GBps": moved / tot / 1e6, "frac_of_hbm_peak_moved": moved / tot / 1e6 / peak,
    "legs_open_after_arbitration": open_frac,
    "all_packets_out_180_bytes": bool((sizes == 180).all()),
}, indent=1))
