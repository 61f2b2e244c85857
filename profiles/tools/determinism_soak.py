#!/usr/bin/env python
"""Soak: the fused path (and then the packet path / state machines) re-run many times on the same device-resident inputs must reproduce its outputs bit for
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

# ---- the packet path and the per-channel state machines: same host inputs, repeated calls, byte-identical outputs
from igate4xsoftphonedsp_b200 import _native as N          # noqa: E402

rng = np.random.default_rng(5)
F, Cn, G = 50, 4096, 4
pk = rng.integers(0, 256, (F * Cn, 180), dtype=np.uint8)
pk[:, 0] = 0x90
pk[:, 1] = rng.choice(np.array([8, 0, 123, 18, 96], np.uint8), F * Cn)
sizes = rng.choice(np.array([180, 180, 180, 20, 12, 100, 8], np.uint32), F * Cn)
rtp12 = rng.integers(0, 256, (F, Cn, 12), dtype=np.uint8)
pay = rng.integers(0, 256, (F, Cn, 160), dtype=np.uint8)
ctl = np.zeros((F, Cn), dtype=N.CTL_DT)
for name in ("pttstatus", "sqlstatus", "pttpriority", "ed137_bssi", "pttid", "callRecorder"):
    ctl[name] = rng.integers(0, 2, (F, Cn)).astype(np.uint8)
present = (rng.integers(0, 4, (F, Cn)) != 0).astype(np.uint8)


def blob(*arrs):
    return b"".join(np.ascontiguousarray(a).tobytes() for a in arrs if a is not None)


def once():
    out = []
    fields, payload = vp.ed137_parse(pk, sizes)
    out.append(blob(fields, payload))
    st = np.zeros(Cn, dtype=N.STATE_DT)
    out.append(blob(*vp.ed137_pack(rtp12, pay, st, ctl, now_ms0=1000, flags=ig.F_REF_QUIRKS), st))
    rxs = np.zeros(Cn, dtype=N.RX_STATE_DT)
    ev = vp.rx_track(fields.reshape(F, Cn), rxs, present, now_ms0=1000)
    out.append(blob(ev, rxs))
    for mode in (N.ARB_CLIENT_PTT, N.ARB_SERVER_BEST):
        legs, br = np.zeros(Cn, dtype=N.ARB_LEG_DT), np.zeros(Cn // G, dtype=N.ARB_BRIDGE_DT)
        gq = vp.gate_arbitrate(ev, legs, br, G, mode=mode)
        out.append(blob(gq, legs, br))
    r = vp.process_batch(pay, synth.laws(Cn), gq, synth.out_laws(Cn // G), G)
    out.append(blob(*vp.event_summary(r["meter"], gq)))
    out.append(blob(vp.wav_images(pay[:, :64], ref_quirks=True)))
    return out


# ---- the packet form of the fused kernel (same slot refill, 4320-byte items): device inputs, many launches
pk_d = torch.from_numpy(pk.reshape(F, Cn, 180)).to(dev)
sz = rng.choice(np.array([180, 180, 180, 20, 100, 179], np.uint32), F * Cn)
fields_h, _ = vp.ed137_parse(pk, sz, want_payload=False)
fields_d = torch.from_numpy(fields_h.view(np.int32).reshape(F * Cn, 4)).to(dev)
law_d, ol_d = torch.from_numpy(synth.laws(Cn)).to(dev), torch.from_numpy(synth.out_laws(Cn // G)).to(dev)
gain_d = torch.from_numpy(synth.gains(F, Cn // G, G).view(np.int16)).to(dev)
refp = None
for r in range(300):
    o = vp.process_packets(pk_d, fields_d, law_d, gain_d, ol_d)
    h = tuple(int(o[k].view(torch.uint8).to(torch.int64).sum()) ^ int((o[k].view(torch.uint8).flatten().to(torch.int64) *
              (torch.arange(o[k].numel() * o[k].element_size(), device=dev) % 1000003 + 1)).sum()) for k in ("mix", "enc", "meter", "bmeter"))
    if refp is None:
        refp = h
    elif h != refp:
        bad += 1
        print(f"MISMATCH process_packets rep {r}")
print("process_packets: 300 runs identical" if bad == 0 else f"mismatches {bad}")

ref = once()
names = ("parse", "pack", "rx_track", "arb_client", "arb_server", "summary", "wav_images")
for r in range(40):
    for nme, a, b in zip(names, ref, once()):
        if a != b:
            bad += 1
            print(f"MISMATCH {nme} rep {r}")
print("packet path / state machines: 40 runs identical" if bad == 0 else f"mismatches {bad}")
print("SOAK", "OK" if bad == 0 else f"FAILED ({bad})")
