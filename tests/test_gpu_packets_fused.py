"""GPU parity of igd_process_packets (the fused path reading the codes straight out of the raw
ED-137 packets) against the oracle on the payload igd_ed137_parse extracts from the same packets:
transport_rtp_cb's rule (TransportAdapter.cpp:240-316: payload = bytes 20..size; a keep-alive never
reaches the stream, so the leg is silent -- as is a truncated / dropped / absent packet) composed
with the decode -> meter -> mix -> encode path."""
import numpy as np
import pytest
import torch

import oracle_py as O
import igate4xsoftphonedsp_b200 as ig
from igate4xsoftphonedsp_b200 import _native as N
from test_gpu_fused import check

pytestmark = pytest.mark.gpu
G = 4


def make(F, B, seed, ragged, gains=(0, 0, 256)):
    rng = np.random.default_rng(seed)
    Cn = B * G
    pk = rng.integers(0, 256, (F, Cn, 180), dtype=np.uint8)
    pk[..., 0] = 0x90
    pk[..., 1] = rng.choice(np.array([8, 0, 8, 0, 123, 18, 96], np.uint8), (F, Cn))
    if ragged:
        sizes = rng.choice(np.array([180, 180, 180, 180, 20, 12, 100, 21, 179, 37, 8, 200, 1100, 0], np.uint32), (F, Cn))
    else:
        sizes = np.full((F, Cn), 180, np.uint32)
    law = rng.integers(0, 2, Cn).astype(np.uint8)
    out_law = rng.integers(0, 2, B).astype(np.uint8)
    gain = rng.choice(np.array(gains, np.uint16), (F, Cn))
    return pk, sizes, law, gain, out_law


def no_audio(pk, sizes):
    """a packet that is not a whole G.711 audio frame: the leg is silent on that tick (IGD_GAIN_NO_AUDIO rule:
    pt 0 / 8 and at least 160 payload bytes -- igd_ed137_parse reports payload_len = min(size - 20, 160) -- and not
    dropped, i.e. size - 20 < 1024)"""
    pt = pk[..., 1] & 0x7F
    return ~(((pt == 0) | (pt == 8)) & (sizes >= 180) & (sizes < 1044))


def want_of(vp, pk, sizes, law, gain, out_law, signed=0):
    F, Cn, _ = pk.shape
    fields, payload = vp.ed137_parse(pk.reshape(F * Cn, 180), sizes.reshape(-1))
    g = np.where(no_audio(pk, sizes), gain | N.GAIN_NO_AUDIO, gain).astype(np.uint16)
    return fields.reshape(F, Cn), O.process_batch(payload.reshape(F, Cn, 160), law, g, out_law, G,
                                                  signed_char=signed, threads=8)


@pytest.mark.parametrize("F,B,ragged", [(7, 5, True), (3, 1, True), (1, 1, False), (40, 64, False), (50, 333, True),
                                        (77, 250, False)])
@pytest.mark.parametrize("kern", [0, N.F_KERNEL_W])      # quarter-lane kernel (default) and its predecessor
def test_packets_in_equals_parse_then_process_batch(vp, F, B, ragged, kern):
    pk, sizes, law, gain, out_law = make(F, B, 100 * F + B, ragged)
    fields, want = want_of(vp, pk, sizes, law, gain, out_law)
    check(vp.process_packets(pk, fields, law, gain, out_law, flags=kern), want)


@pytest.mark.parametrize("kern", [0, N.F_KERNEL_W])
def test_packets_in_general_gains_and_signed_char(vp, kern):
    pk, sizes, law, gain, out_law = make(9, 13, 5, True, gains=(0, 13, 64, 128, 256, 300))
    fields, want = want_of(vp, pk, sizes, law, gain, out_law, signed=1)
    check(vp.process_packets(pk, fields, law, gain, out_law, flags=ig.F_SIGNED_CHAR | kern), want)


@pytest.mark.parametrize("kern", [0, N.F_KERNEL_W])
def test_packets_in_device_pointers_repeated_launches(vp, kern):
    """many items per warp, L2-resident packets, device buffers: every launch equals the oracle."""
    F, B = 60, 400
    pk, sizes, law, gain, out_law = make(F, B, 77, True)
    fields, want = want_of(vp, pk, sizes, law, gain, out_law)
    dev = "cuda:0"
    d = [torch.from_numpy(a).to(dev) for a in (pk, fields.view(np.int32).reshape(F, B * G, 4), law,
                                               gain.view(np.int16), out_law)]
    for rep in range(4):
        r = vp.process_packets(*d, flags=kern)
        torch.cuda.synchronize()
        got = {"mix": r["mix"].cpu().numpy(), "enc": r["enc"].cpu().numpy(),
               "meter": r["meter"].cpu().numpy().view(ig.METER_DT).reshape(F, B * G),
               "bmeter": r["bmeter"].cpu().numpy().view(ig.BRIDGE_DT).reshape(F, B)}
        check(got, want)


@pytest.mark.parametrize("legs,B,F", [(2, 9, 7), (3, 5, 6), (8, 6, 5), (32, 2, 4), (1, 11, 3)])
def test_packets_in_any_leg_count(vp, legs, B, F):
    """leg counts other than 4 (e.g. the 32 inbound call slots of a CLIENT softphone, roip_ed137.cpp:141-150):
    same rule, payload extracted inside the library"""
    rng = np.random.default_rng(legs * 100 + B)
    Cn = B * legs
    pk = rng.integers(0, 256, (F, Cn, 180), dtype=np.uint8)
    pk[..., 0] = 0x90
    pk[..., 1] = rng.choice(np.array([8, 0, 8, 0, 123, 96], np.uint8), (F, Cn))
    sizes = rng.choice(np.array([180, 180, 180, 20, 100, 0, 200], np.uint32), (F, Cn))
    law = rng.integers(0, 2, Cn).astype(np.uint8)
    out_law = rng.integers(0, 2, B).astype(np.uint8)
    gain = rng.choice(np.array([0, 256, 256, 64], np.uint16), (F, Cn))
    fields, payload = vp.ed137_parse(pk.reshape(F * Cn, 180), sizes.reshape(-1))
    g = np.where(no_audio(pk, sizes), gain | N.GAIN_NO_AUDIO, gain).astype(np.uint16)
    want = O.process_batch(payload.reshape(F, Cn, 160), law, g, out_law, legs)
    check(vp.process_packets(pk, fields.reshape(F, Cn), law, gain, out_law, legs=legs), want)
    dev = "cuda:0"
    d = [torch.from_numpy(a).to(dev) for a in (pk, fields.view(np.int32).reshape(F, Cn, 4), law, gain.view(np.int16), out_law)]
    r = vp.process_packets(*d, legs=legs, want=("mix", "enc"))
    torch.cuda.synchronize()
    assert np.array_equal(r["mix"].cpu().numpy(), want[0]) and np.array_equal(r["enc"].cpu().numpy(), want[1])
