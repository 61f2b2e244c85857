"""(1) No-audio leg-frames are silent (IGD_GAIN_NO_AUDIO): the reference never hands a keep-alive, dropped or
truncated packet to the stream (TransportAdapter.cpp:298-315), so its bridge hears silence while the PTT
hold-off still keeps the gate open (roip_ed137.cpp:6140-6147).  All fused kernel variants vs the oracle, and the
INTEGRATION.md receive chain end to end: PTT release followed by keep-alives must give a zero mix.
(2) Event logger / VU feed from REAL fused-kernel records: fused -> igd_event_summary ->
igd_ptt_released_json == the oracle's message of the oracle's summary of the same batch (the message format
itself is pinned to the reference's createPTTEventDataLogger by tests/test_ref_pins.py)."""
import ctypes as C

import numpy as np
import pytest

import oracle_py as O
import rx_arb_cases as R
import igate4xsoftphonedsp_b200 as ig
from igate4xsoftphonedsp_b200 import _native as N
from igate4xsoftphonedsp_b200 import eventlog as E
from igate4xsoftphonedsp_b200 import synth
from test_gpu_fused import check, make

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("G,B,F", [(4, 37, 23), (1, 50, 9), (2, 33, 12), (3, 21, 10), (8, 11, 7), (32, 5, 6), (5, 9, 8)])
def test_no_audio_flag_all_kernel_variants(vp, G, B, F):
    codes, law, gain, out_law = make(F, B, G, random_codes=True, seed=G)
    rng = np.random.default_rng(100 + G)
    gain = rng.choice(np.array([0, 256, 256, 13, 128], np.uint16), gain.shape)
    flag = rng.random(gain.shape) < 0.3
    gain = np.where(flag, gain | N.GAIN_NO_AUDIO, gain).astype(np.uint16)
    want = O.process_batch(codes, law, gain, out_law, G)
    check(vp.process_batch(codes, law, gain, out_law, G), want)
    m = want[2]
    assert (m["hi"][flag] == 0).all() and (m["sumsq_lo"][flag] == 0).all() and np.isinf(m["rms_dbfs"][flag]).all()
    # every leg silent -> digital silence out
    allsil = np.full(gain.shape, 256 | N.GAIN_NO_AUDIO, np.uint16)
    got = vp.process_batch(codes, law, allsil, out_law, G)
    assert not got["mix"].any() and (got["bmeter"]["n_open"] == 0).all()


def test_no_audio_flag_generic_kernel(vp):
    codes, law, gain, out_law = make(9, 7, 6, random_codes=True, seed=3)
    gain = np.where(np.random.default_rng(1).random(gain.shape) < 0.4, gain | N.GAIN_NO_AUDIO, gain).astype(np.uint16)
    check(vp.process_batch(codes, law, gain, out_law, 6, flags=N.F_GENERIC_KERNEL), O.process_batch(codes, law, gain, out_law, 6))


def test_event_summary_skips_no_audio_frames(vp):
    codes, law, gain, out_law = make(60, 8, 4, random_codes=True, seed=9)
    gain = np.where(np.random.default_rng(2).random(gain.shape) < 0.5, gain | N.GAIN_NO_AUDIO, gain).astype(np.uint16)
    r = vp.process_batch(codes, law, gain, out_law, 4)
    rec, _ = vp.event_summary(r["meter"], gain)
    want = O.event_summary(O.process_batch(codes, law, gain, out_law, 4)[2], gain)
    assert rec.tobytes() == want.tobytes()
    assert (rec["count"] == ((gain != 0) & ((gain & N.GAIN_NO_AUDIO) == 0)).sum(axis=0)).all()


def _ptt_release_stream(F, B, G, release_at):
    """every leg: PTT (type 1) audio packets until `release_at`, then 20-byte keep-alives with PTT off"""
    L = O.lib()
    rng = np.random.default_rng(5)
    Cn = B * G
    pk = np.zeros((F, Cn, 180), np.uint8)
    sizes = np.zeros((F, Cn), np.uint32)
    hdr = np.zeros(20, np.uint8)
    for f in range(F):
        for c in range(Cn):
            talking = f < release_at and c % G == 0          # one talker per bridge
            if talking:
                L.orc_hdr_write(hdr.ctypes.data, 2, 0, 1, 0, int(f == 0), 8 if c % 2 == 0 else 0, f, 160 * f, c, 0x0167, 1,
                                (1 << 29) | 0x00013100)
                pk[f, c, :20] = hdr
                pk[f, c, 20:] = rng.integers(0, 256, 160)
                sizes[f, c] = 180
            else:
                L.orc_hdr_write(hdr.ctypes.data, 2, 0, 1, 0, 0, 123, f, 160 * f, c, 0x0167, 1, (1 << 22) | 0x00013100)
                pk[f, c, :20] = hdr
                sizes[f, c] = 20
    return pk, sizes


@pytest.mark.parametrize("form", ["packets", "two_call"])
def test_ptt_release_followed_by_keepalives_is_silent(vp, form):
    """INTEGRATION.md receive chain: parse -> rx_track -> gate_arbitrate(silence) -> fused.  The CLIENT arbitration
    holds a released PTT for five more ticks with the gate at 2.0 while the radio already sends keep-alives."""
    F, B, G, rel = 30, 6, 4, 12
    pk, sizes = _ptt_release_stream(F, B, G, rel)
    Cn = B * G
    law = (np.arange(Cn) % 2 == 1).astype(np.uint8)                  # PT 8 = A-law on even channels
    out_law = np.zeros(B, np.uint8)
    fields, payload = vp.ed137_parse(pk.reshape(F * Cn, 180), sizes.reshape(-1))
    fields = fields.reshape(F, Cn)
    state = np.zeros(Cn, N.RX_STATE_DT)
    ev = vp.rx_track(fields, state, now_ms0=1000)
    legs, bridges = np.zeros(Cn, N.ARB_LEG_DT), np.zeros(B, N.ARB_BRIDGE_DT)
    gain = vp.gate_arbitrate(ev, legs, bridges, G, mode=N.ARB_CLIENT_PTT, silence=True)
    talk = gain[:, 0::G]
    assert ((talk[:rel] & 0x7FFF) == 256).all() and ((talk[:rel] & N.GAIN_NO_AUDIO) == 0).all()
    assert ((talk[rel:rel + 5] & 0x7FFF) == 256).all(), "the hold-off keeps the gate open"
    assert ((gain[rel:] & N.GAIN_NO_AUDIO) != 0).all(), "...but keep-alive ticks carry no audio"
    if form == "packets":
        got = vp.process_packets(pk, fields, law, gain, out_law)
    else:
        got = vp.process_batch(payload.reshape(F, Cn, 160), law, gain, out_law, G)
    assert got["mix"][:rel].any()
    assert not got["mix"][rel:].any(), "keep-alives after the PTT release must not be decoded as G.711 code 0"
    assert (got["meter"]["hi"][rel:] == 0).all() and (got["bmeter"]["n_open"][rel:] == 0).all()
    # and both forms equal the oracle on the parsed payload with the same gains
    check(got, O.process_batch(payload.reshape(F, Cn, 160), law, gain, out_law, G))
    # without the silence flag the same gains would have mixed the zero-filled payload as full-scale samples
    plain = vp.gate_arbitrate(ev, np.zeros(Cn, N.ARB_LEG_DT), np.zeros(B, N.ARB_BRIDGE_DT), G, mode=N.ARB_CLIENT_PTT)
    assert np.array_equal(plain, gain & 0x7FFF)
    if form == "packets":        # ...except that the packet form knows the packets: silent either way
        assert not vp.process_packets(pk, fields, law, plain, out_law)["mix"][rel:].any()


def test_ptt_event_and_vu_messages_from_fused_records(vp):
    """SURVEY 8f-4 on the GPU: real fused-kernel records -> igd_event_summary -> the supervisor messages"""
    F, B, G = 120, 5, 4
    codes, law, gain, out_law = make(F, B, G, ch0=64)
    r = vp.process_batch(codes, law, gain, out_law, G)
    rec, db = vp.event_summary(r["meter"], gain)
    omix, oenc, ometer, obm = O.process_batch(codes, law, gain, out_law, G)
    orec = O.event_summary(ometer, gain)
    assert rec.tobytes() == orec.tobytes()
    buf = C.create_string_buffer(2048)
    for ch in range(B * G):
        av, mx, mn, bmav = O.summary_db(orec[ch])
        n = O.lib().orc_ptt_event_json(buf, 2048, 2, b"pptTest_released", av, mx, mn, b"sip:radio@10.0.0.7", bmav,
                                       int(orec["bm_max"][ch]), int(orec["bm_min"][ch]))
        want = buf.raw[:n].decode()
        got = E.ptt_released_json(2, rec[ch], db[ch], "sip:radio@10.0.0.7")
        if orec["count"][ch]:
            # same message: identical layout and integers; the three dB values are fp32 on the device
            # (1e-4 dB contract) and print with six significant digits, so compare them as numbers
            import json
            import re
            num = re.compile(r"(?<=:)-?\d+\.?\d*(?:e[+-]?\d+)?(?=[, ])")
            assert num.sub("#", got) == num.sub("#", want), ch
            g, w = json.loads(got), json.loads(want)
            for k in w:
                if isinstance(w[k], float):
                    assert abs(g[k] - w[k]) <= 1e-4 + 2e-6 * abs(w[k]), (ch, k)
                else:
                    assert g[k] == w[k], (ch, k)
    assert (orec["count"] > 0).sum() >= B * G // 2
    # VU feed of one tick from the same records: peaks in LSB and dBFS as the oracle's records give them
    f = 7
    got = E.vu_meter_json_from_records(r["meter"][f, :4], r["bmeter"][f, :4])
    want = E.vu_meter_json_from_records(ometer[f, :4], obm[f, :4])
    assert got == want and got.startswith('{"menuID":"broadcastVUMeter","in1":')


@pytest.mark.parametrize("want", [("enc", "meter", "bmeter"), ("enc",), ("mix",), ("meter",), ("bmeter", "mix")])
def test_optional_outputs_host_and_device(vp, want):
    """any of mix / enc / meter / bmeter may be NULL: what is asked for is unchanged, the rest is never written"""
    import torch
    F, B, G = 40, 130, 4                       # several items per warp
    codes, law, gain, out_law = make(F, B, G, random_codes=True, seed=11)
    full = O.process_batch(codes, law, gain, out_law, G)
    names = ("mix", "enc", "meter", "bmeter")
    got = vp.process_batch(codes, law, gain, out_law, G, want=want)           # host form (chunk pipeline)
    assert set(got) == set(want)
    for k in want:
        w = full[names.index(k)]
        if k == "meter":
            assert np.array_equal(got[k]["hi"], w["hi"]) and np.array_equal(got[k]["sumsq_lo"], w["sumsq_lo"])
        else:
            assert got[k].tobytes() == w.tobytes(), k
    dev = "cuda:0"
    d = [torch.from_numpy(a).to(dev) for a in (codes, law, gain.view(np.int16), out_law)]
    r = vp.process_batch(*d, G, want=want)                                    # device form
    torch.cuda.synchronize()
    for k in want:
        w = full[names.index(k)]
        g = r[k].cpu().numpy()
        if k == "meter":
            g = g.view(ig.METER_DT).reshape(F, B * G)
            assert np.array_equal(g["hi"], w["hi"]) and np.array_equal(g["sumsq_lo"], w["sumsq_lo"])
        elif k == "bmeter":
            assert g.view(ig.BRIDGE_DT).reshape(F, B).tobytes() == w.tobytes()
        else:
            assert g.tobytes() == w.tobytes(), k
    with pytest.raises(ig.IgdError):
        vp.process_batch(codes, law, gain, out_law, G, want=())


def test_packets_host_form_is_chunked_and_optional(vp):
    """igd_process_packets with host buffers: several 32 MiB chunks, outputs enc + meter only"""
    from test_gpu_packets_fused import make as mkp, want_of
    F, B = 60, 2048                          # 60 * 8192 * 180 B = 88 MB -> 3 chunks
    pk, sizes, law, gain, out_law = mkp(F, B, 4, True)
    fields, want = want_of(vp, pk, sizes, law, gain, out_law)
    got = vp.process_packets(pk, fields, law, gain, out_law, want=("enc", "meter"))
    assert set(got) == {"enc", "meter"}
    assert got["enc"].tobytes() == want[1].tobytes()
    assert np.array_equal(got["meter"]["hi"], want[2]["hi"]) and np.array_equal(got["meter"]["sumsq_lo"], want[2]["sumsq_lo"])
