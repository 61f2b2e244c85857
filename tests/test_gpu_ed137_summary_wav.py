"""GPU parity for the ED-137 RTP header extension (pack + parse), the PTT event
summary and the WavWriter sink -- byte-exact against the oracle and against the
fixtures produced through the reference's own struct / WavWriter."""
import json
import os

import numpy as np
import pytest

import oracle_py as O
import tx_scenarios as T
import igate4xsoftphonedsp_b200 as ig
from igate4xsoftphonedsp_b200 import synth

pytestmark = pytest.mark.gpu
L = O.lib()


@pytest.mark.parametrize("s", T.SCENARIOS, ids=[s["name"] for s in T.SCENARIOS])
@pytest.mark.parametrize("quirks", [True, False])
def test_pack_scenarios(vp, s, quirks):
    want_pk, want_sz, want_bm, ads = T.run_oracle(s)
    if not quirks:
        want_pk, want_bm = T.clean_expectation(s, want_pk, want_sz)
    st = T.gpu_inputs(s)
    pk, sz, bm = vp.ed137_pack(s["rtp12"], s["payload"], st, ctl=s["ctl"], now_ms0=s["now0"],
                               tick_ms=s["tick_ms"], flags=ig.F_REF_QUIRKS if quirks else 0)
    assert np.array_equal(sz, want_sz)
    F, Cn = sz.shape
    for f in range(F):
        for c in range(Cn):
            n = int(sz[f, c])
            assert pk[f, c, :n].tobytes() == want_pk[f, c, :n].tobytes(), (s["name"], f, c)
    assert np.array_equal(bm, want_bm)
    for c, a in enumerate(ads):                        # carried sender state == reference adapter fields
        for name in ("packetCnt", "firstR2SPacket", "trxSlaveEnableChangedCount", "r2sSendtime",
                     "rxSlaveEnable", "txSlaveEnable", "pttstatus", "sqlstatus", "rtpFalse"):
            assert int(st[name][c]) == int(getattr(a, name)), (s["name"], c, name)


def test_pack_state_carries_across_batches(vp):
    s = T.SCENARIOS[0]
    want_pk, want_sz, _, _ = T.run_oracle(s)
    st = T.gpu_inputs(s)
    h = s["F"] // 2
    a = vp.ed137_pack(s["rtp12"][:h], s["payload"][:h], st, ctl=s["ctl"][:h], now_ms0=s["now0"],
                      tick_ms=s["tick_ms"])
    b = vp.ed137_pack(s["rtp12"][h:], s["payload"][h:], st, ctl=s["ctl"][h:],
                      now_ms0=s["now0"] + h * s["tick_ms"], tick_ms=s["tick_ms"])
    assert np.array_equal(np.concatenate([a[1], b[1]]), want_sz)
    hdr = np.concatenate([a[0], b[0]])[..., :20]
    sent = want_sz > 0
    assert np.array_equal(hdr[sent], want_pk[..., :20][sent])


@pytest.mark.parametrize("name", ["trx_out_gated", "fuzz64"])
def test_pack_quirk_q2_stale_payload_across_batches(vp, name):
    """one tick per call (what the host shim does): the adapters' stale send buffers travel in stale_payload"""
    s = [x for x in T.SCENARIOS if x["name"] == name][0]
    want_pk, want_sz, want_bm, _ = T.run_oracle(s)
    st = T.gpu_inputs(s)
    Cn = len(s["legs"])
    stale = np.zeros((Cn, 160), np.uint8)
    step = 7
    for f0 in range(0, s["F"], step):
        sl = slice(f0, f0 + step)
        pk, sz, bm = vp.ed137_pack(s["rtp12"][sl], s["payload"][sl], st, ctl=s["ctl"][sl],
                                   now_ms0=s["now0"] + f0 * s["tick_ms"], tick_ms=s["tick_ms"],
                                   flags=ig.F_REF_QUIRKS, stale_payload=stale)
        assert np.array_equal(sz, want_sz[sl]) and np.array_equal(bm, want_bm[sl])
        for f in range(sz.shape[0]):
            for c in range(Cn):
                n = int(sz[f, c])
                assert pk[f, c, :n].tobytes() == want_pk[f0 + f, c, :n].tobytes(), (name, f0 + f, c)


def test_parse_golden_reference_headers(vp, golden_dir):
    g = json.load(open(os.path.join(golden_dir, "ed137_ref_headers.json")))
    n = len(g["cases"])
    pk = np.zeros((n, 180), np.uint8)
    rng = np.random.default_rng(1)
    pk[:, 20:] = rng.integers(0, 256, (n, 160))
    for i, c in enumerate(g["cases"]):
        pk[i, :20] = np.frombuffer(bytes.fromhex(c["hex"]), np.uint8)
    fields, pay = vp.ed137_parse(pk)
    for i, c in enumerate(g["cases"]):
        ok = c["pt"] in (8, 0, 18, 123)
        f = fields[i]
        assert f["pt"] == c["pt"] and f["accepted"] == ok and f["keepalive"] == (c["pt"] == 123)
        assert f["word"] == (c["parsed_by_ref"][11] if ok else 0)
        assert f["length_raw"] == (c["parsed_by_ref"][10] if ok else 0)
        of = O.Fields()
        L.orc_ed137_fields_from_word(int(f["word"]), of)
        assert (f["ptt_type"], f["ptt_id"], f["squelch"], f["bss"]) == (of.ptt_type, of.ptt_id, of.squelch, of.bss)
        assert f["flags"] == of.active | (of.rrc_present << 1) | (of.main_tx_used << 2) | (of.main_rx_used << 3)
        assert f["payload_len"] == 160
    assert np.array_equal(pay, pk[:, 20:])


def test_parse_roundtrip_of_packed_packets_and_ragged_sizes(vp):
    s = T.SCENARIOS[-1]
    st = T.gpu_inputs(s)
    pk, sz, _ = vp.ed137_pack(s["rtp12"], s["payload"], st, ctl=s["ctl"], now_ms0=s["now0"], tick_ms=s["tick_ms"])
    F, Cn = sz.shape
    flat, sizes = pk.reshape(F * Cn, 180), sz.reshape(-1).copy()
    sizes[5], sizes[6], sizes[7] = 12, 19, 23            # truncated / ragged packets
    fields, pay = vp.ed137_parse(flat, sizes)
    for i in range(F * Cn):
        n = int(sizes[i])
        f = fields[i]
        if n < 20:
            assert f["flags"] & ig.EDF_DROPPED and not f["accepted"] and f["payload_len"] == 0
            continue
        assert f["payload_len"] == n - 20
        want_w = int.from_bytes(flat[i, 16:20].tobytes(), "big") if f["accepted"] else 0
        assert f["word"] == want_w and f["pt"] == (flat[i, 1] & 0x7F)
        assert pay[i, :n - 20].tobytes() == flat[i, 20:n].tobytes() and (pay[i, n - 20:] == 0).all()
    a = O.Adapter()                                      # and the oracle's receive callback agrees
    L.orc_adapter_init(a, 1, 0, b"TRx", 200, 0)
    for i in range(0, F * Cn, 37):
        n = int(sizes[i])
        r = L.orc_transport_rtp_cb(a, np.ascontiguousarray(flat[i]).ctypes.data, n, 0, 1, 0)
        if r == 1:
            assert not fields[i]["keepalive"] and not (fields[i]["flags"] & ig.EDF_DROPPED)
            if fields[i]["accepted"]:
                assert L.orc_get_ed137_value(a) == fields[i]["word"]
        elif r == 0:
            assert fields[i]["keepalive"]


def test_event_summary(vp):
    F, B, G = 130, 9, 4
    rng = np.random.default_rng(2)
    codes = rng.integers(0, 256, (F, B * G, 160), dtype=np.uint8)
    law, out_law, gain = synth.laws(B * G), synth.out_laws(B), synth.gains(F, B, G)
    gain[:, 5] = 0                                        # a channel that never opens
    res = vp.process_batch(codes, law, gain, out_law, G)
    got, db = vp.event_summary(res["meter"], gain)
    _, _, meter, _ = O.process_batch(codes, law, gain, out_law, G)
    want = O.event_summary(meter, gain)
    assert got.tobytes() == want.astype(ig.SUMMARY_DT).tobytes()
    for c in range(B * G):
        av, mx, mn, bm = O.summary_db(want[c])
        if want[c]["count"] == 0:
            assert np.isnan(db[c]["level_av_db"])
            continue
        assert abs(db[c]["level_av_db"] - av) < 1e-4 and abs(db[c]["level_max_db"] - mx) < 1e-4
        assert abs(db[c]["level_min_db"] - mn) < 1e-4 and db[c]["bm_av"] == bm


def test_wav_image_reference_bytes(vp, golden_dir):
    payload = np.fromfile(os.path.join(golden_dir, "wavwriter_payload.bin"), np.uint8)
    ref = open(os.path.join(golden_dir, "wavwriter_ref.bin"), "rb").read()
    assert vp.wav_image(payload, 8000, ig.LAW_ULAW, ref_quirks=True).tobytes() == ref
    clean = vp.wav_image(payload, 8000, ig.LAW_ALAW, ref_quirks=False).tobytes()
    assert clean[:4] == b"RIFF" and clean[20:22] == b"\x06\x00" and clean[22:24] == b"\x01\x00"
    assert clean[34:36] == b"\x08\x00" and clean[44:] == payload.tobytes()
    assert int.from_bytes(clean[40:44], "little") == payload.size and int.from_bytes(clean[4:8], "little") == 36 + payload.size


def test_keepalive_matches_reference_sendR2SStatus(vp):
    """batched sendR2SStatus (TransportAdapter.cpp:422-633) tick by tick against the oracle."""
    import keepalive_cases as K
    legs, hdr, ctl = K.make(300, 60, seed=5)
    want_pk, want_sz, want_h = K.oracle_walk(legs, hdr, ctl)
    st = K.initial_state(legs)
    h = np.ascontiguousarray(hdr.copy())
    for t in range(ctl.shape[0]):
        K.apply_setters(st, ctl[t])
        sz = vp.ed137_keepalive(h, st, 10_000 + 40 * t)
        assert np.array_equal(sz, want_sz[t]), t
        sent = sz == 20
        assert np.array_equal(h[sent], want_pk[t][sent]), t
    assert np.array_equal(h, want_h) and int(want_sz.sum()) > 0


@pytest.mark.parametrize("quirks", [True, False])
def test_wav_images_of_many_channels_from_the_batch_layout(vp, quirks):
    """recorder sink for many calls at once: image k == the single-channel igd_wav_image / the oracle's
    WavWriter restatement of that channel's frames, gathered out of codes [F][C][160]."""
    F, Cn = 37, 23
    rng = np.random.default_rng(3)
    codes = rng.integers(0, 256, (F, Cn, 160), dtype=np.uint8)
    law = rng.integers(0, 2, Cn).astype(np.uint8)
    chans = np.array([22, 0, 7, 7, 13], np.uint32)
    got = vp.wav_images(codes, chans, law, ref_quirks=quirks)
    L = O.lib()
    for k, ch in enumerate(chans):
        pay = np.ascontiguousarray(codes[:, ch].reshape(-1))
        one = vp.wav_image(pay, law=int(law[ch]), ref_quirks=quirks)
        assert got[k].tobytes() == one.tobytes(), (k, ch)
        if quirks:        # byte-for-byte the reference's WavWriter (WavWriter.cpp:63-156)
            hdr = np.zeros(44, np.uint8)
            body = np.zeros(2 * pay.size, np.uint8)
            L.orc_wav_header(hdr.ctypes.data, 8000, pay.size)
            L.orc_wav_body(pay.ctypes.data, pay.size, body.ctypes.data)
            assert got[k].tobytes() == hdr.tobytes() + body.tobytes()
    allc = vp.wav_images(codes, None, law, ref_quirks=quirks)
    assert allc.shape[0] == Cn and allc[7].tobytes() == got[2].tobytes()


@pytest.mark.parametrize("stride,device", [(180, False), (180, True), (184, False), (64, False)])
def test_header_only_parse_equals_the_full_parse(vp, stride, device):
    """payload_out = NULL takes the thread-per-packet header kernel: same field records as the tile /
    warp-per-packet kernels on ragged sizes, unknown payload types and short strides."""
    import torch
    rng = np.random.default_rng(stride)
    n = 3001
    pk = rng.integers(0, 256, (n, stride), dtype=np.uint8)
    pk[:, 1] = rng.choice(np.array([8, 0, 123, 18, 96, 0x88, 0xFB], np.uint8), n)
    sizes = rng.choice(np.array([stride, stride, 180, 20, 19, 12, 0, 100, 21, 1100, 2000], np.uint32), n)
    if device:
        dpk, dsz = torch.from_numpy(pk).to("cuda:0"), torch.from_numpy(sizes.view(np.int32)).to("cuda:0")
        full = vp.ed137_parse(dpk, dsz)[0].cpu().numpy()
        hdr = vp.ed137_parse(dpk, dsz, want_payload=False)[0].cpu().numpy()
    else:
        full = vp.ed137_parse(pk, sizes)[0]
        hdr = vp.ed137_parse(pk, sizes, want_payload=False)[0]
    assert full.tobytes() == hdr.tobytes()
    assert vp.ed137_parse(pk, None, want_payload=False)[0].tobytes() == vp.ed137_parse(pk, None)[0].tobytes()
