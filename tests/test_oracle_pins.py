"""The oracle against every pin we have: SURVEY Appendix B hashes, ITU known
answers, Python's independent audioop codec, the reference's own struct
custom_rtp_hdr / WavWriter (committed golden fixtures, regenerated live when
oracle/_ref is present), and hand-computed values of the reference formulas."""
import ctypes as C
import glob
import hashlib
import json
import math
import os
import tempfile
import warnings

import numpy as np
import pytest

import oracle_py as O
from igate4xsoftphonedsp_b200 import synth

L = O.lib()

# SURVEY.md Appendix B (computed when the reference was surveyed)
SURVEY_SHA = {
    "alaw_decode": "e04788d110e58ff8c70c93b8480190d973e3b67876b6119abbaec766cc75c174",
    "ulaw_decode": "3dab54339e520bb2c924826e3b72a917a2b612e9fd12fc867500f1d983a75827",
    "alaw_encode": "dd15ee5ab9d2c23c335e8f5e1fab5e24b902c759eeeccb09c231bd96b441391f",
    "ulaw_encode": "5ee7cf5f273f842d2234121e4cb0c98d6b20a99ac29026f94e05b36955b195be",
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_g711_table_hashes(golden_dir):
    got = {"alaw_decode": sha(O.decode_table(0)), "ulaw_decode": sha(O.decode_table(1)),
           "alaw_encode": sha(O.encode_table(0)), "ulaw_encode": sha(O.encode_table(1))}
    assert got == SURVEY_SHA
    assert got == json.load(open(os.path.join(golden_dir, "g711_pins.json")))["sha256"]


def test_g711_known_answers():
    assert [L.orc_lin2alaw(v) for v in (0, -1, 32767, -32768)] == [0xD5, 0x55, 0xAA, 0x2A]
    assert [L.orc_lin2ulaw(v) for v in (0, -1, 32767, -32768)] == [0xFF, 0x7F, 0x80, 0x00]
    assert L.orc_alaw2lin(0xD5) == 8 and L.orc_ulaw2lin(0xFF) == 0
    assert O.decode_table(0).max() == 32256 and O.decode_table(1).max() == 32124


def test_g711_vs_audioop():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        audioop = pytest.importorskip("audioop")
    codes = bytes(range(256))
    assert np.array_equal(np.frombuffer(audioop.alaw2lin(codes, 2), "<i2"), O.decode_table(0))
    assert np.array_equal(np.frombuffer(audioop.ulaw2lin(codes, 2), "<i2"), O.decode_table(1))
    pcm = np.arange(0, 32768, dtype=np.int16)        # encoders agree on every non-negative input
    assert np.array_equal(np.frombuffer(audioop.lin2alaw(pcm.tobytes(), 2), "u1"), O.g711_encode(pcm, 0))
    assert np.array_equal(np.frombuffer(audioop.lin2ulaw(pcm.tobytes(), 2), "u1"), O.g711_encode(pcm, 1))


def test_g711_roundtrip_idempotent():
    c = np.arange(256, dtype=np.uint8)
    assert np.array_equal(O.g711_encode(O.g711_decode(c, 0), 0), c)
    ru = O.g711_encode(O.g711_decode(c, 1), 1)
    diff = np.nonzero(ru != c)[0]
    assert diff.tolist() == [0x7F] and ru[0x7F] == 0xFF      # u-law "-0" folds onto "+0"


def test_cfg1_tone_levels():
    """BASELINE config 1: one channel, 60 s of 1 kHz tone, A-law enc/dec + meter per 20 ms frame."""
    pcm = synth.tone_1k(60 * 8000)
    frames = pcm.reshape(-1, 160)
    assert frames.shape[0] == 3000
    s, p = C.c_uint64(), C.c_uint32()
    L.orc_frame_power(frames[0].ctypes.data, 160, C.byref(s), C.byref(p))
    exact = 10 * math.log10((2 * 16384**2 + 4 * 11585**2) / 8.0) - 20 * math.log10(32768)
    assert abs(L.orc_rms_dbfs(s.value, 160) - exact) < 1e-9
    assert abs(L.orc_rms_dbfs(s.value, 160) - (-9.03)) < 1e-3
    assert abs(L.orc_peak_dbfs(p.value) - (-6.0206)) < 1e-4
    codes = O.g711_encode(frames, 0)
    dec = O.g711_decode(codes, 0)
    assert np.array_equal(dec[0], dec[1234])                   # period 8 divides 160
    assert np.abs(dec.astype(int) - frames.astype(int)).max() <= 512   # A-law step at |x| < 16384*1.0
    assert np.array_equal(O.g711_encode(dec, 0), codes)


def test_ed137_header_vs_reference_struct(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "ed137_ref_headers.json")))
    assert g["sizeof"] == 20
    for c in g["cases"]:
        buf = np.zeros(20, np.uint8)
        L.orc_hdr_write(buf.ctypes.data, c["v"], c["p"], c["x"], c["cc"], c["m"], c["pt"], c["seq"], c["ts"],
                        c["ssrc"], c["profile"], c["length"], c["word"])
        assert buf.tobytes().hex() == c["hex"]
        f = O.Fields()
        L.orc_ed137_fields_from_word(c["parsed_by_ref"][11], C.byref(f))
        assert f.word == c["word"]
        assert f.ptt_type == c["word"] >> 29 and f.bss == (c["word"] & 0xF8) >> 3


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built (no /root/reference here)")
def test_ed137_header_vs_reference_struct_live():
    R = O.ref()
    rng = np.random.default_rng(5)
    for _ in range(500):
        m, pt = int(rng.integers(0, 2)), int(rng.integers(0, 128))
        seq, ts, ssrc, w = (int(rng.integers(0, 65536)), int(rng.integers(0, 2**32)),
                            int(rng.integers(0, 2**32)), int(rng.integers(0, 2**32)))
        a, b = np.zeros(20, np.uint8), np.zeros(20, np.uint8)
        R.ref_hdr_build(a.ctypes.data, 2, 0, 1, 0, m, pt, seq, ts, ssrc, 0x0167, 1, w)
        L.orc_hdr_write(b.ctypes.data, 2, 0, 1, 0, m, pt, seq, ts, ssrc, 0x0167, 1, w)
        assert a.tobytes() == b.tobytes()


def test_send_rtp_header_stamp_matches_reference_struct(golden_dir):
    """transport_send_rtp's header mutation as done through the reference bit-fields."""
    g = json.load(open(os.path.join(golden_dir, "ed137_ref_headers.json")))
    for s in g["stamps"]:
        a = O.Adapter()
        L.orc_adapter_init(C.byref(a), 1, 0, b"TRx", 200, 0)
        # reproduce (m, word, pt) through the oracle's sender: craft state that yields them
        before = bytes.fromhex(s["before"])
        pkt = np.frombuffer(before[:12] + bytes(160), np.uint8).copy()
        out = np.zeros(256, np.uint8)
        n = L.orc_transport_send_rtp(C.byref(a), pkt.ctypes.data, pkt.size, 0, out.ctypes.data, 0, 0)
        assert n == 20
        after = bytes.fromhex(s["after"])
        # fixed extension words and the X bit are identical whatever the state
        assert out[12:16].tobytes() == after[12:16] == bytes([0x01, 0x67, 0x00, 0x01])
        assert out[0] == after[0] == 0x90
        assert out[2:12].tobytes() == after[2:12]


def test_wavwriter_vs_reference(golden_dir):
    payload = np.fromfile(os.path.join(golden_dir, "wavwriter_payload.bin"), np.uint8)
    ref = open(os.path.join(golden_dir, "wavwriter_ref.bin"), "rb").read()
    hdr = np.zeros(44, np.uint8)
    body = np.zeros(2 * payload.size, np.uint8)
    assert L.orc_wav_header(hdr.ctypes.data, 8000, payload.size) == 44
    assert L.orc_wav_body(payload.ctypes.data, payload.size, body.ctypes.data) == 2 * payload.size
    assert hdr.tobytes() + body.tobytes() == ref


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built")
def test_wavwriter_vs_reference_live():
    R = O.ref()
    payload = np.random.default_rng(3).integers(0, 256, 1000, dtype=np.uint8)
    with tempfile.TemporaryDirectory() as d:
        assert R.ref_wavwriter_run(os.path.join(d, "x_").encode(), 16000, payload.ctypes.data, payload.size, 33) == 0
        data = open(glob.glob(os.path.join(d, "x_*.wav"))[0], "rb").read()
    hdr = np.zeros(44, np.uint8)
    body = np.zeros(2 * payload.size, np.uint8)
    L.orc_wav_header(hdr.ctypes.data, 16000, payload.size)
    L.orc_wav_body(payload.ctypes.data, payload.size, body.ctypes.data)
    assert hdr.tobytes() + body.tobytes() == data


def test_bytemean_and_percent_formulas():
    p = np.array([0xD5] * 160, np.uint8)
    assert L.orc_bytemean(p.ctypes.data, 160, 0) == 0xD5
    assert L.orc_bytemean(p.ctypes.data, 160, 1) == (int((0xD5 - 256) * 160 / 160) & 0xFF)
    q = np.arange(160, dtype=np.uint8)
    assert L.orc_bytemean(q.ctypes.data, 160, 0) == sum(range(160)) // 160
    mixed = np.array([200, 100, 7], np.uint8)          # signed: (-56+100+7)/3 = 17
    assert L.orc_bytemean(mixed.ctypes.data, 3, 1) == 17 and L.orc_bytemean(mixed.ctypes.data, 3, 0) == 102
    neg = np.array([200, 200, 1], np.uint8)            # signed: (-111)/3 = -37 -> uint8 219
    assert L.orc_bytemean(neg.ctypes.data, 3, 1) == 219
    assert [L.orc_percent(v) for v in (30000, 15000, 299, 300, 32767, -300)] == [100, 50, 0, 1, 109, -1]


def test_gain_and_mix_semantics():
    assert [L.orc_gain_adj(v) for v in (2.0, 0.0, 0.1, 0.5, 1.0)] == [256, 0, 13, 64, 128]
    assert L.orc_apply_gain(20000, 256) == 32767 and L.orc_apply_gain(-20000, 256) == -32768
    assert L.orc_apply_gain(-1, 13) == -1 and L.orc_apply_gain(1, 13) == 0      # arithmetic shift floors
    assert L.orc_apply_gain(12345, 128) == 12345


def test_event_summary_hand_computed():
    meter = np.zeros((4, 2), dtype=O.METER_DT)
    S = [[16000, 5], [160, 7], [320000000000, 9], [48000, 11]]
    bm = [[10, 1], [250, 2], [30, 3], [40, 4]]
    for f in range(4):
        for c in range(2):
            meter[f, c]["sumsq_lo"] = S[f][c] & 0xFFFFFFFF
            meter[f, c]["hi"] = (S[f][c] >> 32) | (bm[f][c] << 8)
    gain = np.array([[256, 0], [256, 0], [0, 0], [13, 0]], np.uint16)
    out = O.event_summary(meter, gain)
    r = out[0]
    assert (r["count"], r["sum_s"], r["max_s"], r["min_s"]) == (3, 16000 + 160 + 48000, 48000, 160)
    assert (r["bm_sum"], r["bm_max"], r["bm_min"]) == (300, 250, 10)
    av, mx, mn, bmav = O.summary_db(r)
    assert abs(av - 10 * math.log10((64160 / 160) / 3)) < 1e-12
    assert abs(mx - 10 * math.log10(300)) < 1e-12 and abs(mn - 10 * math.log10(1)) < 1e-12
    assert bmav == 100
    e = out[1]                                           # never open: reference init values
    assert (e["count"], e["max_s"], e["min_s"], e["bm_max"], e["bm_min"]) == (0, 0, 255 * 160, 0, 255)


def test_rx_callback_semantics():
    a = O.Adapter()
    L.orc_adapter_init(C.byref(a), 1, 0, b"TRx", 200, 0)
    hdr = np.zeros(20, np.uint8)
    L.orc_hdr_write(hdr.ctypes.data, 2, 0, 1, 0, 0, 8, 1, 160, 7, 0x0167, 1, 0x304131F8)
    pkt = np.concatenate([hdr, np.full(160, 0xD5, np.uint8)])
    assert L.orc_transport_rtp_cb(C.byref(a), pkt.ctypes.data, 180, 5, 1, 0) == 1
    assert L.orc_get_ed137_value(C.byref(a)) == 0x304131F8 and a.IncomingRTP == 0xD5 and a.r2sPacket == 5
    assert a.checkEvents_calls == 1 and a.payloadsize == 0x0100      # length stored un-swapped
    ka = hdr.copy()
    ka[1] = 123
    assert L.orc_transport_rtp_cb(C.byref(a), ka.ctypes.data, 20, 9, 1, 0) == 0
    assert a.checkEvents_calls == 2 and a.rtpAudio == 0
    bad = hdr.copy()
    bad[1] = 96                                             # PT not in {8,0,18,123}: word not latched
    bad[16:20] = 0
    L.orc_transport_rtp_cb(C.byref(a), np.concatenate([bad, pkt[20:]]).ctypes.data, 180, 11, 1, 0)
    assert L.orc_get_ed137_value(C.byref(a)) == 0x304131F8
    big = np.zeros(20 + 1024, np.uint8)
    assert L.orc_transport_rtp_cb(C.byref(a), big.ctypes.data, big.size, 13, 1, 0) == -1


def test_fused_golden_is_stable(golden_dir):
    g = np.load(os.path.join(golden_dir, "fused_cfg2.npz"))
    mix, enc, meter, bmeter = O.process_batch(g["codes"], g["law"], g["gain"], g["out_law"], 4)
    assert np.array_equal(mix, g["mix"]) and np.array_equal(enc, g["enc"])
    assert np.array_equal(meter.view(np.uint32).reshape(g["meter"].shape), g["meter"])
    assert np.array_equal(bmeter.view(np.uint32).reshape(g["bmeter"].shape), g["bmeter"])
    assert (bmeter["n_open"] == 2).all()                  # the synthetic gate pattern keeps 2 legs open
    assert (np.abs(mix.astype(int)) == 32767).any() or (mix == -32768).any()   # saturation is exercised


def test_tx_scenarios_golden_is_stable(golden_dir):
    import tx_scenarios as T
    g = {s["name"]: s for s in json.load(open(os.path.join(golden_dir, "ed137_tx_scenarios.json")))}
    for s in T.SCENARIOS:
        pk, sizes, bm, _ = T.run_oracle(s)
        assert sha(pk) == g[s["name"]]["sha256_packets"], s["name"]
        assert sizes.tolist() == g[s["name"]]["sizes"]


def test_threaded_oracle_equals_scalar():
    F, B, G = 6, 9, 4
    rng = np.random.default_rng(1)
    codes = rng.integers(0, 256, (F, B * G, 160), dtype=np.uint8)
    gain = synth.gains(F, B, G)
    a = O.process_batch(codes, synth.laws(B * G), gain, synth.out_laws(B), G)
    b = O.process_batch(codes, synth.laws(B * G), gain, synth.out_laws(B), G, threads=4)
    for x, y in zip(a, b):
        assert x.tobytes() == y.tobytes()


def test_frame_codec_tables_equal_the_routines():
    """orc_g711_decode / orc_g711_encode index tables (like pjmedia's default alaw_ulaw_table.c);
    the tables must be the per-sample routines on every input."""
    L = O.lib()
    pcm = np.arange(-32768, 32768, dtype=np.int16)
    codes = np.arange(256, dtype=np.uint8)
    for law, enc1, dec1 in ((0, L.orc_lin2alaw, L.orc_alaw2lin), (1, L.orc_lin2ulaw, L.orc_ulaw2lin)):
        assert O.g711_encode(pcm, law).tolist() == [enc1(int(v)) for v in pcm]
        assert O.g711_decode(codes, law).tolist() == [dec1(int(c)) for c in codes]


def test_rx_arb_keepalive_golden_is_stable(golden_dir):
    """the oracle's receive walk, gate arbitration, sendR2SStatus and event message on the seeded cases
    are pinned by committed hashes (tests/golden/make_golden.py -> rx_arb_keepalive.json)."""
    import ctypes as C
    import keepalive_cases as K
    import rx_arb_cases as R
    from igate4xsoftphonedsp_b200 import _native as N
    g = json.load(open(os.path.join(golden_dir, "rx_arb_keepalive.json")))
    pkts, sizes, present = R.make_rx_stream(300, 24, seed=3)
    assert sha(pkts) + sha(sizes) + sha(present) == g["rx_walk"]["inputs"]
    ev, st = R.oracle_rx_walk(pkts, sizes, present)
    assert sha(ev) == g["rx_walk"]["events"] and sha(st) == g["rx_walk"]["state"]
    for name, mode in (("client_ptt", N.ARB_CLIENT_PTT), ("server_best", N.ARB_SERVER_BEST)):
        w = R.make_arb_words(200, 9, 4, mode, seed=4)
        assert sha(w) == g[name]["inputs"]
        gain, legs, br = R.oracle_arb_walk(w, 4, mode)
        assert (sha(gain), sha(legs), sha(br)) == (g[name]["gain"], g[name]["legs"], g[name]["bridges"])
    legs, hdr, ctl = K.make(40, 120, seed=2)
    pk, sz, hfin = K.oracle_walk(legs, hdr, ctl)
    assert (sha(pk), sha(sz), sha(hfin)) == (g["keepalive"]["packets"], g["keepalive"]["sizes"], g["keepalive"]["final_headers"])
    buf = C.create_string_buffer(1024)
    n = O.lib().orc_ptt_event_json(buf, 1024, 3, b"pptTest_released", 61.25, 70.5, 12.0412, b"sip:radio1@10.0.0.5", 133, 201, 17)
    assert buf.raw[:n].decode() == g["ptt_event_json"]
