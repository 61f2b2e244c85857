"""CPU checks of the RX-liveness walk and the gate arbitration: the device math
(igd_math.cuh compiled for the host) against the oracle's restatement of the
reference, plus hand-computed behaviour the reference's code implies."""
import ctypes as C

import numpy as np

import hostbuild_py as H
import oracle_py as O
import rx_arb_cases as R
from igate4xsoftphonedsp_b200 import _native as N


def test_watchdog_hangs_up_on_the_sixth_late_tick():
    """roip_ed137.cpp:1767-1780: r2sCount counts late ticks, hang-up when it is already 5."""
    L = O.lib()
    cnt = C.c_int(0)
    out = [L.orc_r2s_watchdog(1000 + 40 * k, 0, 200, C.byref(cnt)) for k in range(8)]
    assert out == [1, 1, 1, 1, 1, 3, 1, 1] and cnt.value == 8
    assert L.orc_r2s_watchdog(500, 0, 200, C.byref(cnt)) == 0 and cnt.value == 0      # 500 <= 600: on time
    assert L.orc_r2s_watchdog(601, 0, 200, C.byref(cnt)) == 1


def test_client_ptt_priority_hand_case():
    """leg 0 presses normal PTT, leg 1 overrides with emergency, releases are held 5 ticks."""
    G = 2
    w = np.zeros((16, G), np.uint32)
    w[1:6, 0] = 1 << 29
    w[3:5, 1] = 4 << 29
    gain, legs, br = R.oracle_arb_walk(w, G, N.ARB_CLIENT_PTT)
    assert gain[0].tolist() == [0, 0]
    assert gain[1].tolist() == [256, 0] and gain[2].tolist() == [256, 0]
    assert gain[3].tolist() == [0, 256]                  # emergency wins, pressed loser muted (:6164-6174)
    # leg 1 drops to 0 at tick 5: held as "1" for five ticks (:6140-6147), muted on the sixth
    assert [int(g) for g in gain[5:11, 1]] == [256, 256, 256, 256, 256, 0]
    assert legs["on"].tolist() == [0, 0] and int(br["ptt_level"][0]) == 0


def test_server_best_signal_hand_case():
    """two radios squelch together: after 5 ticks only the better BSS is unmuted (:6026-6109)."""
    G = 4
    w = np.zeros((12, G), np.uint32)
    w[2:10, 0] = (1 << 28) | (10 << 3)
    w[2:10, 2] = (1 << 28) | (20 << 3)
    gain, legs, br = R.oracle_arb_walk(w, G, N.ARB_SERVER_BEST)
    assert not gain[:6].any()
    assert gain[6].tolist() == [0, 0, 256, 0] and gain[9].tolist() == [0, 0, 256, 0]
    assert not gain[10].any()                            # squelch closed -> MUTE (:5717-5718)
    assert int(br["sqlStatusCount"][0]) == 0 and int(br["sqlStatusOn"][0]) == 0


def _emul_rx(fields, present, now0, tick, period, wd_ticks):
    L = H.lib()
    F, Cn = fields.shape
    st = np.zeros(Cn, N.RX_STATE_DT)
    ev = np.zeros((F, Cn), N.RX_EVENT_DT)
    for c in range(Cn):
        s = st[c:c + 1].copy()
        for f in range(F):
            fl = fields[f, c:c + 1].copy()
            wd = wd_ticks > 0 and f % wd_ticks == wd_ticks - 1
            e = L.emul_rx_step(s.ctypes.data, fl.ctypes.data, int(present[f, c]), int(wd), now0 + f * tick, period)
            ev[f, c] = (s["ed137_value"][0], e, s["r2sCount"][0], 0)
        st[c] = s[0]
    return ev, st


def _host_parse(pkts, sizes):
    """the fields igd_ed137_parse produces, computed here only to drive the host build of the walk"""
    F, Cn, _ = pkts.shape
    f = np.zeros((F, Cn), N.FIELDS_DT)
    sz = sizes.astype(np.int64)
    short = sz < 20
    navail = np.minimum(sz, R.STRIDE) // 4
    pt = np.where(navail >= 1, pkts[..., 1] & 0x7F, 0)
    acc = ~short & np.isin(pt, [8, 0, 18, 123])
    word = (pkts[..., 16].astype(np.uint32) << 24) | (pkts[..., 17].astype(np.uint32) << 16) | \
           (pkts[..., 18].astype(np.uint32) << 8) | pkts[..., 19]
    f["word"] = np.where(acc, word, 0)
    f["length_raw"] = np.where(acc, pkts[..., 14].astype(np.uint16) | (pkts[..., 15].astype(np.uint16) << 8), 0)
    f["pt"] = pt
    f["accepted"] = acc
    f["keepalive"] = ~short & (pt == 123)
    dropped = short | (sz - 20 >= 1024)
    f["flags"] = np.where(dropped, N.EDF_DROPPED, 0)
    f["payload_len"] = np.where(dropped, 0, sz - 20)
    return f


def test_rx_walk_device_math_matches_oracle():
    pkts, sizes, present = R.make_rx_stream(300, 24, seed=3)
    want_ev, want_st = R.oracle_rx_walk(pkts, sizes, present)
    got_ev, got_st = _emul_rx(_host_parse(pkts, sizes), present, 1000, 20, 200, 2)
    assert np.array_equal(got_ev["flags"], want_ev["flags"])
    assert np.array_equal(got_ev["word"], want_ev["word"]) and np.array_equal(got_ev["r2sCount"], want_ev["r2sCount"])
    assert got_st.tobytes() == want_st.tobytes()
    f = want_ev["flags"]
    assert (f & N.RXE_EDGE).any() and (f & N.RXE_HANGUP).any() and (f & N.RXE_DROPPED).any()


def test_arbitration_device_math_matches_oracle():
    L = H.lib()
    for mode, G in [(N.ARB_CLIENT_PTT, 4), (N.ARB_CLIENT_PTT, 7), (N.ARB_SERVER_BEST, 4), (N.ARB_SERVER_BEST, 3)]:
        B, F = 9, 200
        w = R.make_arb_words(F, B, G, mode, seed=G)
        active = (np.random.default_rng(1).random(B * G) < 0.85).astype(np.uint8)
        want, wl, wb = R.oracle_arb_walk(w, G, mode, active)
        legs = np.zeros(B * G, N.ARB_LEG_DT)
        br = np.zeros(B, N.ARB_BRIDGE_DT)
        got = np.zeros_like(want)
        for b in range(B):
            lg = legs[b * G:(b + 1) * G].copy()
            bb = br[b:b + 1].copy()
            act = np.ascontiguousarray(active[b * G:(b + 1) * G])
            for f in range(F):
                wv = np.ascontiguousarray(w[f, b * G:(b + 1) * G])
                L.emul_arb_tick(mode, bb.ctypes.data, lg.ctypes.data, wv.ctypes.data, act.ctypes.data, G)
                got[f, b * G:(b + 1) * G] = lg["gain_q7"]
            legs[b * G:(b + 1) * G] = lg
            br[b] = bb[0]
        assert np.array_equal(got, want) and legs.tobytes() == wl.tobytes() and br.tobytes() == wb.tobytes()
        assert (want == 256).any()


def test_keepalive_device_math_matches_oracle():
    """igd_ed137_r2s_step (host build) against the oracle's sendR2SStatus restatement, header bytes included."""
    import keepalive_cases as K
    L = H.lib()
    legs, hdr, ctl = K.make(40, 120, seed=2)
    want_pk, want_sz, want_h = K.oracle_walk(legs, hdr, ctl)
    st = K.initial_state(legs)
    h = hdr.copy()
    o = (C.c_uint * 5)()
    T, Cn = ctl.shape
    sent = 0
    for t in range(T):
        K.apply_setters(st, ctl[t])
        for c in range(Cn):
            s1 = st[c:c + 1].copy()
            L.emul_r2s_step(s1.ctypes.data, 10_000 + 40 * t, int(h[c, 1] & 0x7F), o)
            st[c] = s1[0]
            word, size, pt123, marker, written = [int(x) for x in o]
            if written:       # what k_ed137_keepalive stamps into the send buffer
                h[c, 0] |= 0x10
                h[c, 1] = (h[c, 1] & 0x7F) | (0x80 if marker else 0)
                if pt123:
                    h[c, 1] = (h[c, 1] & 0x80) | 123
                h[c, 12:16] = [0x01, 0x67, 0x00, 0x01]
                h[c, 16:20] = list(int(word).to_bytes(4, "big"))
            assert size == want_sz[t, c], (t, c)
            if size:
                assert h[c].tobytes() == want_pk[t, c].tobytes(), (t, c)
                sent += 1
    assert np.array_equal(h, want_h) and sent > 100
