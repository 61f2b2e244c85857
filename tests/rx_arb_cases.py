"""Seeded RX-stream and gate-arbitration cases plus their oracle walks (test-only).

RX: packets are fed one by one through the oracle's transport_rtp_cb restatement
(oracle/igd_oracle.c, TransportAdapter.cpp:240-316) and the R2S watchdog
(roip_ed137.cpp:1767-1780); the GPU path is igd_ed137_parse + igd_rx_track.
Arbitration: the oracle restates roip_ed137.cpp:6124-6231 (CLIENT) and
:5627-5719 + :5985-6121 (SERVER best signal) tick by tick.
"""
import ctypes as C

import numpy as np

import oracle_py as O
from igate4xsoftphonedsp_b200 import _native as N

STRIDE = 180


def make_rx_stream(F, Cn, seed=0):
    """packets u8 [F][C][180], sizes u32 [F][C], present u8 [F][C]"""
    rng = np.random.default_rng(seed)
    L = O.lib()
    pkts = rng.integers(0, 256, (F, Cn, STRIDE), dtype=np.uint8)
    sizes = np.full((F, Cn), 180, np.uint32)
    present = np.ones((F, Cn), np.uint8)
    hdr = np.zeros(20, np.uint8)
    for c in range(Cn):
        mode = c % 6
        f = 0
        while f < F:
            run = int(rng.integers(1, 40))
            kind = rng.choice(["audio", "ka", "gap", "odd"], p=[0.45, 0.35, 0.15, 0.05]) if mode else "audio"
            if kind == "gap":
                run = int(rng.integers(1, 90))         # long enough for the watchdog to strike six times
            word = int(rng.integers(0, 1 << 32)) if rng.random() < 0.7 else 0
            for k in range(f, min(F, f + run)):
                pt = {"audio": int(rng.choice([8, 0, 18])), "ka": 123, "gap": 8, "odd": int(rng.integers(0, 128))}[kind]
                L.orc_hdr_write(hdr.ctypes.data, 2, 0, 1, 0, int(k == 0), pt, k & 0xFFFF, 160 * k, c, 0x0167, 1, word)
                pkts[k, c, :20] = hdr
                if kind == "ka":
                    sizes[k, c] = 20
                elif kind == "gap":
                    present[k, c] = 0
                elif kind == "odd":
                    sizes[k, c] = int(rng.choice([0, 7, 12, 19, 20, 21, 100, 179, 180, 1043, 1044, 4000]))
            f += run
    return pkts, sizes, present


def frame_flag(pkt, size):
    """IGD_RXE_FRAME: a forwarded packet that carries a whole G.711 audio frame (pt 0 / 8, at least 160 payload
    bytes) -- the build's own notion (include/igate_dsp.h), derived from what the reference forwards to the stream"""
    return N.RXE_FRAME if (int(pkt[1]) & 0x7F) in (0, 8) and 180 <= size < 1044 else 0


def ref_comparable_sizes(sizes):
    """The same stream restricted to sizes on which the REFERENCE's behaviour is defined, so that it can be
    replayed through the reference's own transport_rtp_cb (tests/ref_py.run_rx):
      * 276 < size < 1044: the reference copies size-20 bytes into its 256-byte payload_buff
        (TransportAdapter.cpp:286-287 tests `< 1024`), overwriting payload_bufSize and the send buffers,
        and setIncomingRTP then sizes a stack array from the overwritten length (roip_ed137.cpp:6553);
      * size < 20: the reference reads pt / ed137 / length through the header cast (:248-256) beyond the
        received bytes, i.e. whatever the socket buffer still holds.
    The oracle and the CUDA path cap the copy at 256 bytes and do not latch a word from a packet shorter than
    its header; both kinds are mapped onto the reference's own drop rule (size - 20 >= 1024)."""
    bad = ((sizes > 276) & (sizes < 1044)) | (sizes < 20)
    return np.where(bad, 1044 + (sizes % 7), sizes).astype(np.uint32)


def oracle_rx_walk(pkts, sizes, present, now0=1000, tick=20, period=200, wd_ticks=2, frame0=0, state=None):
    """-> events (RX_EVENT_DT [F][C]), final state (RX_STATE_DT [C])"""
    L = O.lib()
    F, Cn = sizes.shape
    ev = np.zeros((F, Cn), N.RX_EVENT_DT)
    st = np.zeros(Cn, N.RX_STATE_DT) if state is None else state.copy()
    big = np.zeros(4096, np.uint8)
    for c in range(Cn):
        a = O.Adapter()
        L.orc_adapter_init(C.byref(a), 1, 0, b"TRx", 200, 0)
        a.r2sPacket = int(st["r2sPacket"][c])
        a.rtpAudio = int(st["rtpAudio"][c])
        a.ed137_value = int(np.array([st["ed137_value"][c]], ">u4").view("<u4")[0])   # stored in network order
        a.payloadsize = int(st["payloadsize"][c])
        cnt = C.c_int(int(st["r2sCount"][c]))
        for f in range(F):
            now = now0 + f * tick
            flags = 0
            if present[f, c]:
                flags |= N.RXE_PACKET
                n = int(sizes[f, c])
                big[:STRIDE] = pkts[f, c]
                calls = a.checkEvents_calls
                r = L.orc_transport_rtp_cb(C.byref(a), big.ctypes.data, n, now, 0, 0)
                if r < 0:
                    flags |= N.RXE_DROPPED
                elif r == 1:
                    flags |= N.RXE_AUDIO | frame_flag(pkts[f, c], n)
                if a.checkEvents_calls != calls:
                    flags |= N.RXE_EDGE
            if wd_ticks > 0 and (frame0 + f) % wd_ticks == wd_ticks - 1:
                w = L.orc_r2s_watchdog(now, a.r2sPacket, period, C.byref(cnt))
                flags |= (N.RXE_LATE if w & 1 else 0) | (N.RXE_HANGUP if w & 2 else 0)
            ev[f, c] = (L.orc_get_ed137_value(C.byref(a)), flags, min(cnt.value, 255), 0)
        st[c] = (a.r2sPacket, L.orc_get_ed137_value(C.byref(a)), a.payloadsize & 0xFFFF, a.rtpAudio, min(cnt.value, 255))
    return ev, st


def make_arb_words(F, B, G, mode, seed=0):
    """latched ED-137 words u32 [F][B*G] with PTT / squelch bursts that overlap between legs"""
    rng = np.random.default_rng(seed)
    w = np.zeros((F, B * G), np.uint32)
    for ch in range(B * G):
        f = 0
        while f < F:
            run = int(rng.integers(1, 30))
            if mode == N.ARB_CLIENT_PTT:
                ptt = int(rng.choice([0, 0, 1, 1, 2, 3, 4]))
                val = (ptt << 29) | (int(rng.integers(0, 64)) << 22)
            else:
                sq = int(rng.random() < 0.5)
                val = (sq << 28) | (int(rng.integers(0, 32)) << 3) | 0x00013100
            w[f:f + run, ch] = val
            f += run
    return w


def oracle_arb_walk(words, G, mode, active=None, legs=None, bridges=None):
    L = O.lib()
    F, Cn = words.shape
    B = Cn // G
    legs = np.zeros(Cn, N.ARB_LEG_DT) if legs is None else legs.copy()
    bridges = np.zeros(B, N.ARB_BRIDGE_DT) if bridges is None else bridges.copy()
    gain = np.zeros((F, Cn), np.uint16)
    fn = L.orc_arb_client_tick if mode == N.ARB_CLIENT_PTT else L.orc_arb_server_best_tick
    words = np.ascontiguousarray(words, np.uint32)
    for b in range(B):
        lg = legs[b * G:(b + 1) * G].copy()
        br = bridges[b:b + 1].copy()
        act = None if active is None else np.ascontiguousarray(active[b * G:(b + 1) * G], np.uint8)
        for f in range(F):
            wv = np.ascontiguousarray(words[f, b * G:(b + 1) * G])
            fn(br.ctypes.data, lg.ctypes.data, wv.ctypes.data, None if act is None else act.ctypes.data, G)
            gain[f, b * G:(b + 1) * G] = lg["gain_q7"]
        legs[b * G:(b + 1) * G] = lg
        bridges[b] = br[0]
    return gain, legs, bridges
