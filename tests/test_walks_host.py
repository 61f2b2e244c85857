"""The warp-per-call walks (csrc/igd_walks.cuh: tick axis across the lanes, forward fills by ballot, full steps only
where the input changes or the state still moves) compiled for the host and run lane for lane on 32 fibers
(tests/hostbuild/walks_emul.cpp) against the oracle -- the very code the GPU kernels k_rxarb_walk / k_plan_walk run,
checked without a GPU.  The -m gpu twins are test_gpu_gateway.py and test_gpu_walks.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import rx_arb_cases as R
import tx_scenarios as T
from igate4xsoftphonedsp_b200 import _native as N

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostbuild", "walks_emul.cpp")
CSRC = os.path.join(HERE, "..", "igate4xsoftphonedsp_b200", "csrc")
OUT = os.path.join(HERE, "hostbuild", "libwalks_emul.so")
PLAN_DT = np.dtype([("word", "<u4"), ("size", "<u2"), ("flags", "u1"), ("reserved", "u1"), ("src_frame", "<i4")])
G = 4
_lib = None


def lib():
    global _lib
    if _lib is None:
        deps = [SRC, os.path.join(CSRC, "igd_walks.cuh"), os.path.join(CSRC, "igd_math.cuh")]
        if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(p) for p in deps):
            subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-o", OUT, SRC])
        L = C.CDLL(OUT)
        vp, i = C.c_void_p, C.c_int
        L.emul_rxarb_walk.argtypes = [i, i, i, i, i, i, i, C.c_longlong, vp, vp, vp, vp, vp, vp, vp, vp]
        for fn in (L.emul_plan_walk, L.emul_plan_ref):
            fn.argtypes = [i, i, C.c_uint, C.c_uint, i, C.c_longlong, vp, vp, vp, vp, vp]
        _lib = L
    return _lib


def ptr(a):
    return None if a is None else a.ctypes.data


def rx_case(F, B, seed, mode):
    Cn = B * G
    pk, sizes, present = R.make_rx_stream(F, Cn, seed=seed)
    sizes = np.minimum(np.where(present == 1, sizes, 0), 4000).astype(np.uint32)
    w = R.make_arb_words(F, B, G, mode, seed=seed + 1)
    for k in range(4):
        pk[..., 16 + k] = (w >> (24 - 8 * k)) & 0xFF
    return pk, sizes


def run_rxarb(pk, sizes, mode, state, now0, wd_ticks=2, frame0=0, active=None, want_events=True):
    F, Cn = pk.shape[:2]
    rx, legs, br = state
    ev = np.zeros((F, Cn), N.RX_EVENT_DT) if want_events else None
    gain = np.full((F, Cn), 0x5555, np.uint16)
    rc = lib().emul_rxarb_walk(F, Cn // G, mode, 20, 200, wd_ticks, frame0, now0, ptr(pk), ptr(sizes), ptr(active), ptr(rx),
                               ptr(legs), ptr(br), ptr(ev), ptr(gain))
    assert rc == 0, "the lanes of a warp diverged around a collective"
    return ev, gain


@pytest.mark.parametrize("F,B,mode,seed,wd", [(60, 9, N.ARB_CLIENT_PTT, 1, 2), (33, 12, N.ARB_SERVER_BEST, 2, 2),
                                              (7, 1, N.ARB_CLIENT_PTT, 3, 1), (131, 6, N.ARB_CLIENT_PTT, 4, 3),
                                              (97, 7, N.ARB_SERVER_BEST, 5, 0), (64, 5, N.ARB_SERVER_BEST, 6, 5),
                                              (1, 3, N.ARB_CLIENT_PTT, 7, 2), (32, 4, N.ARB_CLIENT_PTT, 8, 2)])
def test_rxarb_walk_equals_the_oracle_walks(F, B, mode, seed, wd):
    Cn = B * G
    pk, sizes = rx_case(F, B, seed, mode)
    now0 = 1_000_000
    present = (sizes != 0).astype(np.uint8)
    rng = np.random.default_rng(seed)
    active = None if seed % 2 else (rng.random(Cn) < 0.8).astype(np.uint8)
    ev_w, st_w = R.oracle_rx_walk(pk, sizes, present, now0=now0, wd_ticks=wd)
    g_w, lg_w, br_w = R.oracle_arb_walk(np.ascontiguousarray(ev_w["word"]), G, mode, active=active)
    g_w = np.where((ev_w["flags"] & N.RXE_FRAME) == 0, g_w | N.GAIN_NO_AUDIO, g_w).astype(np.uint16)
    # one call
    st = (np.zeros(Cn, N.RX_STATE_DT), np.zeros(Cn, N.ARB_LEG_DT), np.zeros(B, N.ARB_BRIDGE_DT))
    ev, gain = run_rxarb(pk, sizes, mode, st, now0, wd_ticks=wd, active=active)
    assert ev.tobytes() == ev_w.tobytes()
    assert np.array_equal(gain, g_w)
    assert st[0].tobytes() == st_w.tobytes() and st[1].tobytes() == lg_w.tobytes() and st[2].tobytes() == br_w.tobytes()
    # the same stream in three calls: every piece of state is carried by the state arrays, the watchdog phase by frame0
    st = (np.zeros(Cn, N.RX_STATE_DT), np.zeros(Cn, N.ARB_LEG_DT), np.zeros(B, N.ARB_BRIDGE_DT))
    cuts = sorted({0, F // 3, F // 3 + min(F, 33), F})
    gains = []
    for f0, f1 in zip(cuts[:-1], cuts[1:]):
        if f1 > f0:
            _, gpart = run_rxarb(np.ascontiguousarray(pk[f0:f1]), np.ascontiguousarray(sizes[f0:f1]), mode, st, now0 + 20 * f0,
                                 wd_ticks=wd, frame0=f0, active=active, want_events=False)
            gains.append(gpart)
    assert np.array_equal(np.concatenate(gains), g_w)
    assert st[0].tobytes() == st_w.tobytes() and st[1].tobytes() == lg_w.tobytes() and st[2].tobytes() == br_w.tobytes()


def test_rxarb_walk_without_sizes_every_packet_is_a_whole_frame():
    F, B, mode = 70, 3, N.ARB_CLIENT_PTT
    pk, _ = rx_case(F, B, 40, mode)
    sizes = np.full((F, B * G), 180, np.uint32)
    ev_w, st_w = R.oracle_rx_walk(pk, sizes, np.ones((F, B * G), np.uint8), now0=5000)
    g_w, lg_w, br_w = R.oracle_arb_walk(np.ascontiguousarray(ev_w["word"]), G, mode)
    g_w = np.where((ev_w["flags"] & N.RXE_FRAME) == 0, g_w | N.GAIN_NO_AUDIO, g_w).astype(np.uint16)
    st = (np.zeros(B * G, N.RX_STATE_DT), np.zeros(B * G, N.ARB_LEG_DT), np.zeros(B, N.ARB_BRIDGE_DT))
    ev, gain = run_rxarb(pk, None, mode, st, 5000)
    assert ev.tobytes() == ev_w.tobytes() and np.array_equal(gain, g_w)
    assert st[0].tobytes() == st_w.tobytes() and st[1].tobytes() == lg_w.tobytes() and st[2].tobytes() == br_w.tobytes()


def run_plan(fn, F, Cn, flags, plen, tick, now0, ctl, payload, state):
    st = state.copy()
    plan = np.zeros((F, Cn), PLAN_DT)
    plan["src_frame"] = -77
    last = np.full(Cn, -99, np.int32)
    assert fn(F, Cn, flags, plen, tick, now0, ptr(ctl), ptr(payload), ptr(st), ptr(plan), ptr(last)) == 0
    return plan, st, last


@pytest.mark.parametrize("flags", [0, N.F_REF_QUIRKS])
def test_plan_walk_equals_transport_send_rtp_tick_by_tick_on_every_tx_scenario(flags):
    L = lib()
    for s in T.SCENARIOS:
        F, Cn = s["payload"].shape[:2]
        st0 = T.gpu_inputs(s)
        ctl = None if s["ctl"] is None else np.ascontiguousarray(s["ctl"])
        payload = np.ascontiguousarray(s["payload"])
        for pl in (payload, None):
            want = run_plan(L.emul_plan_ref, F, Cn, flags, 160, s["tick_ms"], s["now0"], ctl, pl, st0)
            got = run_plan(L.emul_plan_walk, F, Cn, flags, 160, s["tick_ms"], s["now0"], ctl, pl, st0)
            for a, b, what in zip(got, want, ("plan", "state", "last_src")):
                assert a.tobytes() == b.tobytes(), (s["name"], what)


@pytest.mark.parametrize("seed,p_hold,tick,ka", [(1, 0.97, 20, 200), (2, 0.6, 20, 40), (3, 0.995, 20, 200), (4, 0.9, 7, 30),
                                                 (5, 0.99, 20, 0), (6, 0.99, 20, -5), (7, 0.98, 0, 200), (8, 0.99, 35, 1000)])
def test_plan_walk_random_setters_throttles_and_clocks(seed, p_hold, tick, ka):
    L = lib()
    rng = np.random.default_rng(seed)
    F, Cn = 150 + 17 * seed, 24
    legs = [dict(radiocall=int(rng.random() < 0.9), callIn=int(rng.integers(0, 2)), calltype=T.CALLTYPES[int(rng.integers(0, len(T.CALLTYPES)))],
                 keepalive=ka, slave=[None, (1, 0), (1, 1), (0, 1)][int(rng.integers(0, 4))]) for _ in range(Cn)]
    st0 = T.gpu_inputs(dict(legs=legs, now0=1_000_000))
    ctl = T._random_ctl(F, Cn, seed + 100, p_hold=p_hold)
    payload = rng.integers(0, 256, (F, Cn, 160), dtype=np.uint8)
    payload[rng.random((F, Cn)) < 0.5] = 0xd5                      # the stuck-audio detector's pattern, in runs and alone
    for f0, nf in ((0, F), (0, 40), (40, F - 40)):                   # a call boundary restarts the steady-state bookkeeping
        for use_ctl in (True, False):
            c = np.ascontiguousarray(ctl[f0:f0 + nf]) if use_ctl else None
            p = np.ascontiguousarray(payload[f0:f0 + nf])
            want = run_plan(L.emul_plan_ref, nf, Cn, N.F_REF_QUIRKS, 160, tick, 1_000_000 + f0 * tick, c, p, st0)
            got = run_plan(L.emul_plan_walk, nf, Cn, N.F_REF_QUIRKS, 160, tick, 1_000_000 + f0 * tick, c, p, st0)
            for a, b, what in zip(got, want, ("plan", "state", "last_src")):
                assert a.tobytes() == b.tobytes(), (what, f0, use_ctl)
            st0 = want[1] if f0 == 0 and nf == 40 else st0


def test_rxarb_walk_watchdog_strikes_saturate_at_255():
    """a leg that stays silent for 600 watchdog ticks: the strike count is a uint8 that stops at 255 (igd_rx_step)"""
    F, B, mode = 600, 1, N.ARB_CLIENT_PTT
    pk, sizes = rx_case(F, B, 50, mode)
    sizes[20:, 1] = 0
    sizes[300:, 2] = 0
    present = (sizes != 0).astype(np.uint8)
    ev_w, st_w = R.oracle_rx_walk(pk, sizes, present, now0=777, wd_ticks=1)
    assert ev_w["r2sCount"].max() == 255
    st = (np.zeros(B * G, N.RX_STATE_DT), np.zeros(B * G, N.ARB_LEG_DT), np.zeros(B, N.ARB_BRIDGE_DT))
    ev, _ = run_rxarb(pk, sizes, mode, st, 777, wd_ticks=1)
    assert ev.tobytes() == ev_w.tobytes() and st[0].tobytes() == st_w.tobytes()


@pytest.mark.parametrize("mode,seed", [(N.ARB_CLIENT_PTT, 61), (N.ARB_SERVER_BEST, 62), (N.ARB_CLIENT_PTT, 63), (N.ARB_SERVER_BEST, 64)])
def test_rxarb_walk_from_arbitrary_start_states(mode, seed):
    """whatever the state arrays hold (counters mid-way, a uint8 lastTxmsec about to wrap, a level nobody holds): the
    walk continues exactly as the reference's per-tick code would"""
    F, B = 90, 16
    Cn = B * G
    rng = np.random.default_rng(seed)
    pk, sizes = rx_case(F, B, seed, mode)
    present = (sizes != 0).astype(np.uint8)
    rx0 = np.zeros(Cn, N.RX_STATE_DT)
    rx0["r2sPacket"] = 1_000_000 - rng.integers(0, 2000, Cn)
    rx0["ed137_value"] = rng.integers(0, 1 << 32, Cn, dtype=np.uint64).astype(np.uint32)
    rx0["payloadsize"] = rng.integers(0, 1 << 16, Cn)
    rx0["rtpAudio"] = rng.integers(0, 2, Cn)
    rx0["r2sCount"] = rng.choice([0, 1, 4, 5, 6, 200, 254, 255], Cn)
    lg0 = np.zeros(Cn, N.ARB_LEG_DT)
    lg0["last"] = rng.integers(0, 8 if mode == N.ARB_CLIENT_PTT else 2, Cn)
    lg0["msec"] = rng.choice([0, 1, 2, 4, 5, 6, 7, 200, 254, 255], Cn)
    lg0["on"] = rng.integers(0, 2, Cn)
    lg0["rssi"] = rng.integers(-1, 32, Cn)
    lg0["gain_q7"] = rng.choice([0, 256, 13], Cn)
    br0 = np.zeros(B, N.ARB_BRIDGE_DT)
    br0["ptt_level"] = rng.integers(0, 8, B)
    br0["sqlStatusCount"] = rng.choice([0, 1, 3, 4, 5, 9, 2_000_000_000], B)
    br0["sqlStatusOn"] = rng.integers(0, 2, B)
    ev_w, st_w = R.oracle_rx_walk(pk, sizes, present, now0=1_000_000, wd_ticks=2, state=rx0)
    g_w, lg_w, br_w = R.oracle_arb_walk(np.ascontiguousarray(ev_w["word"]), G, mode, legs=lg0, bridges=br0)
    g_w = np.where((ev_w["flags"] & N.RXE_FRAME) == 0, g_w | N.GAIN_NO_AUDIO, g_w).astype(np.uint16)
    st = (rx0.copy(), lg0.copy(), br0.copy())
    ev, gain = run_rxarb(pk, sizes, mode, st, 1_000_000)
    assert ev.tobytes() == ev_w.tobytes() and np.array_equal(gain, g_w)
    assert st[0].tobytes() == st_w.tobytes() and st[1].tobytes() == lg_w.tobytes() and st[2].tobytes() == br_w.tobytes()


def steady_rx_case(F, B, seed, mode):
    """long stretches of one packet kind per leg (a radio that keeps sending audio, or keep-alives), a few odd packets and
    gaps in between: most 32-tick steps take the walk's uniform-step path, the rest the general one"""
    rng = np.random.default_rng(seed)
    Cn = B * G
    L = R.O.lib()
    pk = rng.integers(0, 256, (F, Cn, 180), dtype=np.uint8)
    sizes = np.full((F, Cn), 180, np.uint32)
    hdr = np.zeros(20, np.uint8)
    w = R.make_arb_words(F, B, G, mode, seed=seed + 1)
    for c in range(Cn):
        f = 0
        while f < F:
            run = int(rng.integers(40, 160))
            ka = rng.random() < 0.4
            pt = 123 if ka else int(rng.choice([8, 0]))
            for k in range(f, min(F, f + run)):
                L.orc_hdr_write(hdr.ctypes.data, 2, 0, 1, 0, 0, pt, k & 0xFFFF, 160 * k, c, 0x0167, 1, int(w[k, c]))
                pk[k, c, :20] = hdr
                sizes[k, c] = 20 if ka else 180
            f += run
    odd = rng.random((F, Cn)) < 0.004
    sizes[odd] = rng.choice([0, 12, 100, 1044], int(odd.sum()))
    return pk, sizes


@pytest.mark.parametrize("mode,seed,wd", [(N.ARB_CLIENT_PTT, 71, 2), (N.ARB_SERVER_BEST, 72, 3), (N.ARB_CLIENT_PTT, 73, 1)])
def test_rxarb_walk_long_uniform_stretches(mode, seed, wd):
    F, B = 330, 5
    Cn = B * G
    pk, sizes = steady_rx_case(F, B, seed, mode)
    present = (sizes != 0).astype(np.uint8)
    rx0 = np.zeros(Cn, N.RX_STATE_DT)
    rx0["r2sCount"] = np.arange(Cn) % 7
    rx0["rtpAudio"] = np.arange(Cn) % 2
    rx0["r2sPacket"] = 1_000_000 - 5000
    ev_w, st_w = R.oracle_rx_walk(pk, sizes, present, now0=1_000_000, wd_ticks=wd, state=rx0)
    g_w, lg_w, br_w = R.oracle_arb_walk(np.ascontiguousarray(ev_w["word"]), G, mode)
    g_w = np.where((ev_w["flags"] & N.RXE_FRAME) == 0, g_w | N.GAIN_NO_AUDIO, g_w).astype(np.uint16)
    st = (rx0.copy(), np.zeros(Cn, N.ARB_LEG_DT), np.zeros(B, N.ARB_BRIDGE_DT))
    ev, gain = run_rxarb(pk, sizes, mode, st, 1_000_000, wd_ticks=wd)
    assert ev.tobytes() == ev_w.tobytes() and np.array_equal(gain, g_w)
    assert st[0].tobytes() == st_w.tobytes() and st[1].tobytes() == lg_w.tobytes() and st[2].tobytes() == br_w.tobytes()


@pytest.mark.parametrize("seed", range(100, 140))
def test_rxarb_walk_random_shapes_modes_and_call_splits(seed):
    """random tick counts (also fewer than one warp step and one past a step boundary), bridge counts, modes, watchdog
    periods, active masks, streams (ragged or steady) and call boundaries: always the oracle's events, gains and state"""
    rng = np.random.default_rng(seed)
    F = int(rng.choice([1, 2, 31, 32, 33, 64, 65, 97, 150]))
    B = int(rng.integers(1, 5))
    mode = int(rng.choice([N.ARB_CLIENT_PTT, N.ARB_SERVER_BEST]))
    wd = int(rng.integers(0, 5))
    Cn = B * G
    pk, sizes = (steady_rx_case if rng.random() < 0.5 else rx_case)(F, B, seed, mode)
    if rng.random() < 0.3:
        sizes = None
    active = None if rng.random() < 0.5 else (rng.random(Cn) < 0.75).astype(np.uint8)
    now0 = int(rng.integers(0, 1 << 40))
    sz = np.full((F, Cn), 180, np.uint32) if sizes is None else sizes
    ev_w, st_w = R.oracle_rx_walk(pk, sz, (sz != 0).astype(np.uint8), now0=now0, wd_ticks=wd)
    g_w, lg_w, br_w = R.oracle_arb_walk(np.ascontiguousarray(ev_w["word"]), G, mode, active=active)
    g_w = np.where((ev_w["flags"] & N.RXE_FRAME) == 0, g_w | N.GAIN_NO_AUDIO, g_w).astype(np.uint16)
    cuts = sorted({0, F} | {int(x) for x in rng.integers(0, F + 1, int(rng.integers(0, 4)))})
    st = (np.zeros(Cn, N.RX_STATE_DT), np.zeros(Cn, N.ARB_LEG_DT), np.zeros(B, N.ARB_BRIDGE_DT))
    evs, gains = [], []
    for f0, f1 in zip(cuts[:-1], cuts[1:]):
        if f1 > f0:
            e, g = run_rxarb(np.ascontiguousarray(pk[f0:f1]), None if sizes is None else np.ascontiguousarray(sizes[f0:f1]), mode, st,
                             now0 + 20 * f0, wd_ticks=wd, frame0=f0, active=active)
            evs.append(e)
            gains.append(g)
    assert np.concatenate(evs).tobytes() == ev_w.tobytes()
    assert np.array_equal(np.concatenate(gains), g_w)
    assert st[0].tobytes() == st_w.tobytes() and st[1].tobytes() == lg_w.tobytes() and st[2].tobytes() == br_w.tobytes()
