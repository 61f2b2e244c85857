"""The walks with the tick axis across the lanes (k_rxarb_walk, k_plan_walk: calls of >= 8 ticks) and the one-thread-per-bridge
receive walk of wide, short calls (k_rxarb_bridge: from 65 536 channels, fewer than 200 ticks) against the
thread-per-channel kernels they replace (IGD_F_WALK_SERIAL: k_rx_track, k_gate_arbitrate, k_ed137_plan), which
test_gpu_rx_arb.py / test_gpu_ed137_summary_wav.py hold to the oracle.  The same code runs lane for lane on the
host against the oracle in test_walks_host.py; test_gpu_gateway.py holds the gateway call (which runs the lane
walks by default) to the composition of the separately verified entry points and to the oracle."""
import numpy as np
import pytest

import tx_scenarios as T
import igate4xsoftphonedsp_b200 as ig
from igate4xsoftphonedsp_b200 import _native as N
from test_gpu_gateway import G, make_case, tx_state_of

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("F,B,mode,seed,wd", [(8, 3, N.ARB_CLIENT_PTT, 1, 2), (100, 40, N.ARB_SERVER_BEST, 2, 2),
                                              (333, 70, N.ARB_CLIENT_PTT, 3, 3), (64, 9000, N.ARB_CLIENT_PTT, 4, 2),
                                              (45, 8200, N.ARB_SERVER_BEST, 5, 1), (200, 5, N.ARB_SERVER_BEST, 6, 0),
                                              # wide and short: one thread per bridge walks liveness + arbitration (k_rxarb_bridge)
                                              (24, 16400, N.ARB_CLIENT_PTT, 7, 2), (10, 16500, N.ARB_SERVER_BEST, 8, 3)])
def test_gateway_lane_walks_equal_the_thread_per_channel_walks(vp, F, B, mode, seed, wd):
    case = make_case(F, B, seed, mode)
    Cn = B * G
    rng = np.random.default_rng(seed)
    active = None if seed % 2 else (rng.random(Cn) < 0.8).astype(np.uint8)
    res = []
    for flags in (0, N.F_WALK_SERIAL):
        for cuts in ((0, F), (0, F // 2 + 1, F)):          # one call, and the same ticks in two calls (state carried)
            st = (np.zeros(Cn, N.RX_STATE_DT), np.zeros(Cn, N.ARB_LEG_DT), np.zeros(B, N.ARB_BRIDGE_DT), tx_state_of(case))
            parts = []
            for f0, f1 in zip(cuts[:-1], cuts[1:]):
                sl = slice(f0, f1)
                parts.append(vp.gateway_process(np.ascontiguousarray(case["pk"][sl]), case["law"], case["out_law"], st[0], st[1], st[2],
                                                np.ascontiguousarray(case["rtp12"][sl]), st[3], rx_sizes=np.ascontiguousarray(case["sizes"][sl]),
                                                tx_ctl=np.ascontiguousarray(case["ctl"][sl]), mode=mode, now_ms0=case["now0"] + 20 * f0,
                                                wd_ticks=wd, frame0=f0, active=active, flags=flags,
                                                want=("rx_events", "gain_q7", "bmeter", "enc")))
            out = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
            res.append((out, st))
    ref_out, ref_st = res[-2]                                # thread-per-channel walks, one call
    for out, st in res[:2] + res[3:]:
        for k in ref_out:
            assert out[k].tobytes() == ref_out[k].tobytes(), k
        for a, b in zip(st[:3], ref_st[:3]):
            assert a.tobytes() == b.tobytes()
        a, b = st[3].copy(), ref_st[3].copy()
        a["rtpFalse"] = 0; b["rtpFalse"] = 0
        assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize("flags", [0, N.F_REF_QUIRKS])
def test_pack_lane_walk_equals_the_thread_per_sender_walk_on_every_tx_scenario(vp, flags):
    for s in T.SCENARIOS:
        got = []
        for walk in (0, N.F_WALK_SERIAL):
            st = T.gpu_inputs(s)
            pk, sz, bm = vp.ed137_pack(s["rtp12"], s["payload"], st, ctl=s["ctl"], now_ms0=s["now0"], tick_ms=s["tick_ms"],
                                       flags=flags | walk)
            got.append((pk, sz, bm, st))
        for a, b in zip(*got):
            assert a.tobytes() == b.tobytes(), s["name"]


@pytest.mark.parametrize("mode,seed", [(N.ARB_CLIENT_PTT, 81), (N.ARB_SERVER_BEST, 82)])
def test_gateway_lane_walk_on_long_uniform_stretches_against_the_oracle(vp, mode, seed):
    """legs that keep sending one kind of packet for seconds: the walk's uniform-step path (and the general one around
    the odd packets in between), events / gains / state against the oracle walks"""
    import rx_arb_cases as R
    from test_walks_host import steady_rx_case
    F, B = 300, 40
    Cn = B * G
    pk, sizes = steady_rx_case(F, B, seed, mode)
    case = make_case(F, B, seed, mode)
    st = (np.zeros(Cn, N.RX_STATE_DT), np.zeros(Cn, N.ARB_LEG_DT), np.zeros(B, N.ARB_BRIDGE_DT), tx_state_of(case))
    got = vp.gateway_process(pk, case["law"], case["out_law"], st[0], st[1], st[2], case["rtp12"], st[3], rx_sizes=sizes,
                             tx_ctl=case["ctl"], mode=mode, now_ms0=case["now0"], wd_ticks=2, want=("rx_events", "gain_q7"))
    ev_w, st_w = R.oracle_rx_walk(pk, sizes, (sizes != 0).astype(np.uint8), now0=case["now0"], wd_ticks=2)
    g_w, lg_w, br_w = R.oracle_arb_walk(np.ascontiguousarray(ev_w["word"]), G, mode)
    g_w = np.where((ev_w["flags"] & N.RXE_FRAME) == 0, g_w | N.GAIN_NO_AUDIO, g_w).astype(np.uint16)
    assert got["rx_events"].tobytes() == ev_w.tobytes() and np.array_equal(got["gain_q7"], g_w)
    assert st[0].tobytes() == st_w.tobytes() and st[1].tobytes() == lg_w.tobytes() and st[2].tobytes() == br_w.tobytes()
