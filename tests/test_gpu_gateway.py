"""igd_gateway_process: received ED-137 packets of every leg in -> finished ED-137 packets of every bridge out,
four kernel launches, nothing but packets and state across the API.  Checked (1) against the composition of the
separately verified entry points (parse -> rx_track -> gate_arbitrate(silence) -> process_packets -> ed137_pack)
and (2) directly against the oracle / the reference: the receive walk (transport_rtp_cb), checkEvents, the
decode -> mix -> encode arithmetic and the sender (transport_send_rtp, clean mode = no Q2/Q3 quirks)."""
import numpy as np
import pytest
import torch

import oracle_py as O
import rx_arb_cases as R
import tx_scenarios as T
import igate4xsoftphonedsp_b200 as ig
from igate4xsoftphonedsp_b200 import _native as N
from igate4xsoftphonedsp_b200 import synth

pytestmark = pytest.mark.gpu
G = 4


def make_case(F, B, seed, mode):
    rng = np.random.default_rng(seed)
    Cn = B * G
    pk, sizes, present = R.make_rx_stream(F, Cn, seed=seed)
    sizes = np.where(present == 1, sizes, 0).astype(np.uint32)
    sizes = np.minimum(sizes, 4000).astype(np.uint32)
    # words that make the arbitration do something
    w = R.make_arb_words(F, B, G, mode, seed=seed + 1)
    pk[..., 16] = (w >> 24) & 0xFF
    pk[..., 17] = (w >> 16) & 0xFF
    pk[..., 18] = (w >> 8) & 0xFF
    pk[..., 19] = w & 0xFF
    law = rng.integers(0, 2, Cn).astype(np.uint8)
    out_law = rng.integers(0, 2, B).astype(np.uint8)
    legs = [dict(radiocall=1, callIn=int(rng.integers(0, 2)), calltype=T.CALLTYPES[int(rng.integers(0, 4))],
                 keepalive=int(rng.choice([40, 200])), slave=None) for _ in range(B)]
    ctl = T._random_ctl(F, B, seed + 2, p_hold=0.85)
    rtp12 = synth.rtp12(F, B, [8 if i % 2 == 0 else 0 for i in range(B)])
    return dict(pk=pk, sizes=sizes, law=law, out_law=out_law, legs=legs, ctl=ctl, rtp12=rtp12, now0=1_000_000)


def tx_state_of(case):
    s = dict(legs=case["legs"], now0=case["now0"])
    return T.gpu_inputs(s)


def composed(vp, case, mode, F, B):
    Cn = B * G
    fields, payload = vp.ed137_parse(case["pk"].reshape(F * Cn, 180), case["sizes"].reshape(-1))
    fields = fields.reshape(F, Cn)
    rx_state = np.zeros(Cn, N.RX_STATE_DT)
    ev = vp.rx_track(fields, rx_state, present=(case["sizes"] != 0).astype(np.uint8), now_ms0=case["now0"])
    legs, bridges = np.zeros(Cn, N.ARB_LEG_DT), np.zeros(B, N.ARB_BRIDGE_DT)
    gain = vp.gate_arbitrate(ev, legs, bridges, G, mode=mode, silence=True)
    r = vp.process_packets(case["pk"], fields, case["law"], gain, case["out_law"])
    tx_state = tx_state_of(case)
    pkts, sz, bm = vp.ed137_pack(case["rtp12"], r["enc"], tx_state, ctl=case["ctl"], now_ms0=case["now0"], flags=0)
    return dict(ev=ev, gain=gain, fused=r, tx_pkts=pkts, tx_sizes=sz, tx_level=bm, rx_state=rx_state, legs=legs,
                bridges=bridges, tx_state=tx_state, payload=payload.reshape(F, Cn, 160))


@pytest.mark.parametrize("F,B,mode,seed", [(60, 9, N.ARB_CLIENT_PTT, 1), (33, 50, N.ARB_SERVER_BEST, 2), (7, 1, N.ARB_CLIENT_PTT, 3),
                                           (120, 130, N.ARB_CLIENT_PTT, 4),
                                           # >= 32 768 channels: the packet-fed liveness walk, and the batch walked in
                                           # L2-sized slices of ticks (state carried from slice to slice)
                                           (37, 8192, N.ARB_CLIENT_PTT, 5), (9, 8200, N.ARB_SERVER_BEST, 6),
                                           # the real-time shape of a wide batch (a few ticks, >= 32 768 channels) and a wide
                                           # call: one thread per bridge walks liveness + arbitration (k_rxarb_bridge)
                                           (5, 8300, N.ARB_CLIENT_PTT, 7), (12, 16390, N.ARB_SERVER_BEST, 8)])
@pytest.mark.parametrize("kern", [0, N.F_KERNEL_W])      # quarter-lane fused kernel (default) and its predecessor
def test_gateway_equals_composition_of_verified_calls(vp, F, B, mode, seed, kern):
    case = make_case(F, B, seed, mode)
    want = composed(vp, case, mode, F, B)
    Cn = B * G
    rx_state, legs, bridges = np.zeros(Cn, N.RX_STATE_DT), np.zeros(Cn, N.ARB_LEG_DT), np.zeros(B, N.ARB_BRIDGE_DT)
    tx_state = tx_state_of(case)
    got = vp.gateway_process(case["pk"], case["law"], case["out_law"], rx_state, legs, bridges, case["rtp12"], tx_state,
                             rx_sizes=case["sizes"], tx_ctl=case["ctl"], mode=mode, now_ms0=case["now0"], flags=kern,
                             want=("rx_events", "gain_q7", "meter", "bmeter", "mix", "enc"))
    assert got["rx_events"].tobytes() == want["ev"].tobytes()
    assert np.array_equal(got["gain_q7"], want["gain"])
    assert np.array_equal(got["mix"], want["fused"]["mix"]) and np.array_equal(got["enc"], want["fused"]["enc"])
    assert np.array_equal(got["meter"]["hi"], want["fused"]["meter"]["hi"])
    assert got["bmeter"].tobytes() == want["fused"]["bmeter"].tobytes()
    assert np.array_equal(got["tx_sizes"], want["tx_sizes"])
    assert np.array_equal(got["tx_pkts"], want["tx_pkts"])
    for a, b in ((rx_state, want["rx_state"]), (legs, want["legs"]), (bridges, want["bridges"])):
        assert a.tobytes() == b.tobytes()
    # sender state: identical except rtpFalse, the reference's stuck-audio diagnostic counter (TransportAdapter.cpp:
    # 657-673, never read), which the gateway form does not maintain: its payload is produced after the walk
    a, b = tx_state.copy(), want["tx_state"].copy()
    a["rtpFalse"] = 0
    b["rtpFalse"] = 0
    assert a.tobytes() == b.tobytes()
    # the level of an outgoing audio packet (setOutgoingRTP, clean) is the bridge record's byte-mean
    audio = (got["tx_sizes"] == 180) & ((got["tx_pkts"][..., 1] & 0x7F) != 123)
    assert np.array_equal(got["bmeter"]["bytemean_out"][audio], want["tx_level"][audio])
    if F * B > 500:
        assert (got["tx_sizes"] == 180).sum() > 0 and (got["tx_sizes"] == 20).sum() > 0


def test_gateway_against_the_oracle_end_to_end(vp):
    """small case, everything recomputed by the oracle from the raw packets"""
    F, B, mode = 48, 3, N.ARB_CLIENT_PTT
    case = make_case(F, B, 11, mode)
    Cn = B * G
    rx_state, legs, bridges = np.zeros(Cn, N.RX_STATE_DT), np.zeros(Cn, N.ARB_LEG_DT), np.zeros(B, N.ARB_BRIDGE_DT)
    tx_state = tx_state_of(case)
    got = vp.gateway_process(case["pk"], case["law"], case["out_law"], rx_state, legs, bridges, case["rtp12"], tx_state,
                             rx_sizes=case["sizes"], tx_ctl=case["ctl"], mode=mode, now_ms0=case["now0"], wd_ticks=2,
                             want=("rx_events", "gain_q7", "mix", "enc"))
    present = (case["sizes"] != 0).astype(np.uint8)
    ev, st = R.oracle_rx_walk(case["pk"], case["sizes"], present, now0=case["now0"])
    assert got["rx_events"].tobytes() == ev.tobytes() and rx_state.tobytes() == st.tobytes()
    g, lg, br = R.oracle_arb_walk(np.ascontiguousarray(ev["word"]), G, mode)
    g = np.where((ev["flags"] & N.RXE_FRAME) == 0, g | N.GAIN_NO_AUDIO, g).astype(np.uint16)
    assert np.array_equal(got["gain_q7"], g) and legs.tobytes() == lg.tobytes() and bridges.tobytes() == br.tobytes()
    payload = np.zeros((F, Cn, 160), np.uint8)
    full = case["sizes"] >= 180
    payload[full] = case["pk"][full][:, 20:180]
    mix, enc, _, _ = O.process_batch(payload, case["law"], g, case["out_law"], G)
    assert np.array_equal(got["mix"], mix) and np.array_equal(got["enc"], enc)
    s = dict(name="gw", legs=case["legs"], F=F, ctl=case["ctl"], tick_ms=20, now0=case["now0"], payload=enc, rtp12=case["rtp12"])
    pk, sz, _, _ = T.run_oracle(s)                       # the reference-exact sender (quirks included) ...
    pk, _ = T.clean_expectation(s, pk, sz)               # ... with this tick's payload in every audio packet (Q2 off)
    assert np.array_equal(got["tx_sizes"], sz)
    assert np.array_equal(got["tx_pkts"], pk)


def test_gateway_device_buffers_many_ticks(vp):
    F, B, mode = 200, 600, N.ARB_CLIENT_PTT
    case = make_case(F, B, 21, mode)
    want = composed(vp, case, mode, F, B)
    dev = "cuda:0"
    Cn = B * G
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    rx_state = t(np.zeros(Cn, N.RX_STATE_DT).view(np.int32).reshape(Cn, 4))
    legs = t(np.zeros(Cn, N.ARB_LEG_DT).view(np.int32).reshape(Cn, 2))
    bridges = t(np.zeros(B, N.ARB_BRIDGE_DT).view(np.int32).reshape(B, 4))
    tx_state = t(tx_state_of(case).view(np.int32).reshape(B, 10))
    for rep in range(2):
        rx_state.zero_(); legs.zero_(); bridges.zero_()
        tx_state.copy_(t(tx_state_of(case).view(np.int32).reshape(B, 10)))
        got = vp.gateway_process(t(case["pk"]), t(case["law"]), t(case["out_law"]), rx_state, legs, bridges, t(case["rtp12"]),
                                 tx_state, rx_sizes=t(case["sizes"].view(np.int32)), tx_ctl=t(case["ctl"].view(np.int32).reshape(F, B, 2)),
                                 mode=mode, now_ms0=case["now0"], want=("bmeter",))
        torch.cuda.synchronize()
        assert np.array_equal(got["tx_sizes"].cpu().numpy().view(np.uint32), want["tx_sizes"])
        assert np.array_equal(got["tx_pkts"].cpu().numpy(), want["tx_pkts"])


def test_gateway_call_is_cuda_graph_capturable_with_its_three_stage_pipeline(vp):
    """64 ticks = 4 chunks: liveness walk, arbitration and fused kernel run on three streams of the library, forked from
    and joined back to the caller's stream by events -- the whole call captures into one CUDA graph and replays with the
    same bytes as the eager call."""
    F, B, mode = 64, 300, N.ARB_CLIENT_PTT
    case = make_case(F, B, 31, mode)
    dev = "cuda:0"
    Cn = B * G
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    st0 = tx_state_of(case).view(np.int32).reshape(B, 10)
    rx_state = t(np.zeros((Cn, 4), np.int32)); legs = t(np.zeros((Cn, 2), np.int32)); bridges = t(np.zeros((B, 4), np.int32))
    tx_state = t(st0)
    args = (t(case["pk"]), t(case["law"]), t(case["out_law"]), rx_state, legs, bridges, t(case["rtp12"]), tx_state)
    kw = dict(rx_sizes=t(case["sizes"].view(np.int32)), tx_ctl=t(case["ctl"].view(np.int32).reshape(F, B, 2)), mode=mode,
              now_ms0=case["now0"], want=("bmeter", "meter", "enc", "gain_q7"))

    def reset():
        rx_state.zero_(); legs.zero_(); bridges.zero_(); tx_state.copy_(t(st0))

    vp.use_torch_stream()
    eager = vp.gateway_process(*args, **kw)
    torch.cuda.synchronize()
    want = {k: v.cpu().numpy().copy() for k, v in eager.items()}
    reset()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        vp.use_torch_stream()
        vp.gateway_process(*args, out=eager, **kw)            # warm-up on the side stream (scratch, function attributes)
    torch.cuda.synchronize()
    reset()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        vp.use_torch_stream()
        vp.gateway_process(*args, out=eager, **kw)
    vp.use_torch_stream()
    for rep in range(2):
        reset()
        for v in eager.values():
            v.zero_()
        gr.replay()
        torch.cuda.synchronize()
        for k in want:
            assert np.array_equal(eager[k].cpu().numpy(), want[k]), (rep, k)
