"""GPU parity of the batched RX front-end (igd_ed137_parse + igd_rx_track) and of the gate
arbitration (igd_gate_arbitrate) against the oracle's restatement of transport_rtp_cb
(TransportAdapter.cpp:240-316), the R2S watchdog (roip_ed137.cpp:1767-1780) and
checkEvents' gate decisions (roip_ed137.cpp:5627-5719, 5985-6121, 6124-6231)."""
import numpy as np
import pytest
import torch

import oracle_py as O
import rx_arb_cases as R
import igate4xsoftphonedsp_b200 as ig
from igate4xsoftphonedsp_b200 import _native as N
from igate4xsoftphonedsp_b200 import synth

pytestmark = pytest.mark.gpu


def rx_front_end(vp, pkts, sizes, present, state, **kw):
    F, Cn, stride = pkts.shape
    fields, _ = vp.ed137_parse(pkts.reshape(F * Cn, stride), sizes.reshape(-1), want_payload=False)
    return vp.rx_track(fields.reshape(F, Cn), state, present, **kw)


@pytest.mark.parametrize("F,Cn,wd", [(400, 70, 2), (64, 1, 1), (1, 33, 2), (257, 130, 0)])
def test_rx_front_end_matches_reference_walk(vp, F, Cn, wd):
    pkts, sizes, present = R.make_rx_stream(F, Cn, seed=F + Cn)
    want_ev, want_st = R.oracle_rx_walk(pkts, sizes, present, now0=5000, wd_ticks=wd)
    st = np.zeros(Cn, N.RX_STATE_DT)
    ev = rx_front_end(vp, pkts, sizes, present, st, now_ms0=5000, wd_ticks=wd)
    assert ev.tobytes() == want_ev.tobytes()
    assert st.tobytes() == want_st.tobytes()


def test_rx_state_carries_across_batches(vp):
    F, Cn = 300, 40
    pkts, sizes, present = R.make_rx_stream(F, Cn, seed=11)
    want_ev, want_st = R.oracle_rx_walk(pkts, sizes, present, now0=0, wd_ticks=2)
    st = np.zeros(Cn, N.RX_STATE_DT)
    parts = []
    for a, b in [(0, 101), (101, 102), (102, 300)]:       # odd split: the watchdog phase must carry (frame0)
        parts.append(rx_front_end(vp, pkts[a:b], sizes[a:b], present[a:b], st, now_ms0=20 * a, frame0=a))
    assert np.concatenate(parts).tobytes() == want_ev.tobytes() and st.tobytes() == want_st.tobytes()


def test_rx_track_device_pointers(vp):
    F, Cn = 128, 96
    pkts, sizes, present = R.make_rx_stream(F, Cn, seed=5)
    want_ev, want_st = R.oracle_rx_walk(pkts, sizes, present)
    dev = torch.device("cuda", vp.device)
    fields, _ = vp.ed137_parse(torch.from_numpy(pkts.reshape(F * Cn, -1)).to(dev), torch.from_numpy(sizes.reshape(-1).astype(np.int32)).to(dev),
                               want_payload=False)
    st = torch.zeros((Cn, 4), dtype=torch.int32, device=dev)
    ev = vp.rx_track(fields.reshape(F, Cn, 4), st, torch.from_numpy(present).to(dev), now_ms0=1000)
    assert ev.cpu().numpy().tobytes() == want_ev.tobytes()
    assert st.cpu().numpy().tobytes() == want_st.tobytes()


@pytest.mark.parametrize("mode", [N.ARB_CLIENT_PTT, N.ARB_SERVER_BEST])
@pytest.mark.parametrize("G,B,F", [(4, 100, 300), (1, 5, 50), (32, 3, 120), (7, 33, 64), (4, 1, 1)])
def test_gate_arbitration_matches_reference_walk(vp, mode, G, B, F):
    w = R.make_arb_words(F, B, G, mode, seed=G * 10 + B)
    active = (np.random.default_rng(B).random(B * G) < 0.9).astype(np.uint8)
    want, wl, wb = R.oracle_arb_walk(w, G, mode, active)
    legs, br = np.zeros(B * G, N.ARB_LEG_DT), np.zeros(B, N.ARB_BRIDGE_DT)
    got = vp.gate_arbitrate(w, legs, br, G, mode, active)
    assert np.array_equal(got, want) and legs.tobytes() == wl.tobytes() and br.tobytes() == wb.tobytes()
    # split in two calls: the state carries
    legs2, br2 = np.zeros(B * G, N.ARB_LEG_DT), np.zeros(B, N.ARB_BRIDGE_DT)
    h = F // 3
    got2 = np.concatenate([vp.gate_arbitrate(w[:h], legs2, br2, G, mode, active),
                           vp.gate_arbitrate(w[h:], legs2, br2, G, mode, active)]) if h else got
    assert np.array_equal(got2, want)


def test_rx_events_drive_the_gates_that_drive_the_mix(vp):
    """the whole receive side on the GPU: packets -> parse -> liveness/latch -> PTT arbitration ->
    gains -> fused decode/meter/mix/encode, against the same chain through the oracle."""
    F, B, G = 120, 16, 4
    Cn = B * G
    rng = np.random.default_rng(42)
    law = synth.laws(Cn)
    codes = rng.integers(0, 256, (F, Cn, 160), dtype=np.uint8)
    words = R.make_arb_words(F, B, G, N.ARB_CLIENT_PTT, seed=9)
    pkts = np.zeros((F, Cn, 180), np.uint8)
    hdr = np.zeros(20, np.uint8)
    L = O.lib()
    for f in range(F):
        for c in range(Cn):
            L.orc_hdr_write(hdr.ctypes.data, 2, 0, 1, 0, 0, 8 if law[c] == 0 else 0, f, 160 * f, c, 0x0167, 1, int(words[f, c]))
            pkts[f, c, :20] = hdr
    pkts[:, :, 20:] = codes
    sizes = np.full((F, Cn), 180, np.uint32)
    present = np.ones((F, Cn), np.uint8)
    # GPU chain
    fields, payload = vp.ed137_parse(pkts.reshape(F * Cn, 180), sizes.reshape(-1))
    st = np.zeros(Cn, N.RX_STATE_DT)
    ev = vp.rx_track(fields.reshape(F, Cn), st, present)
    legs, br = np.zeros(Cn, N.ARB_LEG_DT), np.zeros(B, N.ARB_BRIDGE_DT)
    gain = vp.gate_arbitrate(ev, legs, br, G, N.ARB_CLIENT_PTT)      # reads igd_rx_event.word in place
    out_law = synth.out_laws(B)
    got = vp.process_batch(payload.reshape(F, Cn, 160), law, gain, out_law, G)
    # oracle chain
    want_ev, _ = R.oracle_rx_walk(pkts, sizes, present, now0=0)
    want_gain, _, _ = R.oracle_arb_walk(want_ev["word"], G, N.ARB_CLIENT_PTT)
    mix, enc, meter, bmeter = O.process_batch(codes, law, want_gain, out_law, G)
    assert np.array_equal(payload.reshape(F, Cn, 160), codes)
    assert np.array_equal(gain, want_gain) and (gain == 256).any()
    assert np.array_equal(got["mix"], mix) and np.array_equal(got["enc"], enc)
    assert got["bmeter"].tobytes() == bmeter.tobytes()


def test_tick_chain_is_cuda_graph_capturable(vp):
    """real-time shape: one tick per call (F = 1) on device buffers; the whole chain (4 kernels) is captured
    in a CUDA graph through the library's stream hook and replayed on new packets -- same bytes as eager."""
    B, G = 96, 4
    Cn = B * G
    dev = torch.device("cuda", vp.device)
    rng = np.random.default_rng(12)
    law = torch.from_numpy(synth.laws(Cn)).to(dev)
    out_law = torch.from_numpy(synth.out_laws(B)).to(dev)

    def packets(seed):
        pk = rng.integers(0, 256, (Cn, 180), dtype=np.uint8)
        hdr = np.zeros(20, np.uint8)
        for c in range(Cn):
            word = (int(rng.integers(0, 5)) << 29) | (int(rng.integers(0, 64)) << 22)
            O.lib().orc_hdr_write(hdr.ctypes.data, 2, 0, 1, 0, 0, 8, seed, 160 * seed, c, 0x0167, 1, word)
            pk[c, :20] = hdr
        return pk

    pk_dev = torch.zeros((Cn, 180), dtype=torch.uint8, device=dev)
    state = {k: torch.zeros(s, dtype=torch.int32, device=dev) for k, s in (("rx", (Cn, 4)), ("legs", (Cn, 2)), ("br", (B, 4)))}
    out = vp.alloc_outputs(1, B, G)
    keep = {}

    def tick():
        fields, payload = vp.ed137_parse(pk_dev)
        ev = vp.rx_track(fields.reshape(1, Cn, 4), state["rx"])
        keep["gain"] = vp.gate_arbitrate(ev, state["legs"], state["br"], G, N.ARB_CLIENT_PTT)
        vp.process_batch(payload.reshape(1, Cn, 160), law, keep["gain"], out_law, G, out=out)

    ticks = [packets(k) for k in range(6)]
    # eager reference run
    eager = []
    for pk in ticks:
        pk_dev.copy_(torch.from_numpy(pk))
        tick()
        torch.cuda.synchronize()
        eager.append((out["mix"].cpu().numpy().copy(), out["enc"].cpu().numpy().copy(), keep["gain"].cpu().numpy().copy()))
    for v in state.values():
        v.zero_()
    # captured once, replayed per tick
    pk_dev.copy_(torch.from_numpy(ticks[0]))
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        tick()                                   # warm-up on the side stream (sets the function attributes)
    torch.cuda.synchronize()
    for v in state.values():
        v.zero_()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        tick()
    for k, pk in enumerate(ticks):
        pk_dev.copy_(torch.from_numpy(pk))
        gr.replay()
        torch.cuda.synchronize()
        assert np.array_equal(out["mix"].cpu().numpy(), eager[k][0]), k
        assert np.array_equal(out["enc"].cpu().numpy(), eager[k][1]), k
        assert np.array_equal(keep["gain"].cpu().numpy(), eager[k][2]), k
