"""ED-137 sender scenarios shared by the golden generator and the parity tests.

A scenario drives C independent radio/call legs through F frames of
transport_send_rtp (TransportAdapter.cpp:635-874): per-leg construction
arguments, optional slave-enable commands before the first frame, and per
(frame, leg) setter values.  `run_oracle` plays it through the CPU oracle;
`gpu_inputs` turns it into igd_ed137_pack arguments.
"""
import ctypes as C

import numpy as np

import oracle_py as O
from igate4xsoftphonedsp_b200 import _native as N
from igate4xsoftphonedsp_b200 import synth

CALLTYPES = ["TRx", "Tx", "Rx", "Rxonly", "TRx Idle", "Coupling", "Tx Idle", ""]


def _legs(spec):
    return [dict(radiocall=1, callIn=0, calltype="TRx", keepalive=200, slave=None) | s for s in spec]


def _ctl_from_gates(F, Cn, G=4):
    ctl = synth.ed137_ctl(F, Cn // G, G, N.CTL_DT) if Cn % G == 0 else None
    return ctl


def _mk(name, legs, F, ctl, tick_ms=20, now0=1_000_000, seed=0):
    Cn = len(legs)
    rng = np.random.default_rng(1000 + seed)
    payload = rng.integers(0, 256, (F, Cn, 160), dtype=np.uint8)
    payload[::7, :, 28::10] = 0xD5          # exercises the stuck-audio detector bytes 40/50/60
    pt = [8 if i % 2 == 0 else 0 for i in range(Cn)]
    return dict(name=name, legs=_legs(legs), F=F, ctl=ctl, tick_ms=tick_ms, now0=now0,
                payload=payload, rtp12=synth.rtp12(F, Cn, pt))


def _random_ctl(F, Cn, seed, p_hold=0.9):
    rng = np.random.default_rng(seed)
    ctl = np.zeros((F, Cn), dtype=N.CTL_DT)
    cur = np.zeros(Cn, dtype=N.CTL_DT)
    for f in range(F):
        change = rng.random(Cn) > p_hold
        new = np.zeros(Cn, dtype=N.CTL_DT)
        new["pttstatus"] = rng.integers(0, 2, Cn)
        new["sqlstatus"] = rng.integers(0, 2, Cn)
        new["pttpriority"] = rng.integers(0, 8, Cn)
        new["ed137_bssi"] = rng.integers(0, 32, Cn)
        new["pttid"] = rng.integers(0, 64, Cn)
        new["callRecorder"] = rng.integers(0, 2, Cn)
        cur = np.where(change, new, cur)
        ctl[f] = cur
    return ctl


def _build():
    S = []
    F = 120
    S.append(_mk("trx_out_gated", [dict(calltype="TRx")] * 4, F, _ctl_from_gates(F, 4)))
    S.append(_mk("rx_in_sql", [dict(calltype="Rx", callIn=1)] * 4, F, _ctl_from_gates(F, 4), seed=1))
    S.append(_mk("rxonly_out", [dict(calltype="Rxonly")] * 4, F, _ctl_from_gates(F, 4), seed=2))
    S.append(_mk("tx_in_recorder", [dict(calltype="Tx", callIn=1)] * 4, F, _random_ctl(F, 4, 11), seed=3))
    S.append(_mk("idle_in", [dict(calltype="TRx Idle", callIn=1)] * 4, F, _random_ctl(F, 4, 12), seed=4))
    S.append(_mk("nonradio", [dict(radiocall=0)] * 2, 40, _random_ctl(40, 2, 13), seed=5))
    S.append(_mk("slave_switch", [dict(slave=(1, 0)), dict(slave=(0, 1)), dict(slave=(1, 1)), dict(slave=(0, 0))],
                 F, _ctl_from_gates(F, 4), seed=6))
    S.append(_mk("keepalive_fast_tick", [dict(keepalive=50), dict(keepalive=7), dict(keepalive=0),
                                         dict(keepalive=1000)], 300, np.zeros((300, 4), dtype=N.CTL_DT),
                 tick_ms=7, seed=7))
    S.append(_mk("no_ctl_idle_keepalive", [dict(calltype="TRx")] * 4, 200, None, seed=8))
    rng = np.random.default_rng(99)
    legs = [dict(calltype=CALLTYPES[int(rng.integers(0, len(CALLTYPES)))], callIn=int(rng.integers(0, 2)),
                 radiocall=int(rng.random() > 0.1), keepalive=int(rng.choice([20, 60, 200, 400])),
                 slave=None if rng.random() < 0.5 else (int(rng.integers(0, 2)), int(rng.integers(0, 2))))
            for _ in range(64)]
    S.append(_mk("fuzz64", legs, 200, _random_ctl(200, 64, 14, p_hold=0.8), seed=9))
    return S


SCENARIOS = _build()


def run_oracle(s, signed_char=0):
    """-> (pkts u8 [F][C][180] zero padded, sizes u32 [F][C], bytemean u8 [F][C], adapters)
    exactly as the reference would emit them (quirks Q2/Q3 included)."""
    L = O.lib()
    F, Cn = s["F"], len(s["legs"])
    pk = np.zeros((F, Cn, 180), np.uint8)
    sizes = np.zeros((F, Cn), np.uint32)
    bm = np.zeros((F, Cn), np.uint8)
    ads = []
    for c, leg in enumerate(s["legs"]):
        a = O.Adapter()
        L.orc_adapter_init(C.byref(a), leg["radiocall"], leg["callIn"], leg["calltype"].encode(),
                           leg["keepalive"], s["now0"])
        if leg["slave"] is not None:
            L.orc_setTxRxSlaveEnable(C.byref(a), leg["slave"][0], leg["slave"][1])
        out = np.zeros(256, np.uint8)
        for f in range(F):
            if s["ctl"] is not None:
                k = s["ctl"][f, c]
                L.orc_setAdapterPtt(C.byref(a), int(k["pttstatus"]), int(k["pttpriority"]), int(k["callRecorder"]))
                L.orc_setAdapterQslOn(C.byref(a), int(k["sqlstatus"]), 0, int(k["ed137_bssi"]))
                L.orc_setAdapterPttId(C.byref(a), int(k["pttid"]))
            pkt = np.concatenate([s["rtp12"][f, c], s["payload"][f, c]])
            n = L.orc_transport_send_rtp(C.byref(a), pkt.ctypes.data, pkt.size, s["now0"] + f * s["tick_ms"],
                                         out.ctypes.data, 1, signed_char)
            sizes[f, c] = n
            pk[f, c, :n] = out[:n]
            if n and (out[1] & 0x7F) != 123:
                bm[f, c] = a.OutgoingRTP
        ads.append(a)
    return pk, sizes, bm, ads


def clean_expectation(s, pk, sizes, signed_char=0):
    """Default (non-quirk) library behaviour derived from the oracle's output:
    same headers and sizes; payload = this frame's payload whenever one is sent
    (no stale bytes, quirk Q2 off); byte-mean over the payload bytes (Q3 off)."""
    L = O.lib()
    pk = pk.copy()
    F, Cn = sizes.shape
    bm = np.zeros((F, Cn), np.uint8)
    for f in range(F):
        for c in range(Cn):
            n = int(sizes[f, c])
            if n > 20:
                pk[f, c, 20:n] = s["payload"][f, c, :n - 20]
            if n and (pk[f, c, 1] & 0x7F) != 123:
                p = np.ascontiguousarray(s["payload"][f, c])
                bm[f, c] = L.orc_bytemean(p.ctypes.data, 160, signed_char)
    return pk, bm


def gpu_inputs(s):
    """igd_ed137_state [C] as the adapter constructor + slave command leave it."""
    Cn = len(s["legs"])
    st = np.zeros(Cn, dtype=N.STATE_DT)
    lib = N.load()
    for c, leg in enumerate(s["legs"]):
        one = np.zeros(1, dtype=N.STATE_DT)
        lib.igd_ed137_state_init(one.ctypes.data, leg["radiocall"], leg["callIn"], leg["calltype"].encode(),
                                 leg["keepalive"], s["now0"])
        if leg["slave"] is not None:       # setTxRxSlaveEnable, TransportAdapter.cpp:147-157
            one["rxSlaveEnableChanged"] = leg["slave"][0]
            one["txSlaveEnableChanged"] = leg["slave"][1]
            one["trxSlaveEnableChangedCount"] = 0
        st[c] = one[0]
    return st
