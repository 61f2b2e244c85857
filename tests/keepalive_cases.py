"""Seeded sendR2SStatus (TransportAdapter.cpp:422-633) cases and their oracle walk (test-only)."""
import ctypes as C

import numpy as np

import oracle_py as O
from igate4xsoftphonedsp_b200 import _native as N

CALLTYPES = ["TRx", "Tx", "Rx", "Rxonly", "Idle", "TRxIdle", "Coupling"]


def make(Cn, T, seed=0):
    """per channel: leg config, initial send-buffer header, and per tick the setter values"""
    rng = np.random.default_rng(seed)
    legs = [dict(radiocall=int(rng.random() < 0.95), callIn=int(rng.random() < 0.5),
                 calltype=CALLTYPES[int(rng.integers(0, len(CALLTYPES)))],
                 keepalive=int(rng.choice([40, 200, 1000])),
                 slave=None if rng.random() < 0.5 else (int(rng.integers(0, 2)), int(rng.integers(0, 2))))
            for _ in range(Cn)]
    hdr = rng.integers(0, 256, (Cn, 20), dtype=np.uint8)
    hdr[:, 1] = np.where(rng.random(Cn) < 0.5, 123, hdr[:, 1] & 0x7F) | (hdr[:, 1] & 0x80)
    ctl = np.zeros((T, Cn), dtype=N.CTL_DT)
    for c in range(Cn):
        t = 0
        while t < T:
            run = int(rng.integers(1, 20))
            ctl["pttstatus"][t:t + run, c] = rng.random() < 0.3
            ctl["sqlstatus"][t:t + run, c] = rng.random() < 0.3
            ctl["pttpriority"][t:t + run, c] = rng.integers(0, 8)
            ctl["ed137_bssi"][t:t + run, c] = rng.integers(0, 32)
            ctl["pttid"][t:t + run, c] = rng.integers(0, 64)
            ctl["callRecorder"][t:t + run, c] = rng.random() < 0.3
            t += run
    return legs, hdr, ctl


def oracle_walk(legs, hdr, ctl, now0=10_000, tick=40):
    """-> pkts u8 [T][C][20], sizes u32 [T][C], final header [C][20]"""
    L = O.lib()
    T, Cn = ctl.shape
    pk = np.zeros((T, Cn, 20), np.uint8)
    sizes = np.zeros((T, Cn), np.uint32)
    hfin = np.zeros((Cn, 20), np.uint8)
    out = np.zeros(256, np.uint8)
    for c, leg in enumerate(legs):
        a = O.Adapter()
        L.orc_adapter_init(C.byref(a), leg["radiocall"], leg["callIn"], leg["calltype"].encode(), leg["keepalive"], now0)
        if leg["slave"] is not None:
            L.orc_setTxRxSlaveEnable(C.byref(a), leg["slave"][0], leg["slave"][1])
        for i in range(20):
            a.send_pkt_buff[i] = int(hdr[c, i])
        for t in range(T):
            k = ctl[t, c]
            L.orc_setAdapterPtt(C.byref(a), int(k["pttstatus"]), int(k["pttpriority"]), int(k["callRecorder"]))
            L.orc_setAdapterQslOn(C.byref(a), int(k["sqlstatus"]), 0, int(k["ed137_bssi"]))
            L.orc_setAdapterPttId(C.byref(a), int(k["pttid"]))
            n = L.orc_sendR2SStatus(C.byref(a), now0 + t * tick, out.ctypes.data)
            sizes[t, c] = n
            pk[t, c, :n] = out[:n]
        hfin[c] = [a.send_pkt_buff[i] for i in range(20)]
    return pk, sizes, hfin


def initial_state(legs, now0=10_000):
    lib = N.load()
    st = np.zeros(len(legs), dtype=N.STATE_DT)
    for c, leg in enumerate(legs):
        one = np.zeros(1, dtype=N.STATE_DT)
        lib.igd_ed137_state_init(one.ctypes.data, leg["radiocall"], leg["callIn"], leg["calltype"].encode(),
                                 leg["keepalive"], now0)
        if leg["slave"] is not None:
            one["rxSlaveEnableChanged"], one["txSlaveEnableChanged"] = leg["slave"]
            one["trxSlaveEnableChangedCount"] = 0
        st[c] = one[0]
    return st


def apply_setters(st, k):
    """what setAdapterPtt / setAdapterQslOn / setAdapterPttId write (TransportAdapter.cpp:135-202)"""
    for f in ("pttstatus", "pttpriority", "callRecorder", "sqlstatus", "ed137_bssi", "pttid"):
        st[f] = k[f]
    st["sqlpriority"] = 0
