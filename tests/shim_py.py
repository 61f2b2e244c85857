"""Builds and binds tests/host_cpp/libvtable_driver.so: the PRODUCT shim (host/igate_shim.cpp, compiled with
-DIGD_HAVE_PJSIP against the stand-in pjproject headers of oracle/ref_shim) plus a driver that plays PJSIP
and touches the adapter only through `tp->op->...`, the registered RTP callback and the reference's public
functions.  The walks mirror tests/ref_py.py's so that shim output == reference output is an array compare.
Test-only."""
import ctypes as C
import os
import subprocess

import numpy as np

from igate4xsoftphonedsp_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "igate4xsoftphonedsp_b200")
SRCS = [os.path.join(ROOT, "tests", "host_cpp", "vtable_driver.cpp"), os.path.join(PKG, "host", "igate_shim.cpp"),
        os.path.join(ROOT, "oracle", "ref_shim", "igd_pj_stub_impl.cpp")]
DEPS = SRCS + [os.path.join(PKG, "host", "igate_shim.h"), os.path.join(PKG, "host", "igate_pj_compat.h"),
               os.path.join(ROOT, "oracle", "ref_shim", "igd_pj_stub.h"), os.path.join(ROOT, "include", "igate_dsp.h")]
LIB = os.path.join(ROOT, "tests", "host_cpp", "libvtable_driver.so")


def build():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(s) for s in DEPS):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-fPIC", "-shared", "-DIGD_HAVE_PJSIP",
                               "-I" + os.path.join(ROOT, "oracle", "ref_shim"), "-o", LIB] + SRCS +
                              ["-L" + PKG, "-ligate_dsp", "-Wl,-rpath," + PKG])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        N.load()                          # libigate_dsp.so first (same directory rpath also works)
        S = C.CDLL(build())
        vp = C.c_void_p
        S.shimta_open.restype = vp
        S.shimta_open.argtypes = [C.c_int, C.c_int]
        S.shimta_close.argtypes = [vp]
        S.shimta_set_clock.argtypes = [C.c_longlong]
        S.shimta_create.restype = vp
        S.shimta_create.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int,
                                    C.c_int, C.c_int]
        S.shimta_destroy.argtypes = [vp]
        S.shimta_send_rtp.argtypes = [vp, vp, C.c_size_t]
        S.shimta_flush_tx.argtypes = [vp, C.c_uint]
        S.shimta_take_sent.restype = C.c_size_t
        S.shimta_take_sent.argtypes = [vp, vp]
        S.shimta_sendR2SStatus.argtypes = [vp]
        S.shimta_bank_keepalive.argtypes = [vp]
        S.shimta_rx.argtypes = [vp, vp, C.c_size_t]
        S.shimta_rtcp.argtypes = [vp, vp, C.c_size_t]
        S.shimta_flush_rx.argtypes = [vp]
        S.shimta_watchdog.argtypes = [vp, C.c_int]
        S.shimta_take_stream.argtypes = [vp, vp, C.POINTER(C.c_size_t)]
        S.shimta_take_events.restype = C.c_uint
        S.shimta_take_events.argtypes = [vp, C.POINTER(C.c_uint32)]
        S.shimta_setAdapterPtt.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        S.shimta_setTxRxSlaveEnable.argtypes = [vp, C.c_int, C.c_int]
        S.shimta_setAdapterQslOn.argtypes = [vp, C.c_int, C.c_int, C.c_uint32]
        S.shimta_setAdapterPttId.argtypes = [vp, C.c_int]
        S.shimta_setcallRecorder.argtypes = [vp, C.c_int]
        S.shimta_setCallType.argtypes = [vp, C.c_char_p]
        S.shimta_get_ed137_value.restype = C.c_uint32
        S.shimta_get_ed137_value.argtypes = [vp]
        S.shimta_getR2SStatus.restype = C.c_longlong
        S.shimta_getR2SStatus.argtypes = [vp]
        S.shimta_levels.argtypes = [vp, C.POINTER(C.c_int)]
        S.shimta_counters.argtypes = [vp, C.POINTER(C.c_int)]
        S.shimta_vtable_passthrough.argtypes = [vp]
        S.shimta_encode_sdp.restype = C.c_size_t
        S.shimta_encode_sdp.argtypes = [vp, C.c_char_p, C.c_size_t]
        S.shimta_transport_type.argtypes = [vp]
        _lib = S
    return _lib


def create(S, leg, now, call_id=-1, attach=1):
    S.shimta_set_clock(now)
    h = S.shimta_create(leg["radiocall"], leg["callIn"], leg["calltype"].encode(), call_id, b"idx", b"TRx",
                        leg["keepalive"], 1, 0, attach)
    assert h
    return h


def run_tx(s, flags):
    """tx_scenarios through the shim's vtable: per tick every leg's setters + tp->op->send_rtp, then ONE
    igd_bank_flush_tx; packets are what reached each adapter's slave transport.
    -> pkts [F][C][180], sizes [F][C], OutgoingRTP level after every tick [F][C]"""
    S = lib()
    F, Cn = s["F"], len(s["legs"])
    bank = S.shimta_open(0, Cn)
    assert bank, "igd_bank_open failed (no sm_100 GPU?)"
    pk = np.zeros((F, Cn, 180), np.uint8)
    sizes = np.zeros((F, Cn), np.uint32)
    level = np.zeros((F, Cn), np.uint8)
    hs = []
    for c, leg in enumerate(s["legs"]):
        hs.append(create(S, leg, s["now0"], call_id=c))
        if leg["slave"] is not None:
            S.shimta_setTxRxSlaveEnable(hs[c], leg["slave"][0], leg["slave"][1])
    out = np.zeros(512, np.uint8)
    lv = (C.c_int * 7)()
    for f in range(F):
        S.shimta_set_clock(s["now0"] + f * s["tick_ms"])
        for c in range(Cn):
            if s["ctl"] is not None:
                k = s["ctl"][f, c]
                S.shimta_setAdapterPtt(hs[c], int(k["pttstatus"]), int(k["pttpriority"]), int(k["callRecorder"]))
                S.shimta_setAdapterQslOn(hs[c], int(k["sqlstatus"]), 0, int(k["ed137_bssi"]))
                S.shimta_setAdapterPttId(hs[c], int(k["pttid"]))
            pkt = np.concatenate([s["rtp12"][f, c], s["payload"][f, c]])
            assert S.shimta_send_rtp(hs[c], pkt.ctypes.data, pkt.size) == 0
        assert S.shimta_flush_tx(bank, flags) >= 0
        for c in range(Cn):
            n = S.shimta_take_sent(hs[c], out.ctypes.data)
            sizes[f, c] = n
            pk[f, c, :n] = out[:n]
            S.shimta_levels(hs[c], lv)
            level[f, c] = lv[1]
    for h in hs:
        S.shimta_destroy(h)
    S.shimta_close(bank)
    return pk, sizes, level
