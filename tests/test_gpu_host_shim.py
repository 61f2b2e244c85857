"""The C++ host shim (igate4xsoftphonedsp_b200/host/igate_shim.{h,cpp}) keeps the reference's
TransportAdapter call signatures; this drives it from a C++ program the way RoIP_ED137 would and
compares every emitted packet with the oracle's transport_send_rtp."""
import os
import struct
import subprocess

import numpy as np
import pytest

import oracle_py as O
import tx_scenarios as T
import igate4xsoftphonedsp_b200 as ig
from igate4xsoftphonedsp_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "igate4xsoftphonedsp_b200")
SHIM_SRC = os.path.join(PKG, "host", "igate_shim.cpp")
DRIVER = os.path.join(ROOT, "tests", "host_cpp", "shim_driver.cpp")
BIN = os.path.join(ROOT, "tests", "host_cpp", "shim_driver")


def build_driver():
    srcs = [DRIVER, SHIM_SRC, os.path.join(PKG, "host", "igate_shim.h"), os.path.join(PKG, "host", "igate_pj_compat.h"),
            os.path.join(ROOT, "include", "igate_dsp.h"), os.path.join(PKG, "libigate_dsp.so")]
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", BIN, DRIVER, SHIM_SRC, "-L" + PKG, "-ligate_dsp",
                               "-Wl,-rpath," + PKG])
    return BIN


def test_shim_compiles_and_links_without_a_gpu():
    build_driver()
    out = subprocess.run([BIN], capture_output=True)
    assert out.returncode == 2            # usage error, i.e. it loaded libigate_dsp.so and ran


def write_scenario(path, s, flags):
    F, Cn = s["F"], len(s["legs"])
    ctl = s["ctl"] if s["ctl"] is not None else np.zeros((F, Cn), dtype=N.CTL_DT)
    with open(path, "wb") as f:
        f.write(struct.pack("<iiiiq", F, Cn, s["tick_ms"], flags, s["now0"]))
        for leg in s["legs"]:
            sl = leg["slave"] if leg["slave"] is not None else (-1, -1)
            f.write(struct.pack("<iiiii", leg["radiocall"], leg["callIn"], leg["keepalive"], sl[0], sl[1]))
            f.write(leg["calltype"].encode().ljust(64, b"\0"))
        f.write(np.ascontiguousarray(ctl).tobytes())
        f.write(np.ascontiguousarray(s["rtp12"]).tobytes())
        f.write(np.ascontiguousarray(s["payload"]).tobytes())


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["trx_out_gated", "rx_in_sql", "slave_switch", "fuzz64"])
def test_shim_matches_reference_send_path(tmp_path, name):
    s = [x for x in T.SCENARIOS if x["name"] == name][0]
    if s["ctl"] is None:
        pytest.skip("driver always calls the setters")
    build_driver()
    inp, outp = str(tmp_path / "scn.bin"), str(tmp_path / "res.bin")
    write_scenario(inp, s, ig.F_REF_QUIRKS)
    r = subprocess.run([BIN, inp, outp], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    F, Cn = s["F"], len(s["legs"])
    raw = open(outp, "rb").read()
    o = 0
    sizes = np.frombuffer(raw, "<u4", F * Cn, o).reshape(F, Cn); o += 4 * F * Cn
    pkts = np.frombuffer(raw, "u1", F * Cn * 180, o).reshape(F, Cn, 180); o += F * Cn * 180
    out_level = np.frombuffer(raw, "u1", F * Cn, o).reshape(F, Cn); o += F * Cn
    rx_word = np.frombuffer(raw, "<u4", F * Cn, o).reshape(F, Cn); o += 4 * F * Cn
    in_level = np.frombuffer(raw, "u1", F * Cn, o).reshape(F, Cn); o += F * Cn
    edge = np.frombuffer(raw, "u1", F * Cn, o).reshape(F, Cn); o += F * Cn
    hang = np.frombuffer(raw, "u1", F * Cn, o).reshape(F, Cn)
    want_pk, want_sz, want_bm, _ = T.run_oracle(s)
    assert np.array_equal(sizes, want_sz)
    for f in range(F):
        for c in range(Cn):
            n = int(sizes[f, c])
            assert pkts[f, c, :n].tobytes() == want_pk[f, c, :n].tobytes(), (f, c)
    # trx->OutgoingRTP keeps its last value between audio packets (roip_ed137.cpp:6519-6534)
    L = O.lib()
    for c in range(Cn):
        last, lastw, lastin = 0, 0, 0
        for f in range(F):
            n = int(sizes[f, c])
            if n and (want_pk[f, c, 1] & 0x7F) != 123:
                last = int(want_bm[f, c])
            assert out_level[f, c] == last
            # receive side fed with the emitted packet: word latched for PT in {8,0,18,123}
            if n and (want_pk[f, c, 1] & 0x7F) in (8, 0, 18, 123):
                lastw = int.from_bytes(want_pk[f, c, 16:20].tobytes(), "big")
            assert rx_word[f, c] == lastw
            if n == 180 and (want_pk[f, c, 1] & 0x7F) != 123:
                p = np.ascontiguousarray(want_pk[f, c, 20:180])
                lastin = L.orc_bytemean(p.ctypes.data, 160, 0)
            assert in_level[f, c] == lastin
    # receive-side events: the emitted packets go through the oracle's transport_rtp_cb, the watchdog runs
    # on every second tick (40 ms timer); edges = setIncomingED137Value calls, hang = cause-2001 requests
    import ctypes as C
    for c in range(Cn):
        a = O.Adapter()
        L.orc_adapter_init(C.byref(a), 1, 0, b"TRx", 200, 0)
        a.r2sPacket = s["now0"]
        cnt = C.c_int(0)
        big = np.zeros(4096, np.uint8)
        for f in range(F):
            now = s["now0"] + f * s["tick_ms"]
            n = int(sizes[f, c])
            fired = 0
            if n:
                big[:180] = want_pk[f, c]
                calls = a.checkEvents_calls
                L.orc_transport_rtp_cb(C.byref(a), big.ctypes.data, n, now, 0, 0)
                fired = int(a.checkEvents_calls != calls)
            assert edge[f, c] == fired, (f, c)
            h = 0
            if f & 1:
                h = (L.orc_r2s_watchdog(now, a.r2sPacket, 200, C.byref(cnt)) >> 1) & 1
            assert hang[f, c] == h, (f, c)
    # WavWriter sink: reference-exact bytes
    rec = r.stdout.strip().splitlines()[-1]
    hdr = np.zeros(44, np.uint8)
    pay = np.ascontiguousarray(s["payload"][:3, 0].reshape(-1))
    body = np.zeros(2 * pay.size, np.uint8)
    L.orc_wav_header(hdr.ctypes.data, 8000, pay.size)
    L.orc_wav_body(pay.ctypes.data, pay.size, body.ctypes.data)
    assert open(rec, "rb").read() == hdr.tobytes() + body.tobytes()
