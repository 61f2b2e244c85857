"""The supervisor messages built from the GPU summaries (host formatting, no GPU needed):
Python (eventlog.py) == C++ (host/igate_eventlog.cpp) == the oracle's restatement of
createPTTEventDataLogger's QString (Functions.cpp:2169-2211)."""
import ctypes as C
import json
import os
import subprocess

import numpy as np

import oracle_py as O
from igate4xsoftphonedsp_b200 import _native as N
from igate4xsoftphonedsp_b200 import eventlog as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "igate4xsoftphonedsp_b200", "host", "igate_eventlog.cpp")
WRAP = os.path.join(ROOT, "tests", "host_cpp", "eventlog_wrap.cpp")
OUT = os.path.join(ROOT, "tests", "host_cpp", "libeventlog_test.so")


def cxx():
    newest = max(os.path.getmtime(SRC), os.path.getmtime(WRAP))
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < newest:
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", OUT, WRAP, SRC])
    L = C.CDLL(OUT)
    L.w_ptt_event.restype = C.c_size_t
    L.w_ptt_event.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_char_p, C.c_double, C.c_double, C.c_double,
                              C.c_char_p, C.c_int, C.c_int, C.c_int]
    L.w_ptt_released.restype = C.c_size_t
    L.w_ptt_released.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_char_p]
    L.w_vu.restype = C.c_size_t
    L.w_vu.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.w_vu_rec.restype = C.c_size_t
    L.w_vu_rec.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p]
    return L


def test_ptt_event_logger_message_three_ways():
    L, orc = cxx(), O.lib()
    rng = np.random.default_rng(0)
    vals = [0.0, 1.0, -1.5, 123456.789, 1e-7, 1e15, float("inf"), float("-inf"), float("nan"), 24.0824, -90.309]
    vals += list(10 * np.log10(rng.uniform(1e-3, 1e9, 50)))
    buf1, buf2 = C.create_string_buffer(1024), C.create_string_buffer(1024)
    for i, v in enumerate(vals):
        a, b, c = v, vals[(i + 3) % len(vals)], vals[(i + 7) % len(vals)]
        args = (3, b"pptTest_released", a, b, c, b"sip:radio1@10.0.0.5", i % 256, 255 - i % 256, (7 * i) % 256)
        n1 = L.w_ptt_event(buf1, 1024, *args)
        n2 = orc.orc_ptt_event_json(buf2, 1024, *args)
        py = E.ptt_event_logger_json(3, "pptTest_released", a, b, c, "sip:radio1@10.0.0.5", *args[-3:])
        assert buf1.raw[:n1] == buf2.raw[:n2] == py.encode()
    # the finite ones are valid JSON with the reference's keys (note the blank in "radioUrl ")
    m = json.loads(E.ptt_event_logger_json(1, "pptTest_pressed", 1.5, 2.5, -3.0, "u", 1, 2, 3))
    assert list(m) == ["menuID", "softPhoneID", "Ptt", "level_in_av", "level_in_max", "level_in_min", "radioUrl ",
                       "OutgoingRTPAv", "OutgoingRTPmax", "OutgoingRTPmin"]
    assert m["menuID"] == "PTTEventDataLogger" and m["level_in_min"] == -3.0


def test_released_message_from_summary_records():
    L = cxx()
    rec = np.zeros(1, N.SUMMARY_DT)
    db = np.zeros(1, N.SUMMARY_DB_DT)
    rec["bm_max"], rec["bm_min"] = 201, 17
    db["level_av_db"], db["level_max_db"], db["level_min_db"], db["bm_av"] = 61.25, 70.5, 12.0412, 133
    buf = C.create_string_buffer(1024)
    n = L.w_ptt_released(buf, 1024, 2, rec.ctypes.data, db.ctypes.data, b"sip:r")
    assert buf.raw[:n].decode() == E.ptt_released_json(2, rec[0], db[0], "sip:r")
    m = json.loads(buf.raw[:n])
    assert m["Ptt"] == "pptTest_released" and m["OutgoingRTPAv"] == 133 and m["OutgoingRTPmax"] == 201


def test_vu_meter_message():
    L = cxx()
    a = np.array([100.0, 2000.5, 0.0, 30000.0]); b = a[::-1].copy()
    adb = 10 * np.log10(np.maximum(a, 1e-9)); bdb = adb[::-1].copy()
    buf = C.create_string_buffer(2048)
    n = L.w_vu(buf, 2048, a.ctypes.data, b.ctypes.data, adb.ctypes.data, bdb.ctypes.data)
    py = E.vu_meter_json(a, b, adb, bdb)
    assert buf.raw[:n].decode() == py
    m = json.loads(py)        # the keys RoIP_ED137 reads for softphone N (roip_ed137.cpp:7688-7712)
    for k in range(1, 5):
        assert {f"in{k}", f"out{k}", f"in{k}dB", f"out{k}dB"} <= set(m)
    assert m["menuID"] == "broadcastVUMeter" and m["in2"] == 2000.5


def test_vu_meter_message_from_kernel_records():
    L = cxx()
    meter = np.zeros(4, N.METER_DT)
    bm = np.zeros(4, N.BRIDGE_DT)
    meter["hi"] = np.array([16384, 0, 32768, 8], np.uint32) << 16
    bm["mix_peak"] = [32767, 16, 0, 1000]
    buf = C.create_string_buffer(2048)
    n = L.w_vu_rec(buf, 2048, meter.ctypes.data, bm.ctypes.data)
    py = E.vu_meter_json_from_records(meter, bm)
    assert buf.raw[:n].decode() == py
    assert '"in1":16384' in py and '"in1dB":-6.0206' in py and '"in2dB":-inf' in py and '"out3dB":-inf' in py

