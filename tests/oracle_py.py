"""ctypes binding of the CPU oracle (oracle/liboracle.so) and of the
reference-backed probe (oracle/_ref/libigd_ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, bench.py's cpu_baseline /
`--impl reference` legs and __graft_entry__.smoke() -- never by the product
package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = os.path.join(ORACLE_DIR, "liboracle.so")
_REF = os.path.join(ORACLE_DIR, "_ref", "libigd_ref.so")

FRAME = 160
LAW_ALAW, LAW_ULAW = 0, 1

METER_DT = np.dtype([("sumsq_lo", "<u4"), ("hi", "<u4"), ("rms_dbfs", "<f4"), ("peak_dbfs", "<f4")])
BRIDGE_DT = np.dtype([("bytemean_out", "u1"), ("n_open", "u1"), ("mix_peak", "<u2")])
SUMMARY_DT = np.dtype([("count", "<u4"), ("bm_sum", "<u2"), ("bm_max", "u1"), ("bm_min", "u1"),
                       ("sum_s", "<u8"), ("max_s", "<u8"), ("min_s", "<u8")])


def build(force=False):
    src = os.path.join(ORACLE_DIR, "igd_oracle.c")
    stale = (not os.path.exists(_LIB)) or os.path.getmtime(_LIB) < os.path.getmtime(src)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)


class Batch(C.Structure):
    _fields_ = [("F", C.c_int), ("B", C.c_int), ("G", C.c_int),
                ("codes", C.c_void_p), ("law", C.c_void_p), ("gain_q7", C.c_void_p),
                ("out_law", C.c_void_p), ("mix", C.c_void_p), ("enc", C.c_void_p),
                ("meter", C.c_void_p), ("bmeter", C.c_void_p), ("signed_char", C.c_int)]


class Adapter(C.Structure):
    _fields_ = [("radiostatus", C.c_int), ("pttstatus", C.c_int), ("sqlstatus", C.c_int),
                ("callIn", C.c_int), ("callRecorder", C.c_int),
                ("pttpriority", C.c_int), ("sqlpriority", C.c_int), ("ed137_bssi", C.c_int),
                ("pttid", C.c_int),
                ("rxSlaveEnable", C.c_int), ("txSlaveEnable", C.c_int),
                ("rxSlaveEnableChanged", C.c_int), ("txSlaveEnableChanged", C.c_int),
                ("trxSlaveEnableChangedCount", C.c_int),
                ("firstR2SPacket", C.c_int), ("packetCnt", C.c_int),
                ("keepAlivePeroid", C.c_int), ("rtpFalse", C.c_int), ("rtpAudio", C.c_int),
                ("r2sSendtime", C.c_longlong), ("r2sPacket", C.c_longlong),
                ("ed137_value", C.c_uint32), ("payloadsize", C.c_uint32),
                ("calltype", C.c_char * 64),
                ("send_pkt_buff", C.c_uint8 * 256), ("tmp_payload_buf", C.c_uint8 * 256),
                ("payload_buff", C.c_uint8 * 256),
                ("payload_bufSize", C.c_size_t), ("send_payload_bufSize", C.c_size_t),
                ("IncomingRTP", C.c_uint8), ("OutgoingRTP", C.c_uint8),
                ("checkEvents_calls", C.c_int)]


class Fields(C.Structure):
    _fields_ = [("word", C.c_uint32), ("ptt_type", C.c_int), ("ptt_id", C.c_int),
                ("squelch", C.c_int), ("bss", C.c_int), ("active", C.c_int),
                ("rrc_present", C.c_int), ("main_tx_used", C.c_int), ("main_rx_used", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.orc_alaw2lin.restype = C.c_int16
        L.orc_alaw2lin.argtypes = [C.c_uint8]
        L.orc_ulaw2lin.restype = C.c_int16
        L.orc_ulaw2lin.argtypes = [C.c_uint8]
        L.orc_lin2alaw.restype = C.c_uint8
        L.orc_lin2alaw.argtypes = [C.c_int]
        L.orc_lin2ulaw.restype = C.c_uint8
        L.orc_lin2ulaw.argtypes = [C.c_int]
        L.orc_g711_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        L.orc_g711_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        L.orc_bytemean.restype = C.c_uint8
        L.orc_bytemean.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_frame_power.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
        L.orc_rms_dbfs.restype = C.c_double
        L.orc_rms_dbfs.argtypes = [C.c_uint64, C.c_int]
        L.orc_peak_dbfs.restype = C.c_double
        L.orc_peak_dbfs.argtypes = [C.c_uint32]
        L.orc_percent.restype = C.c_int
        L.orc_percent.argtypes = [C.c_int]
        L.orc_gain_adj.restype = C.c_int
        L.orc_gain_adj.argtypes = [C.c_float]
        L.orc_apply_gain.restype = C.c_int16
        L.orc_apply_gain.argtypes = [C.c_int16, C.c_int]
        L.orc_process_batch.argtypes = [C.POINTER(Batch)]
        L.orc_process_batch_mt.argtypes = [C.POINTER(Batch), C.c_int]
        L.orc_event_summary.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_summary_db.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                     C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.orc_adapter_init.argtypes = [C.POINTER(Adapter), C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_longlong]
        L.orc_setAdapterPtt.argtypes = [C.POINTER(Adapter), C.c_int, C.c_int, C.c_int]
        L.orc_setTxRxSlaveEnable.argtypes = [C.POINTER(Adapter), C.c_int, C.c_int]
        L.orc_setAdapterQslOn.argtypes = [C.POINTER(Adapter), C.c_int, C.c_int, C.c_uint32]
        L.orc_setAdapterPttId.argtypes = [C.POINTER(Adapter), C.c_int]
        L.orc_setcallRecorder.argtypes = [C.POINTER(Adapter), C.c_int]
        L.orc_setCallType.argtypes = [C.POINTER(Adapter), C.c_char_p]
        L.orc_transport_send_rtp.restype = C.c_size_t
        L.orc_transport_send_rtp.argtypes = [C.POINTER(Adapter), C.c_void_p, C.c_size_t, C.c_longlong,
                                             C.c_void_p, C.c_int, C.c_int]
        L.orc_sendR2SStatus.restype = C.c_size_t
        L.orc_sendR2SStatus.argtypes = [C.POINTER(Adapter), C.c_longlong, C.c_void_p]
        L.orc_transport_rtp_cb.restype = C.c_int
        L.orc_transport_rtp_cb.argtypes = [C.POINTER(Adapter), C.c_void_p, C.c_size_t, C.c_longlong,
                                           C.c_int, C.c_int]
        L.orc_get_ed137_value.restype = C.c_uint32
        L.orc_get_ed137_value.argtypes = [C.POINTER(Adapter)]
        L.orc_ed137_fields_from_word.argtypes = [C.c_uint32, C.POINTER(Fields)]
        L.orc_hdr_write.argtypes = [C.c_void_p] + [C.c_int] * 6 + [C.c_uint16, C.c_uint32, C.c_uint32,
                                                                  C.c_uint16, C.c_uint16, C.c_uint32]
        L.orc_wav_header.restype = C.c_size_t
        L.orc_wav_header.argtypes = [C.c_void_p, C.c_int, C.c_size_t]
        L.orc_wav_body.restype = C.c_size_t
        L.orc_wav_body.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_ptt_event_json.restype = C.c_size_t
        L.orc_ptt_event_json.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_char_p, C.c_double, C.c_double, C.c_double,
                                         C.c_char_p, C.c_int, C.c_int, C.c_int]
        L.orc_r2s_watchdog.restype = C.c_int
        L.orc_r2s_watchdog.argtypes = [C.c_longlong, C.c_longlong, C.c_int, C.POINTER(C.c_int)]
        L.orc_arb_client_tick.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_arb_server_best_tick.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib = L
    return _lib


def ref_available():
    return os.path.exists(_REF)


_ref = None


def ref():
    """The reference-backed probe (reference's own ed137_rtp.h + WavWriter.cpp)."""
    global _ref
    if _ref is None:
        R = C.CDLL(_REF)
        R.ref_hdr_sizeof.restype = C.c_int
        R.ref_hdr_offsets.argtypes = [C.POINTER(C.c_int)]
        R.ref_hdr_build.argtypes = [C.c_void_p] + [C.c_int] * 6 + [C.c_uint] * 6
        R.ref_hdr_parse.argtypes = [C.c_void_p, C.POINTER(C.c_uint)]
        R.ref_hdr_stamp.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_int]
        R.ref_wavwriter_run.restype = C.c_int
        R.ref_wavwriter_run.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_uint, C.c_uint]
        _ref = R
    return _ref


# ------------------------------------------------------------------ helpers
def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def decode_table(law):
    f = lib().orc_alaw2lin if law == LAW_ALAW else lib().orc_ulaw2lin
    return np.array([f(c) for c in range(256)], dtype="<i2")


def encode_table(law):
    pcm = np.arange(-32768, 32768, dtype=np.int32).astype(np.int16)
    return g711_encode(pcm, law)


def g711_decode(codes, law):
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    out = np.empty(codes.shape, dtype=np.int16)
    lib().orc_g711_decode(_p(codes), _p(out), codes.size, law)
    return out


def g711_encode(pcm, law):
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    out = np.empty(pcm.shape, dtype=np.uint8)
    lib().orc_g711_encode(_p(pcm), _p(out), pcm.size, law)
    return out


def alloc_outputs(F, Cn, G):
    B = Cn // G
    out = (np.zeros((F, B, FRAME), dtype=np.int16), np.zeros((F, B, FRAME), dtype=np.uint8),
           np.zeros((F, Cn), dtype=METER_DT), np.zeros((F, B), dtype=BRIDGE_DT))
    for a in out:          # touch every page now so that a timed run does not pay first-touch faults
        a.view(np.uint8).reshape(-1)[::4096] = 0
    return out


def process_batch(codes, law, gain_q7, out_law, G, signed_char=0, threads=1, out=None):
    """codes [F][C][160] u8, law [C] u8, gain_q7 [F][C] u16, out_law [B] u8."""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    F, Cn, n = codes.shape
    assert n == FRAME and Cn % G == 0
    B = Cn // G
    law = np.ascontiguousarray(law, dtype=np.uint8)
    gain_q7 = np.ascontiguousarray(gain_q7, dtype=np.uint16)
    out_law = np.ascontiguousarray(out_law, dtype=np.uint8)
    if out is not None:
        mix, enc, meter, bmeter = out
    else:
        mix = np.zeros((F, B, FRAME), dtype=np.int16)
        enc = np.zeros((F, B, FRAME), dtype=np.uint8)
        meter = np.zeros((F, Cn), dtype=METER_DT)
        bmeter = np.zeros((F, B), dtype=BRIDGE_DT)
    b = Batch(F, B, G, _p(codes).value, _p(law).value, _p(gain_q7).value, _p(out_law).value,
              _p(mix).value, _p(enc).value, _p(meter).value, _p(bmeter).value, signed_char)
    if threads > 1:
        lib().orc_process_batch_mt(C.byref(b), threads)
    else:
        lib().orc_process_batch(C.byref(b))
    return mix, enc, meter, bmeter


def event_summary(meter, gain_q7):
    meter = np.ascontiguousarray(meter)
    gain_q7 = np.ascontiguousarray(gain_q7, dtype=np.uint16)
    F, Cn = meter.shape
    out = np.zeros(Cn, dtype=SUMMARY_DT)
    lib().orc_event_summary(_p(meter), _p(gain_q7), F, Cn, _p(out))
    return out


def summary_db(rec):
    a, mx, mn, bm = C.c_double(), C.c_double(), C.c_double(), C.c_int()
    r = np.ascontiguousarray(rec.reshape(1))
    lib().orc_summary_db(_p(r), C.byref(a), C.byref(mx), C.byref(mn), C.byref(bm))
    return a.value, mx.value, mn.value, bm.value
