"""Builds and binds tests/hostbuild/emul.cpp (host build of igd_math.cuh). Test-only."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostbuild", "emul.cpp")
HDR = os.path.join(HERE, "..", "igate4xsoftphonedsp_b200", "csrc", "igd_math.cuh")
OUT = os.path.join(HERE, "hostbuild", "libemul.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        newest = max(os.path.getmtime(SRC), os.path.getmtime(HDR))
        if not os.path.exists(OUT) or os.path.getmtime(OUT) < newest:
            subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-o", OUT, SRC])
        L = C.CDLL(OUT)
        L.emul_alaw2lin.argtypes = [C.c_uint]
        L.emul_ulaw2lin.argtypes = [C.c_uint]
        L.emul_encode_all.argtypes = [C.c_int, C.c_void_p]
        L.emul_rms_dbfs.restype = C.c_float
        L.emul_rms_dbfs.argtypes = [C.c_ulonglong]
        L.emul_peak_dbfs.restype = C.c_float
        L.emul_peak_dbfs.argtypes = [C.c_uint]
        L.emul_bytemean.restype = C.c_uint
        L.emul_bytemean.argtypes = [C.c_int, C.c_int]
        L.emul_fields.argtypes = [C.c_uint, C.POINTER(C.c_uint)]
        L.emul_tx_step.argtypes = [C.c_void_p, C.c_uint, C.c_longlong, C.POINTER(C.c_uint)]
        L.emul_r2s_step.argtypes = [C.c_void_p, C.c_longlong, C.c_uint, C.POINTER(C.c_uint)]
        L.emul_rx_step.restype = C.c_uint
        L.emul_rx_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_int]
        L.emul_arb_tick.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib = L
    return _lib
