"""The C-ABI library loads on a CPU-only box and exports exactly what
include/igate_dsp.h declares; no compute entry point is called here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import igate4xsoftphonedsp_b200 as ig
from igate4xsoftphonedsp_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "igate_dsp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(igd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ig.load()
    declared = header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/igate_dsp.h but not exported"
    assert sorted(N.SYMBOLS) == declared, "python binding table drifted from the header"
    out = subprocess.check_output(["nm", "-D", "--defined-only", N.LIB_PATH], text=True)
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l and "igd_" in l)
    assert exported == declared, "library exports symbols the header does not declare (or vice versa)"


def test_abi_version_and_record_sizes():
    lib = ig.load()
    assert lib.igd_abi_version() == 2
    assert C.sizeof(N.BatchDesc) == 24 + 8 * 8 and C.sizeof(N.PackDesc) == 40 + 8 * 8
    assert N.METER_DT.itemsize == 16 and N.STATE_DT.itemsize == 40 and N.FIELDS_DT.itemsize == 16


def test_library_is_sm100a_only_and_has_blackwell_stores():
    sass = subprocess.run(["cuobjdump", "-lelf", N.LIB_PATH], capture_output=True, text=True)
    if sass.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in sass.stdout and "sm_90" not in sass.stdout and "sm_80" not in sass.stdout


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ig.IgdError, match="IGD_ENODEV"):
        ig.VoicePath(0)


def test_product_never_references_the_oracle():
    pkg = os.path.join(ROOT, "igate4xsoftphonedsp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_py" not in txt and "liboracle" not in txt and "igd_oracle" not in txt, f


def test_host_helpers():
    assert [ig.gain_q7(v) for v in (2.0, 0.0, 0.1, 0.5, 1.0)] == [256, 0, 13, 64, 128]
    assert ig.calltype_flags("TRx") == ig.CT_TXISH
    assert ig.calltype_flags("Rx") == ig.CT_RXONLY and ig.calltype_flags("Rxonly") == ig.CT_RXONLY
    assert ig.calltype_flags("Rxx") == 0
    assert ig.calltype_flags("Tx Idle") == ig.CT_TXISH | ig.CT_IDLE
    st = ig.make_state(3, radiocall=True, callIn=True, calltype="Rx", keepAlivePeroid=150, now_ms=77)
    assert st.shape == (3,) and (st["firstR2SPacket"] == 1).all() and (st["keepAlivePeroid"] == 150).all()
    assert (st["r2sSendtime"] == 77).all() and (st["callIn"] == 1).all() and (st["packetCnt"] == 0).all()
