"""Host build of the kernels' scalar math (csrc/igd_math.cuh, -DIGD_HOST_EMUL)
against the oracle: catches logic errors in the bit tricks without a GPU."""
import ctypes as C

import numpy as np

import hostbuild_py as H
import oracle_py as O
import tx_scenarios as T

E = H.lib()
L = O.lib()


def test_float_exponent_encoder_exhaustive():
    for law in (0, 1):
        out = np.zeros(65536, np.uint8)
        E.emul_encode_all(law, out.ctypes.data)
        assert np.array_equal(out, O.encode_table(law))


def test_decode_formulas_exhaustive():
    a = np.array([E.emul_alaw2lin(c) for c in range(256)], np.int16)
    u = np.array([E.emul_ulaw2lin(c) for c in range(256)], np.int16)
    assert np.array_equal(a, O.decode_table(0)) and np.array_equal(u, O.decode_table(1))


def test_db_maps_within_contract():
    rng = np.random.default_rng(0)
    worst = 0.0
    for s in list(rng.integers(1, 160 * 32768 * 32768, 5000)) + [1, 2, 159, 160, 161, 160 * 32768**2]:
        worst = max(worst, abs(E.emul_rms_dbfs(int(s)) - L.orc_rms_dbfs(int(s), 160)))
    for p in range(1, 32769, 7):
        worst = max(worst, abs(E.emul_peak_dbfs(p) - L.orc_peak_dbfs(p)))
    assert worst < 1e-4                                   # north_star: RMS/dBFS within 1e-4 dB
    assert E.emul_rms_dbfs(0) == -np.inf and E.emul_peak_dbfs(0) == -np.inf


def test_bytemean_division_matches_c():
    for s, n in ((0, 160), (160 * 255, 160), (-111, 3), (-1, 160), (51, 3)):
        assert E.emul_bytemean(s, n) == (int(s / n) & 0xFF)


def test_field_extraction():
    rng = np.random.default_rng(2)
    for w in [0, 0x00013100, 0x104131F8, 0xFFFFFFFF] + [int(x) for x in rng.integers(0, 2**32, 300)]:
        o = (C.c_uint * 5)()
        E.emul_fields(w, o)
        f = O.Fields()
        L.orc_ed137_fields_from_word(w, C.byref(f))
        flags = f.active | (f.rrc_present << 1) | (f.main_tx_used << 2) | (f.main_rx_used << 3)
        assert list(o) == [f.ptt_type, f.ptt_id, f.squelch, f.bss, flags]


def test_sender_state_machine_all_scenarios():
    for s in T.SCENARIOS:
        pk, sizes, _, ads = T.run_oracle(s)
        st = T.gpu_inputs(s)
        F, Cn = sizes.shape
        for c in range(Cn):
            one = st[c:c + 1].copy()
            for f in range(F):
                if s["ctl"] is not None:
                    k = s["ctl"][f, c]
                    for n in ("pttstatus", "pttpriority", "callRecorder", "sqlstatus", "ed137_bssi", "pttid"):
                        one[n] = k[n]
                o = (C.c_uint * 5)()
                E.emul_tx_step(one.ctypes.data, 160, s["now0"] + f * s["tick_ms"], o)
                n = int(sizes[f, c])
                assert o[1] == n, (s["name"], c, f)
                if n:
                    w = int.from_bytes(pk[f, c, 16:20].tobytes(), "big")
                    exp_pt = 123 if o[2] else int(s["rtp12"][f, c, 1] & 0x7F)
                    assert (w, int(pk[f, c, 1] & 0x7F), int(pk[f, c, 1] >> 7)) == (o[0], exp_pt, o[3]), (s["name"], c, f)
            a = ads[c]
            for name in ("packetCnt", "firstR2SPacket", "trxSlaveEnableChangedCount", "r2sSendtime",
                         "rxSlaveEnable", "txSlaveEnable"):
                assert int(one[name][0]) == int(getattr(a, name)), (s["name"], c, name)
