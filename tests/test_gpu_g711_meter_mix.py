"""GPU parity (through the C ABI) for the stand-alone kernels: G.711, meters, mix.
Bit-exact for codes / PCM / integer meters; dB within 1e-4 (north_star)."""
import ctypes as C

import numpy as np
import pytest

import oracle_py as O
import igate4xsoftphonedsp_b200 as ig

pytestmark = pytest.mark.gpu
L = O.lib()
DB_TOL = 1e-4      # north_star: "RMS/dBFS must be within 1e-4 dB"


def meter_oracle(pcm):
    fr = np.ascontiguousarray(pcm.reshape(-1, 160))
    S = np.zeros(fr.shape[0], np.uint64)
    P = np.zeros(fr.shape[0], np.uint32)
    rms = np.zeros(fr.shape[0])
    pk = np.zeros(fr.shape[0])
    for i in range(fr.shape[0]):
        s, p = C.c_uint64(), C.c_uint32()
        L.orc_frame_power(fr[i].ctypes.data, 160, C.byref(s), C.byref(p))
        S[i], P[i] = s.value, p.value
        rms[i], pk[i] = L.orc_rms_dbfs(s.value, 160), L.orc_peak_dbfs(p.value)
    return S, P, rms, pk


def db_close(got, want):
    inf = np.isinf(want)
    assert np.array_equal(np.isinf(got), inf) and (got[inf] < 0).all()
    return np.abs(got[~inf].astype(np.float64) - want[~inf]).max() if (~inf).any() else 0.0


@pytest.mark.parametrize("law", [0, 1])
def test_decode_exhaustive(vp, law):
    codes = np.tile(np.arange(256, dtype=np.uint8), 10)
    assert np.array_equal(vp.g711_decode(codes, law), O.g711_decode(codes, law))


@pytest.mark.parametrize("law", [0, 1])
def test_encode_exhaustive(vp, law):
    pcm = np.arange(-32768, 32768, dtype=np.int32).astype(np.int16)
    got = vp.g711_encode(pcm, law)
    assert np.array_equal(got, O.encode_table(law))
    # the encoder works on packed pairs: every value next to partners of either sign
    for seed in range(3):
        shuf = np.random.default_rng(seed).permutation(pcm)
        assert np.array_equal(vp.g711_encode(shuf, law), O.g711_encode(shuf, law))


@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 159, 161, 4099])
def test_ragged_lengths(vp, n):
    rng = np.random.default_rng(n)
    codes = rng.integers(0, 256, n, dtype=np.uint8)
    pcm = rng.integers(-32768, 32768, n).astype(np.int16)
    for law in (0, 1):
        assert np.array_equal(vp.g711_decode(codes, law), O.g711_decode(codes, law))
        assert np.array_equal(vp.g711_encode(pcm, law), O.g711_encode(pcm, law))


def test_per_channel_law(vp):
    rng = np.random.default_rng(4)
    F, Cn = 5, 7
    law = rng.integers(0, 2, Cn).astype(np.uint8)
    codes = rng.integers(0, 256, (F, Cn, 160), dtype=np.uint8)
    pcm = rng.integers(-32768, 32768, (F, Cn, 160)).astype(np.int16)
    want_d = np.stack([O.g711_decode(codes[:, c], int(law[c])) for c in range(Cn)], axis=1)
    want_e = np.stack([O.g711_encode(pcm[:, c], int(law[c])) for c in range(Cn)], axis=1)
    assert np.array_equal(vp.g711_decode(codes, law), want_d)
    assert np.array_equal(vp.g711_encode(pcm, law), want_e)


def test_device_pointer_mode_matches_host_mode(vp):
    import torch
    rng = np.random.default_rng(8)
    codes = rng.integers(0, 256, (64, 160), dtype=np.uint8)
    t = torch.from_numpy(codes).cuda()
    got = vp.g711_decode(t, 1)
    vp.sync()
    assert got.is_cuda and np.array_equal(got.cpu().numpy(), O.g711_decode(codes, 1))
    back = vp.g711_encode(got, 1)
    vp.sync()
    assert np.array_equal(back.cpu().numpy(), O.g711_encode(O.g711_decode(codes, 1), 1))


def test_frame_meter(vp):
    rng = np.random.default_rng(5)
    pcm = rng.integers(-32768, 32768, (40, 160)).astype(np.int16)
    pcm[0] = 0                      # silence -> -inf
    pcm[1] = -32768                 # maximum sum of squares, |x| = 32768
    pcm[2] = 32767
    pcm[3, :] = 0
    pcm[3, 17] = 1                  # smallest non-zero level
    got = vp.frame_meter(pcm)
    S, P, rms, pk = meter_oracle(pcm)
    gS = got["sumsq_lo"].astype(np.uint64) | ((got["hi"] & 0xFF).astype(np.uint64) << np.uint64(32))
    assert np.array_equal(gS, S) and np.array_equal(got["hi"] >> 16, P)
    assert ((got["hi"] >> 8) & 0xFF == 0).all()
    assert db_close(got["rms_dbfs"], rms) < DB_TOL and db_close(got["peak_dbfs"], pk) < DB_TOL


@pytest.mark.parametrize("flags", [0, ig.F_SIGNED_CHAR])
def test_bytemean(vp, flags):
    rng = np.random.default_rng(6)
    for n, stride, ln in ((50, 160, 160), (33, 180, 160), (9, 24, 7), (4, 256, 256)):
        p = rng.integers(0, 256, (n, stride), dtype=np.uint8)
        got = vp.bytemean(p, ln, flags)
        want = np.array([L.orc_bytemean(np.ascontiguousarray(p[i]).ctypes.data, ln, int(flags != 0))
                         for i in range(n)], np.uint8)
        assert np.array_equal(got, want)


def test_level_percent(vp):
    v = np.array([0, 299, 300, 15000, 30000, 32767, -300, -1, 2**31 - 1], np.int32)
    assert vp.level_percent(v).tolist() == [L.orc_percent(int(x)) for x in v]


@pytest.mark.parametrize("G", [1, 2, 3, 4, 5, 8])
def test_mix(vp, G):
    rng = np.random.default_rng(10 + G)
    F, B = 3, 11
    pcm = rng.integers(-32768, 32768, (F, B * G, 160)).astype(np.int16)
    gain = rng.choice(np.array([0, 13, 64, 128, 256, 300, 1000], np.uint16), (F, B * G))
    got = vp.mix(pcm, gain, G)
    want = np.zeros((F, B, 160), np.int16)
    for f in range(F):
        for b in range(B):
            legs = (C.c_void_p * G)(*[pcm[f, b * G + g].ctypes.data for g in range(G)])
            adj = np.ascontiguousarray(gain[f, b * G:(b + 1) * G])
            L.orc_mix_frame(legs, adj.ctypes.data, G, 160, want[f, b].ctypes.data)
    assert np.array_equal(got, want)
