"""The two fused kernels for G <= 4 -- k_fused_q (quarter-frame lanes, the default of igd_process_batch) and k_fused_w
(five lanes per bridge-frame, IGD_F_KERNEL_W; the packet forms run it) -- against the oracle on the same cases: every
leg count, ragged item tails, both char signs, arbitrary and NO_AUDIO gains, optional outputs; and byte-for-byte
against each other at a size that fills the grid."""
import numpy as np
import pytest

import oracle_py as O
import igate4xsoftphonedsp_b200 as ig
from igate4xsoftphonedsp_b200 import _native as N
from igate4xsoftphonedsp_b200 import synth
from test_gpu_fused import check, make

pytestmark = pytest.mark.gpu
KERNELS = [0, N.F_KERNEL_W]


@pytest.mark.parametrize("G,B,F", [(4, 37, 9), (4, 32, 1), (4, 1, 1), (4, 3, 3), (2, 45, 7), (1, 100, 3), (3, 21, 5),
                                   (1, 1, 1), (2, 1, 9), (3, 1, 7), (4, 1024, 5)])
@pytest.mark.parametrize("signed", [0, 1])
@pytest.mark.parametrize("Q", KERNELS)
def test_q_random_codes(vp, G, B, F, signed, Q):
    codes, law, _, out_law = make(F, B, G, random_codes=True, seed=G * 100 + B)
    rng = np.random.default_rng(7 + signed)
    gain = rng.choice(np.array([0, 0, 13, 64, 128, 256, 300], np.uint16), (F, B * G))
    law = rng.integers(0, 2, B * G).astype(np.uint8)
    got = vp.process_batch(codes, law, gain, out_law, G, flags=Q | (ig.F_SIGNED_CHAR if signed else 0))
    check(got, O.process_batch(codes, law, gain, out_law, G, signed_char=signed))


@pytest.mark.parametrize("G", [1, 2, 3, 4])
@pytest.mark.parametrize("Q", KERNELS)
def test_q_gates_only_fast_path(vp, G, Q):
    """gains 0 / 256 only (the PTT gates of the bench workload): the selector path, no multiply"""
    F, B = 11, 29
    codes, law, gain, out_law = make(F, B, G, random_codes=True, seed=G)
    assert set(np.unique(gain)) <= {0, 256}
    check(vp.process_batch(codes, law, gain, out_law, G, flags=Q), O.process_batch(codes, law, gain, out_law, G))


@pytest.mark.parametrize("G", [1, 2, 3, 4])
@pytest.mark.parametrize("Q", KERNELS)
def test_q_no_audio_legs(vp, G, Q):
    F, B = 6, 19
    codes, law, gain, out_law = make(F, B, G, random_codes=True, seed=40 + G)
    rng = np.random.default_rng(G)
    gain = gain | np.where(rng.random(gain.shape) < 0.2, N.GAIN_NO_AUDIO, 0).astype(np.uint16)
    check(vp.process_batch(codes, law, gain, out_law, G, flags=Q), O.process_batch(codes, law, gain, out_law, G))


@pytest.mark.parametrize("want", [("enc",), ("mix", "bmeter"), ("meter",), ("enc", "meter", "bmeter")])
@pytest.mark.parametrize("Q", KERNELS)
def test_q_optional_outputs(vp, want, Q):
    F, B, G = 5, 23, 4
    codes, law, gain, out_law = make(F, B, G, random_codes=True, seed=5)
    ref = dict(zip(("mix", "enc", "meter", "bmeter"), O.process_batch(codes, law, gain, out_law, G)))
    got = vp.process_batch(codes, law, gain, out_law, G, flags=Q, want=want)
    assert set(got) == set(want)
    for k in want:
        if k == "meter":
            assert np.array_equal(got[k]["sumsq_lo"], ref[k]["sumsq_lo"]) and np.array_equal(got[k]["hi"], ref[k]["hi"])
        else:
            assert got[k].tobytes() == ref[k].tobytes()


@pytest.mark.parametrize("G", [1, 2, 3, 4])
def test_q_equals_w_full_grid(vp, G):
    """more items than warps in the grid (148 x 24 x 8 bridge-frames), ragged tail: several trips through every slot"""
    B, F = 1021, 131
    codes, law, gain, out_law = make(F, B, G, random_codes=True, seed=77)
    a = vp.process_batch(codes, law, gain, out_law, G)
    b = vp.process_batch(codes, law, gain, out_law, G, flags=N.F_KERNEL_W)
    for k in ("mix", "enc", "meter", "bmeter"):
        assert a[k].tobytes() == b[k].tobytes(), k


# ---- eight legs run k_fused_h (default); the other multiples of 8 and IGD_F_KERNEL_W run k_fused_g: both against the oracle
@pytest.mark.parametrize("G,B,F", [(8, 13, 4), (8, 1, 1), (8, 4, 3), (16, 7, 5), (16, 2, 1), (24, 5, 3), (32, 3, 2), (32, 1, 7),
                                   (8, 333, 3)])
@pytest.mark.parametrize("signed", [0, 1])
def test_h_random_codes(vp, G, B, F, signed):
    codes, law, _, out_law = make(F, B, G, random_codes=True, seed=G * 100 + B)
    rng = np.random.default_rng(17 + signed)
    gain = rng.choice(np.array([0, 0, 13, 64, 128, 256, 300], np.uint16), (F, B * G))
    law = rng.integers(0, 2, B * G).astype(np.uint8)
    want = O.process_batch(codes, law, gain, out_law, G, signed_char=signed)
    for kern in KERNELS:
        check(vp.process_batch(codes, law, gain, out_law, G, flags=kern | (ig.F_SIGNED_CHAR if signed else 0)), want)


@pytest.mark.parametrize("G", [8, 16, 24, 32])
def test_h_gates_only_and_no_audio(vp, G):
    F, B = 9, 11
    codes, law, gain, out_law = make(F, B, G, random_codes=True, seed=G)
    check(vp.process_batch(codes, law, gain, out_law, G), O.process_batch(codes, law, gain, out_law, G))
    rng = np.random.default_rng(G)
    gain = gain | np.where(rng.random(gain.shape) < 0.2, N.GAIN_NO_AUDIO, 0).astype(np.uint16)
    check(vp.process_batch(codes, law, gain, out_law, G), O.process_batch(codes, law, gain, out_law, G))
    one = np.zeros_like(gain)                      # a CLIENT softphone: one of the G call slots open
    one[:, ::G] = 256
    check(vp.process_batch(codes, law, one, out_law, G), O.process_batch(codes, law, one, out_law, G))


@pytest.mark.parametrize("want", [("enc",), ("mix", "bmeter"), ("meter",)])
def test_h_optional_outputs(vp, want):
    F, B, G = 5, 9, 8
    codes, law, gain, out_law = make(F, B, G, random_codes=True, seed=5)
    ref = dict(zip(("mix", "enc", "meter", "bmeter"), O.process_batch(codes, law, gain, out_law, G)))
    got = vp.process_batch(codes, law, gain, out_law, G, want=want)
    assert set(got) == set(want)
    for k in want:
        if k == "meter":
            assert np.array_equal(got[k]["sumsq_lo"], ref[k]["sumsq_lo"]) and np.array_equal(got[k]["hi"], ref[k]["hi"])
        else:
            assert got[k].tobytes() == ref[k].tobytes()


@pytest.mark.parametrize("G,B,F", [(8, 509, 67), (8, 1021, 131)])
def test_h_equals_g_full_grid(vp, G, B, F):
    """more items than warps in the grid, ragged tail: several trips through every slot"""
    codes, law, gain, out_law = make(F, B, G, random_codes=True, seed=78)
    a = vp.process_batch(codes, law, gain, out_law, G)
    b = vp.process_batch(codes, law, gain, out_law, G, flags=N.F_KERNEL_W)
    for k in ("mix", "enc", "meter", "bmeter"):
        assert a[k].tobytes() == b[k].tobytes(), k
