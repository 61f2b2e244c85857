"""examples/igd_tick.c: the C ABI used from plain C99 (the header must be valid C, not only C++)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "examples", "igd_tick.c")
PKG = os.path.join(ROOT, "igate4xsoftphonedsp_b200")
BIN = os.path.join(ROOT, "tests", "host_cpp", "igd_tick")


def build():
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           SRC, "-L" + PKG, "-ligate_dsp", "-Wl,-rpath," + PKG, "-o", BIN])


def test_header_is_valid_c99_and_the_example_links():
    build()
    r = subprocess.run([BIN], capture_output=True, text=True)
    # without an sm_100 GPU the library must refuse loudly (IGD_ENODEV = -19), never fall back
    assert r.returncode in (0, 1)
    if r.returncode == 1:
        assert "igd_init failed: -19" in r.stderr


@pytest.mark.gpu
def test_example_runs_one_tick_on_the_gpu():
    build()
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = r.stdout.splitlines()
    # four radios keying PTT types 0..3 in one tick: the priority PTT (leg 3) wins (roip_ed137.cpp:6157-6177)
    assert [ln.split("gain_q7 ")[1].split()[0] for ln in out[:4]] == ["0", "0", "0", "256"]
    assert all("bytemean 213 peak 8" in ln for ln in out[:4])          # A-law silence 0xD5 decodes to +8
    assert out[4] == "bridge: 1 legs open, outgoing byte-mean 212, mix[0] 16"   # 2.0 * 8 = 16 -> A-law 0xD4
