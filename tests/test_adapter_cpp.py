"""The C++ open_spiel adapter: compiled and linked against libtwixt_b200.so and the
open_spiel header shim; parameter handling + renderer checked on the CPU, the reference's own state tests
(twixt_test.cc) on the GPU."""
import json
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = os.path.join(ROOT, "twixt_for_open_spiel_b200")
EXE = os.path.join(HERE, "_build", "adapter_driver")


@pytest.fixture(scope="module")
def driver():
    from twixt_for_open_spiel_b200 import _lib
    _lib.load()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    srcs = [os.path.join(HERE, "adapter_driver.cc"), os.path.join(PKG, "adapter", "twixt_b200_game.cc"),
            os.path.join(PKG, "adapter", "twixt_b200_game.h"), os.path.join(ROOT, "include", "twixt_b200.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(s) > os.path.getmtime(EXE) for s in srcs):
        subprocess.check_call([
            "g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
            "-I", os.path.join(ROOT, "oracle", "shim"), "-I", os.path.join(PKG, "adapter"),
            srcs[0], srcs[1], "-o", EXE, "-L", PKG, "-ltwixt_b200", "-Wl,-rpath," + PKG, "-lpthread", "-ldl"])
    return EXE


def test_adapter_parameters_and_renderer(driver, tmp_path):
    with open(os.path.join(HERE, "golden", "playthrough_n8.json")) as f:
        s0 = json.load(f)["states"][0]["observation_string"]
    want = tmp_path / "state0.txt"
    want.write_bytes(s0.encode("utf-8"))
    res = subprocess.run([driver, "cpu", str(want)], capture_output=True, text=True)
    assert res.returncode == 0 and "OK" in res.stdout, res.stderr


@pytest.mark.gpu
def test_adapter_reference_tests_on_gpu(driver):
    res = subprocess.run([driver, "gpu"], capture_output=True, text=True)
    assert res.returncode == 0 and "OK" in res.stdout, res.stderr
