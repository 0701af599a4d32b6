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


def test_adapter_parameters_and_renderer(driver, tmp_path, oracle_mod):
    """Parameter errors; the C++ renderer on the empty board AND on every kept position of the
    reference-generated string fixture (records from the oracle: links of all eight directions, both
    colours, swap / win / draw suffixes, ANSI on and off)."""
    with open(os.path.join(HERE, "golden", "playthrough_n8.json")) as f:
        s0 = json.load(f)["states"][0]["observation_string"]
    want = tmp_path / "state0.txt"
    want.write_bytes(s0.encode("utf-8"))
    with open(os.path.join(HERE, "golden", "ref_strings.json")) as f:
        games = json.load(f)["games"]
    recs = tmp_path / "records.txt"
    with open(recs, "wb") as out:
        for g in games:
            og = oracle_mod.OracleGame(g["n"])
            for ply in sorted(g["strings"], key=int):
                st = og.new_initial_state()
                st.replay(g["actions"][:int(ply)])
                rec = st.export_record()
                raw = g["strings"][ply].encode("utf-8")
                out.write(("REC %d %d %d %d\n" % (g["n"], 1 if g["ansi"] else 0, len(rec), len(raw))).encode())
                out.write((" ".join(str(int(w)) for w in rec) + "\n").encode())
                out.write(raw)
                out.write(b"\n")
        out.write(b"END\n")
    res = subprocess.run([driver, "cpu", str(want), str(recs)], capture_output=True, text=True)
    assert res.returncode == 0 and "OK" in res.stdout, res.stderr


def write_strings_fixture(path):
    """tests/golden/ref_strings.json as the plain-text file adapter_driver reads."""
    with open(os.path.join(HERE, "golden", "ref_strings.json")) as f:
        games = json.load(f)["games"]
    with open(path, "wb") as out:
        for g in games:
            out.write(("GAME %d %d %d\n" % (g["n"], 1 if g["ansi"] else 0, len(g["actions"]))).encode())
            out.write((" ".join(str(a) for a in g["actions"]) + "\n").encode())
            out.write((" ".join(str(c) for c in g["crcs"]) + "\n").encode())
            out.write(("%d\n" % len(g["strings"])).encode())
            for ply in sorted(g["strings"], key=int):
                raw = g["strings"][ply].encode("utf-8")
                out.write(("%s %d\n" % (ply, len(raw))).encode())
                out.write(raw)
                out.write(b"\n")
        out.write(b"END\n")
    return len(games)


@pytest.mark.gpu
def test_adapter_reference_tests_on_gpu(driver, tmp_path):
    """twixt_test.cc's state tests, then ToString at every ply of the reference-generated string fixture
    (swap, both wins, draw, all eight link directions, ANSI on/off) through the C++ adapter on cuda:0."""
    fixture = tmp_path / "ref_strings.txt"
    assert write_strings_fixture(str(fixture)) >= 10
    res = subprocess.run([driver, "gpu", str(fixture)], capture_output=True, text=True)
    assert res.returncode == 0 and "OK" in res.stdout, res.stderr
