"""The product's per-env rules (csrc/twixt_engine.cuh + twixt_philox.cuh) compiled for the HOST by
tests/host_engine_harness.cc and compared with the oracle -- checks the bit-plane rules, crossing masks,
k-th legal selection, flood fill (incl. its stack-overflow branch, built with a 2-entry stack) and the
Philox stream without a GPU.  The product library itself never runs on the CPU."""
import ctypes as C
import json
import os
import random
import subprocess
import zlib

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "twixt_for_open_spiel_b200", "csrc")


def _build(stack):
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libhostengine_s%d.so" % stack)
    srcs = [os.path.join(HERE, "host_engine_harness.cc")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-x", "c++", "-fPIC", "-shared",
                               "-DTW_TEST_STACK=%d" % stack, "-I", CSRC, "-o", so,
                               os.path.join(HERE, "host_engine_harness.cc")])
    he = C.CDLL(so)
    he.he_record_words.restype = C.c_int
    he.he_init.argtypes = [C.c_void_p, C.c_int]
    he.he_legal_actions.restype = C.c_int
    he.he_legal_actions.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    he.he_legal_count.argtypes = [C.c_void_p, C.c_int]
    he.he_apply.argtypes = [C.c_void_p, C.c_int, C.c_int]
    he.he_current_player.argtypes = [C.c_void_p, C.c_int]
    he.he_observation.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    he.he_playout.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
    he.he_playout_cached.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
    he.he_playout_smem.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
    he.he_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    he.he_select_bit.argtypes = [C.c_uint32, C.c_int]
    return he


@pytest.fixture(scope="module", params=[48, 2], ids=["stack48", "stack2"])
def he(request):
    return _build(request.param)


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def test_crossing_include_is_up_to_date():
    assert subprocess.call(["python", os.path.join(ROOT, "tools", "gen_crossing_table.py"), "--check"]) == 0


def test_select_bit(he):
    rng = random.Random(1)
    for _ in range(2000):
        w = rng.getrandbits(24) | 1
        bits = [i for i in range(32) if (w >> i) & 1]
        k = rng.randrange(len(bits))
        assert he.he_select_bit(w, k) == bits[k]


def test_philox_matches_known_answers(he):
    with open(os.path.join(HERE, "golden", "kats.json")) as f:
        for kat in json.load(f)["philox4x32_10"]:
            out = np.zeros(4, dtype=np.uint32)
            he.he_philox(P(np.asarray(kat["ctr"], dtype=np.uint32)), P(np.asarray(kat["key"], dtype=np.uint32)), P(out))
            assert out.tolist() == kat["out"]


@pytest.mark.parametrize("n", list(range(5, 25)))
def test_rules_match_oracle(he, oracle_mod, n):
    og = oracle_mod.OracleGame(n)
    R = he.he_record_words(n)
    assert R == og.record_words()
    rng = random.Random(100 + n)
    for gi in range(24 if n <= 12 else 6):
        st = og.new_initial_state()
        rec = np.zeros(R, dtype=np.uint32)
        he.he_init(P(rec), n)
        ply, first = 0, None
        while True:
            assert np.array_equal(rec, st.export_record()), (n, gi, ply)
            out = np.zeros(n * n, dtype=np.int64)
            c = he.he_legal_actions(P(rec), n, P(out))
            lb = st.legal_actions()
            assert out[:c].tolist() == lb and he.he_legal_count(P(rec), n) == len(lb)
            assert he.he_current_player(P(rec), n) == st.current_player()
            if ply % 5 == 0 or st.is_terminal():
                obs = np.zeros(12 * n * (n - 2), dtype=np.float32)
                he.he_observation(P(rec), n, P(obs))
                assert np.array_equal(obs, st.observation_tensor(0)), (n, gi, ply)
            if st.is_terminal():
                break
            bad = rng.randrange(-2, n * n + 3)
            if bad not in lb:
                before = rec.copy()
                assert he.he_apply(P(rec), n, bad) == 1 and np.array_equal(before, rec)
            a = first if (ply == 1 and gi % 3 == 0 and first in lb) else rng.choice(lb)
            if ply == 0:
                first = a
            assert he.he_apply(P(rec), n, a) == 0
            st.apply_action(a)
            ply += 1
    # the fused-playout policy: same Philox stream, same k-th legal pick -- with the plain column scan, with
    # the kernel's structure (per-column count cache, moves interleaved with flood visits, swap run ahead of
    # the move) and with that structure on the kernel's shared-memory layout (no bounds tests on the link
    # window).  Small boards get enough streams to contain swaps (1 game in n(n-2)).
    swaps = 0
    for fn in (he.he_playout, he.he_playout_cached, he.he_playout_smem):
        for s in range(90 if n <= 6 else 12):
            rec = np.zeros(R, dtype=np.uint32)
            he.he_init(P(rec), n)
            st = og.new_initial_state()
            if s % 3 == 2:  # from a mid-game position (s = 5, 11, ..: taken over at ply 1, before a possible swap)
                pre = st.playout_philox(7, s, 1 if s % 6 == 5 else 2 + s % 40)
                for a in pre:
                    assert he.he_apply(P(rec), n, a) == 0
            acts = np.zeros(n * n, dtype=np.int64)
            L = fn(P(rec), n, 0x7477697854, s + (n << 33), 1 << 30, P(acts))
            assert acts[:L].tolist() == st.playout_philox(0x7477697854, s + (n << 33)), (n, s)
            assert np.array_equal(rec, st.export_record())
            swaps += int(rec[1] >> 2) & 1
    if n <= 6:
        assert swaps >= 3  # the swap path was really taken


def test_rules_match_reference_generated_fixture(he):
    with open(os.path.join(HERE, "golden", "ref_games.json")) as f:
        data = json.load(f)
    for g in data["games"]:
        n = g["n"]
        rec = np.zeros(he.he_record_words(n), dtype=np.uint32)
        he.he_init(P(rec), n)
        for ply, (player, count, crc_l, crc_o) in enumerate(g["plies"]):
            out = np.zeros(n * n, dtype=np.int64)
            c = he.he_legal_actions(P(rec), n, P(out))
            assert c == count and he.he_current_player(P(rec), n) == player
            assert (zlib.crc32(out[:c].tobytes()) & 0xFFFFFFFF) == crc_l
            obs = np.zeros(12 * n * (n - 2), dtype=np.float32)
            he.he_observation(P(rec), n, P(obs))
            assert (zlib.crc32(obs.tobytes()) & 0xFFFFFFFF) == crc_o
            if ply < len(g["actions"]):
                assert he.he_apply(P(rec), n, g["actions"][ply]) == 0
