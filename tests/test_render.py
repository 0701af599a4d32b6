"""ToString parity: the host renderer over an exported record reproduces the reference's
board picture byte for byte -- against the golden playthrough strings and against the compiled reference."""
import json
import os
import random

import pytest

from helpers import random_game_actions

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_playthrough_strings(oracle_mod):
    from twixt_for_open_spiel_b200.render import board_to_string
    with open(os.path.join(GOLDEN, "playthrough_n8.json")) as f:
        pt = json.load(f)
    st = oracle_mod.OracleGame(8).new_initial_state()
    by_index = {s["index"]: s for s in pt["states"]}
    checked = 0
    for ply in range(36):
        s = by_index.get(ply)
        if s is not None and "observation_string" in s:
            assert board_to_string(st.export_record(), 8, True) == s["observation_string"], ply
            checked += 1
        if ply < 35:
            st.apply_action(pt["actions"][ply])
    assert checked >= 9
    assert board_to_string(st.export_record(), 8, True).endswith("[x has won]")


@pytest.mark.parametrize("n,ansi", [(5, True), (6, False), (8, True), (11, False), (12, True), (24, True), (24, False)])
def test_matches_compiled_reference(oracle_mod, have_ref, n, ansi):
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from twixt_for_open_spiel_b200.render import board_to_string
    rg = oracle_mod.RefGame(n, ansi)
    og = oracle_mod.OracleGame(n)
    proto = rg.new_initial_state()
    rng = random.Random(n)
    for gi in range(6 if n <= 12 else 2):
        acts = random_game_actions(og, rng, force_swap=(gi % 2 == 0))
        rs, os_ = proto.clone(), og.new_initial_state()
        for ply, a in enumerate(acts):
            if ply % 4 == 0:
                assert board_to_string(os_.export_record(), n, ansi) == rs.to_string(), (n, gi, ply)
            rs.apply_action(a)
            os_.apply_action(a)
        assert board_to_string(os_.export_record(), n, ansi) == rs.to_string()
        del rs
    del proto, rg


def test_action_to_string_matches_reference(oracle_mod, have_ref):
    """twixt.cc:67-74; also pinned by StringLegalActions in the playthrough (GPU test)."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from twixt_for_open_spiel_b200.spiel import TwixTState
    for n in (5, 8, 12, 24):
        rg = oracle_mod.RefGame(n)
        rs = rg.new_initial_state()

        class _G:
            def board_size(self):
                return n
        fake = TwixTState.__new__(TwixTState)
        fake._game = _G()
        for a in range(n * n):
            for p in (0, 1):
                assert TwixTState.action_to_string(fake, p, a) == rs.action_to_string(p, a)
        fake._game = None
        del rs, rg


def test_reference_generated_strings(oracle_mod):
    """tests/golden/ref_strings.json (the compiled reference's ToString at every ply of swap / win / draw and
    random games, all eight link directions, ANSI on and off) against the host renderer over oracle records."""
    import zlib
    from twixt_for_open_spiel_b200.render import board_to_string
    with open(os.path.join(GOLDEN, "ref_strings.json")) as f:
        games = json.load(f)["games"]
    dirs = 0
    for g in games:
        st = oracle_mod.OracleGame(g["n"]).new_initial_state()
        for ply in range(len(g["actions"]) + 1):
            text = board_to_string(st.export_record(), g["n"], g["ansi"])
            assert zlib.crc32(text.encode("utf-8")) & 0xFFFFFFFF == g["crcs"][ply], (g["n"], ply)
            if str(ply) in g["strings"]:
                assert text == g["strings"][str(ply)]
            if ply < len(g["actions"]):
                st.apply_action(g["actions"][ply])
        dirs |= g["link_directions"]
    assert dirs == 0xFF
