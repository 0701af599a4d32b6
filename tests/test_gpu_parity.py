"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle, bit for bit.

Every test plays oracle-generated games -- including swaps, draws and illegal
moves -- through libtwixt_b200 in lock-step with the oracle and compares, at
every ply: the ascending legal-action list, the legal mask, current player,
terminal flag, returns, the observation tensor and the complete packed state
record (pegs, links, blocked flags, border flags, counters).
"""
import os
import random

import numpy as np
import pytest

from helpers import SEED, draw_seeking_actions, pad_games, random_game_actions

pytestmark = pytest.mark.gpu


def _lockstep(oracle_mod, n, games, check_obs_every=1):
    """Replay `games` (lists of actions) on the GPU batch and the oracle in lock-step."""
    from twixt_for_open_spiel_b200 import TwixTBatch
    og = oracle_mod.OracleGame(n)
    E = len(games)
    batch = TwixTBatch(n, E, 0, SEED)
    states = [og.new_initial_state() for _ in range(E)]
    acts = pad_games(games)
    maxlen = acts.shape[1]
    for ply in range(maxlen + 1):
        la, cnt = batch.legal_actions()
        mask = batch.legal_mask()
        player = batch.current_player()
        term = batch.is_terminal()
        rets = batch.returns()
        recs = batch.export_state()
        obs = batch.observation() if ply % check_obs_every == 0 or ply == maxlen else None
        for e, st in enumerate(states):
            ol = st.legal_actions()
            assert la[e, :cnt[e]].tolist() == ol, (n, e, ply)
            assert cnt[e] == len(ol)
            m = np.zeros(n * n, dtype=np.uint8)
            m[ol] = 1
            assert np.array_equal(mask[e], m), (n, e, ply)
            assert int(player[e]) == st.current_player(), (n, e, ply)
            assert bool(term[e]) == st.is_terminal()
            assert rets[e].tolist() == st.returns()
            assert np.array_equal(recs[e], st.export_record()), (n, e, ply)
            if obs is not None:
                assert np.array_equal(obs[e].reshape(-1), st.observation_tensor(0)), (n, e, ply)
        if ply == maxlen:
            break
        step = acts[:, ply].copy()
        status = batch.apply(step)
        for e, st in enumerate(states):
            if step[e] >= 0:
                assert status[e] == 0
                st.apply_action(int(step[e]))
            else:
                assert status[e] == 2
    batch.close()


@pytest.mark.parametrize("n", [5, 6, 7, 8, 9, 10, 11, 12, 13, 16, 19, 23, 24])
def test_lockstep_random_games(oracle_mod, n):
    og = oracle_mod.OracleGame(n)
    rng = random.Random(1000 + n)
    num = 96 if n <= 12 else 40
    games = [random_game_actions(og, rng, force_swap=(i % 3 == 0)) for i in range(num)]
    _lockstep(oracle_mod, n, games, check_obs_every=1 if n <= 12 else 3)


@pytest.mark.parametrize("n", list(range(5, 25)))
def test_observation_every_size_late_game(oracle_mod, n):
    """All 12 planes (incl. the sparse blocked-link planes 5/11) on crowded boards, every board size."""
    from twixt_for_open_spiel_b200 import TwixTBatch
    og = oracle_mod.OracleGame(n)
    E = 64
    batch = TwixTBatch(n, E, 0, SEED)
    budget = max(4, (n * n * 3) // 5)
    batch.playout(max_plies=budget)
    obs = batch.observation()
    seen = np.zeros(12, dtype=bool)
    for e in range(E):
        st = og.new_initial_state()
        st.playout_philox(SEED, e, budget)
        want = st.observation_tensor(0)
        assert np.array_equal(obs[e].reshape(-1), want), (n, e)
        seen |= want.reshape(12, -1).any(axis=1)
    assert seen.all(), seen  # every plane carried data, so every plane was really compared
    batch.close()


def test_lockstep_draw_seeking_n5(oracle_mod):
    og = oracle_mod.OracleGame(5)
    games = [draw_seeking_actions(og, p) for p in [(0, 1), (0, 0), (10**6, 10**6), (1, 0), (2, 3), (0, 10**6)]]
    assert games[0] == [5, 2, 6, 3, 7, 8, 9, 11, 10, 12, 13, 16, 14, 17, 15, 18, 19, 21]
    _lockstep(oracle_mod, 5, games)


def test_n5_all_first_moves_with_and_without_swap(oracle_mod):
    """BASELINE config C2: every first move x {swap where legal, two other replies} x random continuations."""
    og = oracle_mod.OracleGame(5)
    rng = random.Random(5)
    games = []
    first_moves = og.new_initial_state().legal_actions()
    assert len(first_moves) == 15
    for f in first_moves:
        st = og.new_initial_state()
        st.apply_action(f)
        replies = st.legal_actions()
        picks = ([f] if f in replies else []) + rng.sample([a for a in replies if a != f], 2)
        for r in picks:
            for _ in range(6):
                s2 = st.clone()
                s2.apply_action(r)
                g = [f, r]
                while not s2.is_terminal():
                    a = rng.choice(s2.legal_actions())
                    s2.apply_action(a)
                    g.append(a)
                games.append(g)
    _lockstep(oracle_mod, 5, games)


def test_illegal_actions_leave_state_untouched(oracle_mod):
    from twixt_for_open_spiel_b200 import SpielFatalError, TwixTBatch
    n = 8
    og = oracle_mod.OracleGame(n)
    rng = random.Random(77)
    E = 64
    batch = TwixTBatch(n, E, 0, SEED)
    states = [og.new_initial_state() for _ in range(E)]
    for ply in range(40):
        acts = np.zeros(E, dtype=np.int32)
        legal_flags = []
        for e, st in enumerate(states):
            la = st.legal_actions()
            if st.is_terminal() or rng.random() < 0.3:
                a = rng.randrange(0, n * n + 5)  # may be legal by chance; may be out of range
            else:
                a = rng.choice(la)
            acts[e] = a
            legal_flags.append(a in la)
        before = batch.export_state()
        status = batch.apply(acts, raise_on_illegal=False)
        after = batch.export_state()
        first_bad = None
        for e, st in enumerate(states):
            if legal_flags[e]:
                assert status[e] == 0
                st.apply_action(int(acts[e]))
            else:
                assert status[e] == 1
                assert np.array_equal(before[e], after[e])
                if first_bad is None:
                    first_bad = int(acts[e])
            assert np.array_equal(after[e], st.export_record())
        if first_bad is not None:  # the reference's message, twixt.h:96
            snapshot = batch.export_state()
            with pytest.raises(SpielFatalError, match=r"^Not a legal action: %d$" % first_bad):
                bad_only = np.where(np.array(legal_flags), -1, acts).astype(np.int32)
                batch.apply(bad_only)
            assert np.array_equal(snapshot, batch.export_state())
    batch.close()


def test_device_pointer_path_matches_host_path(oracle_mod):
    """Outputs written straight into torch CUDA tensors equal the staged host outputs."""
    import torch
    from twixt_for_open_spiel_b200 import TwixTBatch
    n = 12
    og = oracle_mod.OracleGame(n)
    rng = random.Random(3)
    games = [random_game_actions(og, rng, force_swap=(i % 2 == 0), max_plies=30 + i % 20) for i in range(50)]
    E = len(games)
    batch = TwixTBatch(n, E, 0, SEED)
    acts = pad_games(games)
    dev = torch.device("cuda:0")
    for ply in range(acts.shape[1]):
        a_dev = torch.from_numpy(acts[:, ply].copy()).to(dev)
        st_dev = torch.full((E,), 9, dtype=torch.int32, device=dev)
        batch.apply(a_dev, out_status=st_dev)
    batch.synchronize()
    la_h, cnt_h = batch.legal_actions()
    la_d = torch.full((E, batch.max_legal_actions), -1, dtype=torch.int64, device=dev)
    cnt_d = torch.zeros(E, dtype=torch.int32, device=dev)
    batch.legal_actions(out_actions=la_d, out_counts=cnt_d)
    la16_d = torch.zeros((E, batch.max_legal_actions), dtype=torch.int16, device=dev)
    batch.legal_actions(out_actions=la16_d, out_counts=cnt_d)
    mask_d = torch.zeros((E, n * n), dtype=torch.uint8, device=dev)
    batch.legal_mask(out=mask_d)
    obs_d = torch.empty((E,) + batch.obs_shape, dtype=torch.float32, device=dev)
    batch.observation(out=obs_d)
    ret_d = torch.zeros((E, 2), dtype=torch.float32, device=dev)
    batch.returns(out=ret_d)
    batch.synchronize()
    assert np.array_equal(cnt_d.cpu().numpy(), cnt_h)
    for e in range(E):
        assert np.array_equal(la_d[e, :cnt_h[e]].cpu().numpy(), la_h[e, :cnt_h[e]])
        assert np.array_equal(la16_d[e, :cnt_h[e]].cpu().numpy().astype(np.int64), la_h[e, :cnt_h[e]])
    assert np.array_equal(mask_d.cpu().numpy(), batch.legal_mask())
    assert np.array_equal(obs_d.cpu().numpy(), batch.observation())
    assert np.array_equal(ret_d.cpu().numpy(), batch.returns())
    # unaligned float output takes the scalar-store kernel
    raw = torch.empty(E * batch.info.obs_size + 1, dtype=torch.float32, device=dev)
    batch.observation(out=raw[1:])
    batch.synchronize()
    assert np.array_equal(raw[1:].cpu().numpy().reshape(obs_d.shape), obs_d.cpu().numpy())
    batch.close()


@pytest.mark.parametrize("n", [5, 8, 12, 17, 24])
def test_fused_playout_replays_on_oracle(oracle_mod, n):
    """K5: every game the fused kernel plays is regenerated move for move by the oracle's
    restatement of the Philox policy; final records, returns and lengths agree."""
    from twixt_for_open_spiel_b200 import TwixTBatch
    og = oracle_mod.OracleGame(n)
    E = 300 if n <= 12 else 160
    batch = TwixTBatch(n, E, 0, SEED)
    batch.set_stream_base(1 << 20)
    rets, lens, trace = batch.playout(trace=True)
    recs = batch.export_state()
    stats = batch.stats()
    assert stats["games"] == E and stats["plies"] == int(lens.sum())
    assert stats["red_wins"] + stats["blue_wins"] + stats["draws"] == E
    assert bool(batch.is_terminal().all())
    for e in range(E):
        st = og.new_initial_state()
        oa = st.playout_philox(SEED, (1 << 20) + e)
        assert lens[e] == len(oa), (n, e)
        assert trace[:lens[e], e].tolist() == oa, (n, e)
        assert (trace[lens[e]:, e] == 0xFFFF).all()
        assert rets[e].tolist() == st.returns()
        assert np.array_equal(recs[e], st.export_record()), (n, e)
    assert stats["max_length"] == int(lens.max())
    batch.close()


def test_playout_from_midgame_clones_with_stream_ids(oracle_mod):
    """BASELINE config C3 shape: leaves cloned x4, rolled out with explicit stream ids and a ply budget."""
    from twixt_for_open_spiel_b200 import TwixTBatch
    n = 12
    og = oracle_mod.OracleGame(n)
    rng = random.Random(12)
    B, R = 48, 4
    batch = TwixTBatch(n, B + B * R, 0, SEED)
    leaves, games = [], []
    for i in range(B):
        g = random_game_actions(og, rng, force_swap=(i % 4 == 0), max_plies=rng.randrange(0, 61))
        st = og.new_initial_state()
        st.replay(g)
        leaves.append(st)
        games.append(g)
    acts = pad_games(games)
    for ply in range(acts.shape[1]):
        batch.apply(acts[:, ply].copy(), 0)
    src = np.repeat(np.arange(B, dtype=np.int64), R)
    batch.clone_gather(src, B)
    ids = (np.arange(B * R, dtype=np.uint64) * np.uint64(7919)) + np.uint64(5)
    budget = 25
    rets, lens, trace = batch.playout(B, B * R, max_plies=budget, stream_ids=ids, trace=True)
    recs = batch.export_state(B, B * R)
    leaf_recs = batch.export_state(0, B)
    for i in range(B):
        assert np.array_equal(leaf_recs[i], leaves[i].export_record())  # sources untouched
    for j in range(B * R):
        st = leaves[j // R].clone()
        oa = st.playout_philox(SEED, int(ids[j]), budget)
        assert lens[j] == len(oa)
        assert trace[:lens[j], j].tolist() == oa
        assert np.array_equal(recs[j], st.export_record())
        assert rets[j].tolist() == st.returns()
    batch.close()


def test_reset_clone_import_roundtrip(oracle_mod):
    from twixt_for_open_spiel_b200 import TwixTBatch
    n = 9
    batch = TwixTBatch(n, 64, 0, SEED)
    batch.playout(0, 32, max_plies=17)
    a = batch.export_state(0, 32)
    batch.clone(0, 32, 32)
    assert np.array_equal(batch.export_state(32, 32), a)
    batch.reset(0, 32)
    init = oracle_mod.OracleGame(n).new_initial_state().export_record()
    assert all(np.array_equal(r, init) for r in batch.export_state(0, 32))
    batch.import_state(a, 0)
    assert np.array_equal(batch.export_state(0, 32), a)
    other = TwixTBatch(n, 8, 0, 1)
    other.clone_from(0, batch, 5, 8)
    assert np.array_equal(other.export_state(), a[5:13])
    with pytest.raises(ValueError):
        batch.clone(0, 10, 20)  # overlap
    with pytest.raises(ValueError):
        batch.reset(60, 10)  # out of range
    other.close()
    batch.close()


def test_spiel_adapter_reference_kats():
    """The reference's own tests (twixt_test.cc:108-199) re-hosted on the adapter."""
    from twixt_for_open_spiel_b200 import SpielFatalError, load_game
    game = load_game("twixt")
    st = game.new_initial_state()
    assert st.current_player() == 0 and 11 in st.legal_actions()
    st.apply_action(19)
    assert st.current_player() == 1
    st.apply_action(19)  # swap
    assert 19 in st.legal_actions() and 29 not in st.legal_actions()
    assert st.current_player() == 0
    st.apply_action(36)
    la = st.legal_actions()
    assert 19 in la and 29 not in la and 36 not in la

    st = game.new_initial_state()
    assert not st.is_terminal() and len(st.legal_actions()) == 48
    sizes = []
    for a in [21, 38, 15, 11]:
        st.apply_action(a)
        sizes.append(len(st.legal_actions()))
    assert sizes == [48, 46, 46, 44]
    with pytest.raises(SpielFatalError, match=r"^Not a legal action: 11$"):
        st.apply_action(11)
    for a, want in [(27, 44), (17, 42), (42, 42), (45, 40)]:
        st.apply_action(a)
        assert len(st.legal_actions()) == want
    c = st.clone()
    st.apply_action(48)
    assert st.is_terminal() and st.player_return(0) == 1.0 and st.player_return(1) == -1.0
    assert st.current_player() == -4 and st.legal_actions() == []
    assert not c.is_terminal() and len(c.legal_actions()) == 40  # clone is independent

    g5 = load_game("twixt(board_size=5)")
    st = g5.new_initial_state()
    while not st.is_terminal():
        st.apply_action(st.legal_actions()[0])
        st.apply_action(st.legal_actions()[1])
    assert st.returns() == [0.0, 0.0]
    assert st.history() == [5, 2, 6, 3, 7, 8, 9, 11, 10, 12, 13, 16, 14, 17, 15, 18, 19, 21]


def test_cuda_matches_reference_generated_fixture():
    """tests/golden/ref_games.json (outputs of the unmodified reference) replayed through the C ABI."""
    import json
    import os
    import zlib
    from twixt_for_open_spiel_b200 import TwixTBatch
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_games.json")) as f:
        games = json.load(f)["games"]
    by_n = {}
    for g in games:
        by_n.setdefault(g["n"], []).append(g)
    for n, gs in by_n.items():
        batch = TwixTBatch(n, len(gs), 0, SEED)
        acts = pad_games([g["actions"] for g in gs])
        for ply in range(acts.shape[1] + 1):
            la, cnt = batch.legal_actions()
            player = batch.current_player()
            obs = batch.observation()
            for e, g in enumerate(gs):
                if ply >= len(g["plies"]):
                    continue
                p, count, crc_l, crc_o = g["plies"][ply]
                assert int(player[e]) == p and int(cnt[e]) == count, (n, e, ply)
                assert (zlib.crc32(la[e, :count].astype(np.int64).tobytes()) & 0xFFFFFFFF) == crc_l
                assert (zlib.crc32(obs[e].tobytes()) & 0xFFFFFFFF) == crc_o, (n, e, ply)
            if ply < acts.shape[1]:
                batch.apply(acts[:, ply].copy())
        rets, term = batch.returns(), batch.is_terminal()
        for e, g in enumerate(gs):
            assert bool(term[e]) == g["terminal"] and rets[e].tolist() == g["returns"]
        batch.close()


def test_cuda_golden_playthrough():
    """The reference's playthrough.txt (committed as tests/golden/playthrough_n8.json) through the adapter."""
    import json
    import os
    from twixt_for_open_spiel_b200 import load_game
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "playthrough_n8.json")) as f:
        pt = json.load(f)
    game = load_game("twixt")
    assert game.num_distinct_actions() == 64 and game.max_game_length() == 61
    assert game.observation_tensor_shape() == [12, 8, 6] and game.observation_tensor_size() == 576
    st = game.new_initial_state()
    by_index = {s["index"]: s for s in pt["states"]}
    for ply in range(36):
        s = by_index.get(ply)
        if s is not None and "current_player" in s:
            assert st.current_player() == s["current_player"] and st.is_terminal() == s["is_terminal"]
            assert st.returns() == s["returns"] and st.history() == s["history"]
            if "legal_actions" in s:
                assert st.legal_actions() == s["legal_actions"]
                assert [st.action_to_string(st.current_player(), a) for a in st.legal_actions()] == \
                    s["string_legal_actions"]
            for p in (0, 1):
                if "obs_ones_%d" % p in s:
                    assert np.flatnonzero(np.asarray(st.observation_tensor(p))).tolist() == s["obs_ones_%d" % p]
        if ply < 35:
            st.apply_action(pt["actions"][ply])
    assert st.is_terminal() and st.returns() == [1.0, -1.0] and st.current_player() == -4


def test_batched_rollout_evaluator_mcts_shape(oracle_mod):
    """BASELINE config C3: n=12 leaves at random plies x rollout_count=4; means equal the oracle's."""
    from twixt_for_open_spiel_b200.rollout import BatchedRolloutEvaluator
    n, B, R = 12, 256, 4
    og = oracle_mod.OracleGame(n)
    rng = random.Random(33)
    leaves = []
    for i in range(B):
        st = og.new_initial_state()
        st.replay(random_game_actions(og, rng, force_swap=(i % 5 == 0), max_plies=rng.randrange(0, 61)))
        leaves.append(st)
    ev = BatchedRolloutEvaluator(n, n_rollouts=R, max_leaves=300, seed=SEED)
    recs = np.stack([st.export_record() for st in leaves])
    for call in range(2):  # the second call draws fresh streams
        ids = ev.stream_ids(B)
        got = ev.evaluate_records(recs)
        want = np.zeros((B, 2), dtype=np.float64)
        for i, leaf in enumerate(leaves):
            for r in range(R):
                st = leaf.clone()
                st.playout_philox(SEED, int(ids[i * R + r]))
                want[i] += st.returns()
        assert np.array_equal(got, (want / R).astype(np.float32)), call
    assert np.array_equal(ev.batch.export_state(0, B), recs)  # the leaves themselves were not advanced
    ev.close()


def test_full_size_playout_properties(oracle_mod):
    """BASELINE config C4 at full size (n=24, 1 Mi envs): size-independent properties plus an oracle-replayed
    sample spread over the whole range, determinism, and independence of how the range is sharded."""
    from twixt_for_open_spiel_b200 import TwixTBatch
    n, E = 24, 1 << 20
    batch = TwixTBatch(n, E, 0, SEED)
    rets, lens, _ = batch.playout()
    st = batch.stats()
    assert st["games"] == E and st["plies"] == int(lens.astype(np.int64).sum())
    assert bool(batch.is_terminal().all()) and (batch.current_player() == -4).all()
    assert int(lens.max()) == st["max_length"] <= n * n - 3 and int(lens.min()) >= 2 * (n - 1) // 2
    red, blue = int((rets[:, 0] > 0).sum()), int((rets[:, 1] > 0).sum())
    assert (red, blue, E - red - blue) == (st["red_wins"], st["blue_wins"], st["draws"])
    assert np.array_equal(rets[:, 0], -rets[:, 1])  # zero-sum, twixt.h:128
    assert 0.07 < red / E < 0.15 and 0.07 < blue / E < 0.15  # ~0.10 / 0.13 / 0.77
    cnt = batch.legal_actions(0, 4096)[1]
    assert (cnt == 0).all()  # LegalActions() of a terminal state is empty (twixt.h:87-88)
    # an oracle-replayed sample across the range (first, last, and a stride in between)
    og = oracle_mod.OracleGame(n)
    sample = sorted(set([0, 1, 31, 32, 127, 128, E - 1, E - 129] + list(range(7, E, E // 300))))
    recs = {e: batch.export_state(e, 1)[0] for e in sample}
    for e in sample:
        s = og.new_initial_state()
        acts = s.playout_philox(SEED, e)
        assert len(acts) == lens[e] and s.returns() == rets[e].tolist(), e
        assert np.array_equal(recs[e], s.export_record()), e
    # determinism: the same seed replays the same games
    batch.reset()
    rets2, lens2, _ = batch.playout()
    assert np.array_equal(lens, lens2) and np.array_equal(rets, rets2)
    # sharding independence: a shard playing global ids [a, a+m) alone gives the same games
    a, m = 3 * (E // 8) + 5, 10_000
    shard = TwixTBatch(n, m, 0, SEED)
    shard.set_stream_base(a)
    srets, slens, _ = shard.playout()
    assert np.array_equal(slens, lens[a:a + m]) and np.array_equal(srets, rets[a:a + m])
    assert np.array_equal(shard.export_state(0, 64), batch.export_state(a, 64))
    shard.close()
    batch.close()


def test_playout_edge_cases(oracle_mod):
    """Ragged ranges, zero budgets, already-terminal envs, truncated traces, device-side stream ids."""
    import torch
    from twixt_for_open_spiel_b200 import TwixTBatch
    n, E = 7, 333  # not a multiple of the block or warp size
    og = oracle_mod.OracleGame(n)
    batch = TwixTBatch(n, E, 0, SEED)
    # max_plies = 0: nothing moves
    init = batch.export_state()
    rets, lens, _ = batch.playout(max_plies=0)
    assert (lens == 0).all() and (rets == 0).all() and np.array_equal(batch.export_state(), init)
    assert batch.stats()["plies"] == 0 and batch.stats()["games"] == 0
    # a sub-range in the middle, budget 5, trace shorter than the budget
    rets, lens, trace = batch.playout(100, 37, max_plies=5, out_actions=np.zeros((3, 37), dtype=np.uint16))
    assert (lens == 5).all() and trace.shape == (3, 37)
    after = batch.export_state()
    assert np.array_equal(after[:100], init[:100]) and np.array_equal(after[137:], init[137:])
    for j in range(37):
        st = og.new_initial_state()
        acts = st.playout_philox(SEED, 100 + j, 5)
        assert trace[:, j].tolist() == acts[:3]
        assert np.array_equal(after[100 + j], st.export_record())
    # finish everything; envs 100..136 continue from ply 5 with fresh step counters
    ids = torch.arange(1000, 1000 + E, dtype=torch.int64, device="cuda:0")
    rets_d = torch.zeros((E, 2), dtype=torch.float32, device="cuda:0")
    lens_d = torch.zeros(E, dtype=torch.int32, device="cuda:0")
    batch.playout(stream_ids=ids, out_returns=rets_d, out_lengths=lens_d)
    batch.synchronize()
    final = batch.export_state()
    for e in (0, 99, 100, 136, 137, 332):
        st = og.new_initial_state()
        if 100 <= e < 137:
            st.playout_philox(SEED, e, 5)
        acts = st.playout_philox(SEED, 1000 + e)
        assert int(lens_d[e]) == len(acts) and np.array_equal(final[e], st.export_record()), e
        assert rets_d[e].tolist() == st.returns()
    # everything is terminal now: another playout is a no-op and counts no games
    before = batch.stats()
    rets3, lens3, _ = batch.playout()
    assert (lens3 == 0).all() and np.array_equal(rets3, rets_d.cpu().numpy())
    assert batch.stats()["games"] == before["games"] and np.array_equal(batch.export_state(), final)
    # int32 legal lists with a wide stride; untouched tail keeps its fill value
    batch.reset(0, 10)
    la = np.full((10, 60), -7, dtype=np.int32)
    cnt = np.zeros(10, dtype=np.int32)
    batch.legal_actions(0, 10, out_actions=la, out_counts=cnt)
    assert (cnt == n * (n - 2)).all() and (la[:, n * (n - 2):] == -7).all()
    assert la[0, :n * (n - 2)].tolist() == og.new_initial_state().legal_actions()
    with pytest.raises(ValueError):
        batch.legal_actions(0, 10, out_actions=np.zeros((10, 5), dtype=np.int64))  # stride too small
    batch.close()


@pytest.mark.parametrize("n", [5, 7, 8, 13, 18, 24])
def test_legal_list_every_width_stride_and_alignment(oracle_mod, n):
    """K1 list kernel: uint16 / int32 / int64 actions, strides that do and do not keep rows 16-byte aligned,
    a base pointer off the 16-byte grid; lists equal the oracle's, entries past the count keep their fill."""
    import torch
    from twixt_for_open_spiel_b200 import TwixTBatch
    og = oracle_mod.OracleGame(n)
    E = 97
    batch = TwixTBatch(n, E, 0, SEED)
    # envs at different depths: env e plays e % 33 plies in one call (some games at n=5 are over by then)
    depth_of = [e % 33 for e in range(E)]
    for e in range(E):
        if depth_of[e]:
            batch.playout(e, 1, max_plies=depth_of[e], want_returns=False, want_lengths=False)
    recs = batch.export_state()
    want = []
    for e in range(E):
        st = og.new_initial_state()
        st.playout_philox(SEED, e, depth_of[e])
        assert np.array_equal(st.export_record(), recs[e])
        want.append(st.legal_actions())
    m = batch.max_legal_actions
    dev = torch.device("cuda", 0)
    for dtype, fill in ((torch.int16, -7), (torch.int32, -7), (torch.int64, -7)):
        per_vec = 16 // torch.empty(0, dtype=dtype).element_size()
        for stride in (m, m + 1, -(-m // per_vec) * per_vec, -(-m // per_vec) * per_vec + per_vec):
            for off in (0, 1):
                flat = torch.full((E * stride + 8,), fill, dtype=dtype, device=dev)
                out = flat[off:off + E * stride].view(E, stride)
                cnt = torch.zeros(E, dtype=torch.int32, device=dev)
                batch.legal_actions(out_actions=out, out_counts=cnt)
                got, c = out.cpu().numpy(), cnt.cpu().numpy()
                for e in range(E):
                    assert c[e] == len(want[e]), (n, dtype, stride, off, e)
                    assert got[e, :c[e]].tolist() == want[e], (n, dtype, stride, off, e)
                    assert (got[e, c[e]:] == fill).all(), (n, dtype, stride, off, e)
                edge = flat.cpu().numpy()
                assert (edge[:off] == fill).all() and (edge[off + E * stride:] == fill).all()
    batch.close()


def test_serialize_roundtrip_and_record_import(oracle_mod):
    """history (de)serialisation and packed-record import give the same state."""
    from twixt_for_open_spiel_b200 import load_game
    game = load_game("twixt(board_size=9)")
    og = oracle_mod.OracleGame(9)
    rng = random.Random(9)
    acts = random_game_actions(og, rng, force_swap=True, max_plies=40)
    st = game.new_initial_state()
    for a in acts:
        st.apply_action(a)
    again = game.deserialize_state(st.serialize())
    from_rec = game.new_state_from_record(st.export_record())
    ref = og.new_initial_state()
    ref.replay(acts)
    for s in (st, again, from_rec):
        assert np.array_equal(s.export_record(), ref.export_record())
        assert s.legal_actions() == ref.legal_actions() and s.current_player() == ref.current_player()
    assert again.history() == acts and st.to_string() == again.to_string()
    assert "[swapped]" in st.to_string()


def test_playout_flood_stack_overflow_path_on_device(oracle_mod, tmp_path):
    """The flood stack of the playout kernel overflows in ~1 of 10^4 floods (24 entries); a variant build with
    a 4-entry stack makes the closure fallback run constantly, and the games must still match the oracle."""
    import subprocess
    import sys
    import textwrap
    from twixt_for_open_spiel_b200 import build
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    variant = build.build(out=os.path.join(root, "tests", "_build", "libtwixt_b200_stack4.so"),
                          extra_flags=["-DTW_PLAYOUT_STACK_WORDS=4"])
    code = textwrap.dedent("""
        import sys
        import numpy as np
        sys.path.insert(0, %r)
        from oracle import pyoracle
        from twixt_for_open_spiel_b200 import TwixTBatch
        n, E, seed = 24, 1500, 0x7477697854
        b = TwixTBatch(n, E, 0, seed)
        rets, lens, trace = b.playout(trace=True)
        recs = b.export_state()
        og = pyoracle.OracleGame(n)
        for e in range(E):
            st = og.new_initial_state()
            acts = st.playout_philox(seed, e)
            assert trace[:lens[e], e].tolist() == acts, e
            assert np.array_equal(recs[e], st.export_record()), e
        print("variant ok")
    """ % root)
    env = dict(os.environ, TWIXT_B200_LIB=variant)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert res.returncode == 0 and "variant ok" in res.stdout, res.stderr[-2000:]
