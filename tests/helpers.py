"""Shared helpers of the test-suite (test infrastructure)."""
import random
from typing import List

import numpy as np

SEED = 0x7477697854  # "twixT"


def random_game_actions(og, rng: random.Random, force_swap: bool = False, max_plies: int = 10**9) -> List[int]:
    """A random legal game on the oracle; optionally answers the first move with the swap."""
    st = og.new_initial_state()
    acts: List[int] = []
    while not st.is_terminal() and len(acts) < max_plies:
        la = st.legal_actions()
        if len(acts) == 1 and force_swap and acts[0] in la:
            a = acts[0]
        else:
            a = rng.choice(la)
        st.apply_action(a)
        acts.append(a)
    return acts


def draw_seeking_actions(og, pattern=(0, 1)) -> List[int]:
    """twixt_test.cc:185-199: alternately LegalActions().at(p0) / .at(p1)."""
    st = og.new_initial_state()
    acts: List[int] = []
    i = 0
    while not st.is_terminal():
        la = st.legal_actions()
        a = la[min(pattern[i % len(pattern)], len(la) - 1)]
        st.apply_action(a)
        acts.append(a)
        i += 1
    return acts


def pad_games(games: List[List[int]], fill: int = -1) -> np.ndarray:
    """[num_games, max_len] int32 action matrix, `fill` past the end of a game."""
    m = max(len(g) for g in games)
    out = np.full((len(games), m), fill, dtype=np.int32)
    for i, g in enumerate(games):
        out[i, :len(g)] = g
    return out
