#!/usr/bin/env python3
"""The reference README's first command, `example --game=twixt`, two ways on the GPU engine:
(1) one game through the pyspiel-shaped adapter, move by move; (2) a million games in one kernel launch."""
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from twixt_for_open_spiel_b200 import TwixTBatch, load_game  # noqa: E402

game = load_game("twixt(board_size=8,ansi_color_output=False)")
state = game.new_initial_state()
rng = random.Random(0)
while not state.is_terminal():
    action = rng.choice(state.legal_actions())
    print("player %d plays %s" % (state.current_player(), state.action_to_string(state.current_player(), action)))
    state.apply_action(action)
print(state)
print("returns:", state.returns())

batch = TwixTBatch(board_size=24, num_envs=1 << 20, device=0, seed=1)
t0 = time.time()
returns, lengths, _ = batch.playout()
dt = time.time() - t0
print("%d games, %d plies in %.1f ms (incl. copying results to the host): %.2e steps/s; red/blue/draw = %s"
      % (len(lengths), lengths.sum(), dt * 1e3, lengths.sum() / dt,
         [(returns[:, 0] > 0).mean(), (returns[:, 1] > 0).mean(), (returns[:, 0] == 0).mean()]))
