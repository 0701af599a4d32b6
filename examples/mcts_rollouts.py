#!/usr/bin/env python3
"""`mcts_example --game="twixt(board_size=12)" --rollout_count=4` shape: the tree stays on the host, the
leaf evaluation (4 random rollouts per leaf) is batched on the GPU -- here for all children of the root."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from twixt_for_open_spiel_b200 import load_game  # noqa: E402
from twixt_for_open_spiel_b200.rollout import BatchedRolloutEvaluator  # noqa: E402

game = load_game("twixt(board_size=12)")
root = game.new_initial_state()
for a in (40, 77, 65):
    root.apply_action(a)
children = []
for a in root.legal_actions():
    c = root.clone()
    c.apply_action(a)
    children.append((a, c))
ev = BatchedRolloutEvaluator(12, n_rollouts=4, max_leaves=256, seed=7)
values = ev.evaluate([c for _, c in children])  # [num_children, 2] mean returns
mover = root.current_player()
best = int(np.argmax(values[:, mover]))
print("root to move: player %d, %d children evaluated with 4 rollouts each in one launch" % (mover, len(children)))
print("best child by rollout value: %s (%.2f)" % (root.action_to_string(mover, children[best][0]), values[best, mover]))
