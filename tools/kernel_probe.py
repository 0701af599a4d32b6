"""Runs each per-call kernel a few times on mid-game states (n=24) -- the command profiled with ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from twixt_for_open_spiel_b200 import TwixTBatch  # noqa: E402

n, E = 24, 1 << 19
b = TwixTBatch(n, E, 0, 0x7477697854)
b.use_torch_stream()
b.playout(max_plies=200, want_returns=False, want_lengths=False)
dev = torch.device("cuda:0")
acts = torch.zeros((E, b.max_legal_actions), dtype=torch.int16, device=dev)
cnts = torch.zeros(E, dtype=torch.int32, device=dev)
mask = torch.zeros((E, n * n), dtype=torch.uint8, device=dev)
obs = torch.empty((1 << 15,) + b.obs_shape, dtype=torch.float32, device=dev)
for _ in range(3):
    b.legal_actions(out_actions=acts, out_counts=cnts)
    b.legal_mask(out=mask)
    b.observation(0, 1 << 15, out=obs)
idx = (torch.rand(E, device=dev) * cnts.clamp(min=1).float()).long().clamp(max=b.max_legal_actions - 1)
move = acts.gather(1, idx.view(-1, 1)).view(-1).to(torch.int32)
move = torch.where(cnts > 0, move, torch.full_like(move, -1))
status = torch.zeros(E, dtype=torch.int32, device=dev)
b.apply(move, out_status=status)
torch.cuda.synchronize()
print("probe ok", int((status == 0).sum()), int((status == 1).sum()))
