"""Runs each per-call kernel a few times on mid-game states (n=24) -- the command profiled with ncu
(sizes as in bench.py's kernel section: 1 Mi envs for list / mask / apply / validate, 64 Ki for the tensors)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from twixt_for_open_spiel_b200 import TwixTBatch  # noqa: E402

n, E, E4 = 24, 1 << 20, 1 << 16
b = TwixTBatch(n, E, 0, 0x7477697854)
b.use_torch_stream()
b.playout(max_plies=200, want_returns=False, want_lengths=False)
dev = torch.device("cuda:0")
acts = torch.zeros((E, b.max_legal_actions), dtype=torch.int16, device=dev)
acts64 = torch.zeros((E, b.max_legal_actions), dtype=torch.int64, device=dev)
cnts = torch.zeros(E, dtype=torch.int32, device=dev)
mask = torch.zeros((E, n * n), dtype=torch.uint8, device=dev)
obs = torch.empty((E4,) + b.obs_shape, dtype=torch.float32, device=dev)
snap = torch.empty((E, b.record_words), dtype=torch.int32, device=dev)
b.export_state(out=snap)
for _ in range(3):
    b.legal_actions(out_actions=acts, out_counts=cnts)
    b.legal_actions(out_actions=acts64, out_counts=cnts)
    b.legal_mask(out=mask)
    b.observation(0, E4, out=obs)
    b.observation_and_mask(0, E4, out_obs=obs, out_mask=mask[:E4])
    b.import_state(snap)  # validate kernel + copy
idx = (torch.rand(E, device=dev) * cnts.clamp(min=1).float()).long().clamp(max=b.max_legal_actions - 1)
move = acts.gather(1, idx.view(-1, 1)).view(-1).to(torch.int32)
move = torch.where(cnts > 0, move, torch.full_like(move, -1))
status = torch.zeros(E, dtype=torch.int32, device=dev)
for _ in range(3):
    b.import_state(snap)
    b.apply(move, out_status=status)
# replay: 64 Ki fresh envs x the first 200 plies of the games the playout kernel would play
r = TwixTBatch(n, E4, 0, 0x7477697854)
r.use_torch_stream()
_, _, trace = r.playout(max_plies=200, trace=True, want_returns=False, want_lengths=False)
hist_h = trace.astype("int32").T.copy()  # [E4, 200]
hist_h[hist_h == 0xFFFF] = -1  # no move made (game over earlier): end of that history
hist = torch.from_numpy(hist_h).to(dev)
applied = torch.zeros(E4, dtype=torch.int32, device=dev)
for _ in range(2):
    r.reset()
    r.replay(hist, out_applied=applied)
torch.cuda.synchronize()
print("probe ok", int((status == 0).sum()), int((status == 1).sum()), int(applied.min()), int(applied.max()))
