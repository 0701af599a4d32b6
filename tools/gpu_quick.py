"""First-light check on the GPU box: a tiny parity probe and a rough playout timing."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle  # noqa: E402
from twixt_for_open_spiel_b200 import TwixTBatch  # noqa: E402

SEED = 0x7477697854
print(torch.cuda.get_device_name(0))
for n in (5, 8, 24):
    og = pyoracle.OracleGame(n)
    b = TwixTBatch(n, 64, 0, SEED)
    rets, lens, trace = b.playout(trace=True)
    recs = b.export_state()
    bad = 0
    for e in range(64):
        st = og.new_initial_state()
        oa = st.playout_philox(SEED, e)
        if trace[:lens[e], e].tolist() != oa or not np.array_equal(recs[e], st.export_record()):
            bad += 1
    print("n", n, "mismatching envs", bad, "mean len", lens.mean())
    b.close()

for n, E in ((24, 1 << 20), (24, 1 << 18), (8, 1 << 20), (12, 1 << 20)):
    b = TwixTBatch(n, E, 0, SEED)
    b.use_torch_stream()
    rets = torch.zeros((E, 2), dtype=torch.float32, device="cuda")
    lens = torch.zeros(E, dtype=torch.int32, device="cuda")
    for it in range(3):
        b.reset()
        b.stats_reset()
        torch.cuda.synchronize()
        t0 = time.time()
        b.playout(out_returns=rets, out_lengths=lens)
        torch.cuda.synchronize()
        dt = time.time() - t0
        s = b.stats()
        print("n=%d E=%d: %.2f ms, %d plies, %.3e steps/s  red/blue/draw %.3f/%.3f/%.3f maxlen %d" % (
            n, E, dt * 1e3, s["plies"], s["plies"] / dt, s["red_wins"] / E, s["blue_wins"] / E, s["draws"] / E,
            s["max_length"]))
    b.close()
