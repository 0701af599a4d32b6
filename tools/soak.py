"""One-off soak: fused playouts vs the oracle on many more games than the test-suite, every board size."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle  # noqa: E402
from twixt_for_open_spiel_b200 import TwixTBatch  # noqa: E402

SEED = int(sys.argv[1]) if len(sys.argv) > 1 else 20261018
total = 0
t0 = time.time()
for n in range(5, 25):
    E = 20000 if n <= 12 else 6000
    og = pyoracle.OracleGame(n)
    b = TwixTBatch(n, E, 0, SEED)
    b.set_stream_base(n << 40)
    rets, lens, _ = b.playout()
    recs = b.export_state()
    obs = b.observation(0, 256)
    la, cnt = b.legal_actions(0, 256)
    bad = 0
    for e in range(E):
        st = og.new_initial_state()
        acts = st.playout_philox(SEED, (n << 40) + e)
        if len(acts) != lens[e] or st.returns() != rets[e].tolist() or not np.array_equal(recs[e], st.export_record()):
            bad += 1
        elif e < 256 and (not np.array_equal(obs[e].reshape(-1), st.observation_tensor(0)) or cnt[e] != 0):
            bad += 1
    total += int(lens.sum())
    print("n=%2d envs=%d plies=%d mismatches=%d" % (n, E, int(lens.sum()), bad), flush=True)
    assert bad == 0
    b.close()
print("soak ok: %d plies checked in %.0f s" % (total, time.time() - t0))
