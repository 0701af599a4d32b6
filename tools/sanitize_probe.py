"""Small run of every kernel for compute-sanitizer (memcheck): odd and even board sizes, ragged counts."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from twixt_for_open_spiel_b200 import TwixTBatch  # noqa: E402

for n in (5, 7, 8, 12, 13, 24):
    E = 70
    b = TwixTBatch(n, E, 0, 99)
    b.playout(0, 33, max_plies=n * 2)
    la, cnt = b.legal_actions()
    b.legal_mask()
    b.observation()
    acts = np.where(cnt > 0, la[np.arange(E), 0], -1).astype(np.int32)
    b.apply(acts)
    b.clone(0, 35, 35)
    b.clone_gather(np.arange(10, dtype=np.int64), 50)
    b.playout(trace=True)
    b.returns(); b.current_player(); b.is_terminal()
    rec = b.export_state()
    b.import_state(rec)
    b.reset(3, 17)
    b.stats()
    b.close()
print("sanitize probe ok")
