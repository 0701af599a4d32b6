"""World-size-2 gloo test of the multi-GPU host logic (no GPU needed): shards of the global env-id range
play the Philox policy independently (here on the oracle) and only a handful of counters are all-reduced.
The reduced counters must equal a single process playing the whole range -- results do not depend on the
number of ranks."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED = 0x7477697854
N, GLOBAL_ENVS = 6, 101  # odd on purpose: ragged shards


def _play_range(first, count):
    sys.path.insert(0, ROOT)
    from oracle import pyoracle
    og = pyoracle.OracleGame(N)
    c = {"plies": 0, "games": 0, "red_wins": 0, "blue_wins": 0, "draws": 0, "swaps": 0, "max_length": 0,
         "kernel_launches": 0}
    for gid in range(first, first + count):
        st = og.new_initial_state()
        acts = st.playout_philox(SEED, gid)
        r = st.returns()
        c["plies"] += len(acts)
        c["games"] += 1
        c["red_wins"] += r[0] > 0
        c["blue_wins"] += r[1] > 0
        c["draws"] += r[0] == 0
        c["swaps"] += st.board_header()[1]
        c["max_length"] = max(c["max_length"], len(acts))
    return c


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from twixt_for_open_spiel_b200.sharding import reduce_counters, shard_range
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    first, count = shard_range(GLOBAL_ENVS, world, rank)
    local = _play_range(first, count)
    total = reduce_counters(local, dist)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, first, count, total))


def test_shard_ranges_partition_the_env_ids():
    from twixt_for_open_spiel_b200.sharding import shard_range
    for world in (1, 2, 3, 4, 8):
        for total in (0, 1, 7, 8, 101, 1 << 20):
            nxt = 0
            for r in range(world):
                first, count = shard_range(total, world, r)
                assert first == nxt and count in (total // world, total // world + 1)
                nxt = first + count
            assert nxt == total
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_two_rank_reduction_equals_single_process():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = _play_range(0, GLOBAL_ENVS)
    results.sort()
    assert results[0][1:3] == (0, 51) and results[1][1:3] == (51, 50)
    for _, _, _, total in results:
        assert total == {k: int(v) for k, v in single.items()}
