"""AlphaZero-style inference loop (BASELINE config C5 shape): B self-play games live on the GPU; every iteration
ONE kernel launch writes the [B,12,n,n-2] float32 observations and the [B,n*n] uint8 legal masks into torch
tensors, a (stand-in) policy network picks a move per env from the masked logits, and ONE launch applies all B
moves.  Nothing but the final returns crosses the bus.

    python examples/alphazero_batch.py [board_size] [batch]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from twixt_for_open_spiel_b200 import TwixTBatch  # noqa: E402
from twixt_for_open_spiel_b200.producer import ObservationMaskProducer  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    dev = torch.device("cuda:0")
    batch = TwixTBatch(n, B, 0, seed=1)
    producer = ObservationMaskProducer(batch)  # launches on torch's current stream
    # a stand-in for the policy head: one linear layer over the flattened planes
    policy = torch.nn.Linear(12 * n * (n - 2), n * n, device=dev)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    plies = 0
    with torch.no_grad():
        for _ in range(n * n):
            obs, mask = producer.produce()                       # views of the producer's device buffers
            alive = mask.bool().any(dim=1)                       # terminal envs have an all-zero mask
            if not bool(alive.any()):
                break
            logits = policy(obs.flatten(1)).masked_fill(mask == 0, float("-inf"))
            logits[~alive] = 0.0                                 # (no legal move: any value, the env is skipped)
            move = torch.distributions.Categorical(logits=logits).sample().to(torch.int32)
            move = torch.where(alive, move, torch.full_like(move, -1))  # a negative action skips the env
            batch.apply(move, out_status=status)                 # asynchronous, statuses stay on the device
            plies += int(alive.sum())
    assert int((status == 1).sum()) == 0                         # the mask never offered an illegal move
    rets = batch.returns()
    print("games %d  plies %d  red wins %d  blue wins %d  draws %d" % (
        B, plies, int((rets[:, 0] > 0).sum()), int((rets[:, 1] > 0).sum()), int((rets[:, 0] == 0).sum())))
    batch.close()


if __name__ == "__main__":
    main()
