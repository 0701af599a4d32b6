#!/usr/bin/env python3
"""Balls-in-bins model of the shared-memory wavefronts of the K1 list kernel's expansion.

Each of the 32 lanes owns an 18-bit chunk of the flat legal string and writes its set bits, in order, into
the env's dense uint16 row at (exclusive prefix of the chunk populations) + k.  One predicated 2-byte store
instruction per chunk bit: its cost in wavefronts is the largest number of DISTINCT 32-bit words that fall
into one of the 32 banks.  The model reproduces what ncu measures (43 store wavefronts per env at the
bench's 64 % density, profiles/r1_legal_list_u16_v2_ncu_summary.txt) and shows that the conflicts come from
the ragged transpose itself: 4-byte slots or alternating fill directions do not remove them.

usage: python tools/k1_bank_conflict_model.py [trials]
"""
import sys

import numpy as np

N = 24
CHUNK = (N * N + 31) // 32


def random_legal_string(rng, density):
    board = np.zeros((N, N), bool)  # [x, y], action = x*N + y
    if rng.random() < 0.5:
        board[1:N - 1, :] = True    # the player who may not use the first/last column
    else:
        board[:, 1:N - 1] = True    # ... the first/last row
    board &= rng.random((N, N)) < density
    flat = np.concatenate([board.reshape(-1), np.zeros(32 * CHUNK - N * N, bool)])
    return flat.reshape(32, CHUNK)


def wavefronts(words):
    if not words:
        return 0
    per_bank = {}
    for w in set(words):
        per_bank.setdefault(w % 32, set()).add(w)
    return max(len(v) for v in per_bank.values())


def store_wavefronts(chunks, variant):
    counts = chunks.sum(1)
    excl = np.cumsum(counts) - counts
    total = 0
    for j in range(CHUNK):
        words = []
        for lane in range(32):
            jj = CHUNK - 1 - j if (variant == "alternate" and lane & 1) else j
            if chunks[lane, jj]:
                pos = excl[lane] + chunks[lane, :jj].sum()
                words.append(pos if variant == "u32_slots" else pos // 2)
        total += wavefronts(words)
    return total


def main():
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rng = np.random.default_rng(1)
    print("density  " + "  ".join("%10s" % v for v in ("uint16", "alternate", "u32_slots")) + "   (store wavefronts per env; 18 = conflict-free)")
    for density in (0.3, 0.64, 0.9, 1.0):
        envs = [random_legal_string(rng, density) for _ in range(trials)]
        row = [np.mean([store_wavefronts(e, v) for e in envs]) for v in ("uint16", "alternate", "u32_slots")]
        print("%7.2f  " % density + "  ".join("%10.1f" % r for r in row))


if __name__ == "__main__":
    main()
