#!/bin/bash
# playout throughput per board size with the product library (every size has its own compile-time-size kernel)
for n in ${@:-5 8 10 12 14 15 16 17 18 19 20 21 22 23 24}; do
  echo -n "n=$n  "; python tools/ab_playout.py child $n 1048576 3 2>&1 | tail -1
done
