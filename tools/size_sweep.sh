#!/bin/bash
# playout throughput per board size (compile-time-size instantiations: 8, 12, 24; the others run the run-time-size kernel)
for n in 5 8 10 12 16 20 23 24; do
  python bench.py --board-size $n --steps 5 --warmup 3 --no-cpu --no-kernels 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('n=$n', '%.4e steps/s' % d['value'], '%.3f ms' % d['roofline']['kernel_ms'], 'mean plies/game %.1f' % (d['outcomes']['plies']/max(d['outcomes']['games'],1)))"
done
