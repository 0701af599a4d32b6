for lib in "" "/root/repo/tests/_build/lib_prev.so" "" "/root/repo/tests/_build/lib_prev.so"; do
  TWIXT_B200_LIB="$lib" python bench.py --steps 5 --warmup 3 --no-cpu --no-kernels 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('lib=[$lib]', '%.4e' % d['value'], d['roofline']['kernel_ms'], d['outcomes']['plies'], d['outcomes']['red'], d['outcomes']['draws'])"
done
