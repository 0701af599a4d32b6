for lib in "" "/root/repo/tests/_build/lib_prev.so" "" "/root/repo/tests/_build/lib_prev.so"; do
  TWIXT_B200_LIB="$lib" python bench.py --steps 5 --warmup 3 --no-cpu --no-kernels 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('lib=[$lib]', '%.4e' % d['value'], d['roofline']['kernel_ms'])"
done
for lib in "" "/root/repo/tests/_build/lib_l2.so" "" "/root/repo/tests/_build/lib_l2.so"; do
  TWIXT_B200_LIB="$lib" python bench.py --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('lib=[$lib]', {k:(round(v['ms'],4),round(v['frac'],3)) for k,v in d['kernels'].items() if 'legal' in k})"
done
