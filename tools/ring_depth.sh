for rep in 1 2; do
for lib in "" _nb12 _nb16; do
  L=$PWD/sayuri_b200/libsayuri_b200$lib.so
  for net in 10bx128 20bx256; do
    SAYURI_B200_LIB=$L python bench.py --steps 20 --warmup 5 --no-cpu-baseline --eval-threads 0 --net $net 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lib$lib $net value %.0f frac %.4f' % (d['value'], d['roofline']['frac']))"
  done
done
done
