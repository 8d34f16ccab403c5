# development: two-phase sequence stage with 3-byte shared-memory tables (56 lanes per SM)
ZN_SEQ=3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2 | sed "s/^/ZN_SEQ=3 /"
B="--steps 2 --warmup 1 --no-cpu --sustain 0 --no-compress"
for w in realtext realsmall; do for v in 3; do
  ZN_SEQ=$v ZN_ZPROF=1 ZN_ZPROF_SEQ1=1 python bench.py --workload $w $B 2>&1 >/dev/null | grep zpipe | tail -1 | sed "s/^/seq=$v $w: /" | cut -c1-20,100-360
done; done
