# development: far prefetch in the lane-per-stream bit readers; sequence-stage variants again
B="--steps 2 --warmup 1 --no-cpu --sustain 0 --no-compress"
for w in realtext realsmall; do for v in 0 3 4; do
  ZN_SEQ=$v ZN_ZPROF=1 ZN_ZPROF_SEQ1=1 python bench.py --workload $w $B 2>&1 >/dev/null | grep zpipe | tail -1 | sed "s/^/seq=$v $w: /" | cut -c1-20,100-360
done; done
