# development: sequence-stage variants of the pipeline (ZN_SEQ: 0 one-pass, tables in global memory; 3 two-phase, phase 1 with
# tables in shared memory; 4 two-phase, phase 1 with tables in global memory), stage times from ZN_ZPROF
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in realtext realsmall; do for v in 0 3 4; do
  ZN_SEQ=$v ZN_ZPROF=1 python bench.py --workload $w --steps 3 --no-cpu --sustain 0 --no-compress 2>&1 >/dev/null | grep zpipe | tail -1 | sed "s/^/seq=$v $w: /" | cut -c1-20,100-330
done; done
