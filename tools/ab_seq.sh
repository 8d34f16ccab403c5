for v in cg ldg; do for w in realtext realsmall; do
  if [ $v = ldg ]; then export ZN_SEQ_LDG=1; else unset ZN_SEQ_LDG; fi
  ZN_ZPROF=1 python bench.py --workload $w --steps 2 --no-cpu --sustain 0 --no-compress 2>&1 >/dev/null | tail -1 | sed "s/^/$v $w: /" | cut -c1-14,110-330
done; done
