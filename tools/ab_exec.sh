# development: A/B of the exec kernel variants (ZN_EXEC1 = wavefront version; ZN_EXEC2 = shape of the pointer-jumping one)
run() { # name, env...
  name=$1; shift
  for w in realtext realsmall; do
    env "$@" ZN_ZPROF=1 python bench.py --workload $w --steps 2 --no-cpu --sustain 0 --no-compress 2>&1 >/dev/null | tail -1 | sed "s/^/$name $w: /" | cut -c1-20,150-330
  done
}
for s in ${SHAPES:-10242 10243 5124 5123}; do run pj$s ZN_EXEC2=$s; done
