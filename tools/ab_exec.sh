# development: A/B of the exec kernel variants (ZN_EXEC1 = wavefront version; ZN_EXEC2 = shape of the pointer-jumping one)
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run() { # name, env...
  name=$1; shift
  for w in realtext realsmall; do
    env "$@" ZN_ZPROF=1 python bench.py --workload $w --steps 2 --no-cpu --sustain 0 --no-compress 2>&1 >/dev/null | tail -1 | sed "s/^/$name $w: /" | cut -c1-330
  done
}
# run wavefront ZN_EXEC1=1
for s in 10242 5122 5123 1286 1284; do run pj$s ZN_EXEC2=$s; done
