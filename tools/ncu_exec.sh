# development: one ncu capture of the exec kernel on the real-text workload (source-level counters)
ZN_EXEC2=${1:-5122} python bench.py --workload realtext --steps 1 --warmup 1 --no-cpu --sustain 0 --no-compress > /dev/null 2>&1 || exit 1
ZN_EXEC2=${1:-5122} ncu --set full --clock-control none --import-source on -k regex:k_zexec2 -c 1 -o gpurun_out/r2_zexec2 -f \
  python bench.py --workload realtext --steps 1 --warmup 1 --no-cpu --sustain 0 --no-compress > gpurun_out/ncu_exec.log 2>&1
tail -3 gpurun_out/ncu_exec.log
for s in 5122 10242 5123 1286; do for w in realtext realsmall; do ZN_EXEC2=$s ZN_ZPROF=1 python bench.py --workload $w --steps 2 --no-cpu --sustain 0 --no-compress 2>&1 >/dev/null | tail -1 | sed "s/^/pj$s $w: /" | cut -c150-330; done; done
