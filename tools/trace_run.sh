# development: phase counters of the exec kernel (trace build), then the normal build again
ZN_TRACE_BUILD=1 python -m znippy_b200.build --force > gpurun_out/tracebuild.log 2>&1
for w in realtext realsmall; do ZN_ZPROF=1 python bench.py --workload $w --steps 1 --no-cpu --sustain 0 --no-compress 2>&1 >/dev/null | tail -2; done
python -m znippy_b200.build --force > gpurun_out/tracebuild2.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
ZN_HOST_PROF=1 python bench.py --workload small100k --steps 3 --no-cpu --sustain 0 --no-compress 2>&1 >/dev/null | tail -3
