set -x
ZN_TRACE_BUILD=1 python -m znippy_b200.build --force > gpurun_out/tracebuild.log 2>&1
ZN_ZPROF=1 python bench.py --workload realtext --steps 2 --no-cpu --sustain 0 --no-compress > gpurun_out/trace_rt.log 2> gpurun_out/trace_rt.err
tail -4 gpurun_out/trace_rt.err
ZN_ZPROF=1 python bench.py --workload realsmall --steps 2 --no-cpu --sustain 0 --no-compress > gpurun_out/trace_rs.log 2> gpurun_out/trace_rs.err
tail -4 gpurun_out/trace_rs.err
python -m znippy_b200.build --force > gpurun_out/tracebuild2.log 2>&1
ZN_HOST_PROF=1 python tools/bench_archive.py 1.0 3 2> gpurun_out/hostprof.err | tail -2
grep -c . gpurun_out/hostprof.err; tail -12 gpurun_out/hostprof.err
