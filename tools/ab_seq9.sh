python -m pytest tests -m gpu -x -q 2>&1 | tail -2
ZN_SEQ=3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1 | sed "s/^/ZN_SEQ=3 /"
B="--steps 3 --warmup 2 --no-cpu --sustain 0 --no-compress"
ZN_ZPROF=1 ZN_ZPROF_SEQ1=1 python bench.py --workload realtext $B 2>gpurun_out/ab9.err | tail -1 | cut -c1-400
grep zpipe gpurun_out/ab9.err | tail -1 | cut -c100-360
