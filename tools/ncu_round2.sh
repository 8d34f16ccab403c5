# The round's ncu evidence (run under gpurun, one GPU; every profiled command has first exited 0 without ncu):
#   launch list of the default bench, then one `--set full` capture per kernel family
set -x
B="--steps 3 --warmup 3 --no-cpu --sustain 0"
python bench.py $B > gpurun_out/r2_plain.json 2> gpurun_out/r2_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_steps3.csv python bench.py $B > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_decode_ws -c 1 -f -o gpurun_out/r2_ws python bench.py $B --no-compress > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_zstd_blocks -c 1 -f -o gpurun_out/r2_zstd_blocks python bench.py $B > gpurun_out/ncu_d.log 2>&1
python bench.py --workload realtext $B --no-compress > gpurun_out/r2_plain_rt.json 2> gpurun_out/r2_plain_rt.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_zexec2|k_zseq_g|k_zlit|k_ztables" -c 4 -f -o gpurun_out/r2_zpipe python bench.py --workload realtext $B --no-compress > gpurun_out/ncu_c.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_realtext_steps3.csv python bench.py --workload realtext $B --no-compress > gpurun_out/ncu_e.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
