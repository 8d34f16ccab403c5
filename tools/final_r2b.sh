# Round-2 closing evidence for the two-phase sequence stage (run under gpurun, one GPU; every profiled command has first
# exited 0 without ncu): bench lines, ncu --set full of the two kernels, launch list of the real-text step.
python bench.py --steps 10 --warmup 3 --no-cpu --sustain 0 --no-compress > gpurun_out/r2b_bench_text2g_quick.json 2> gpurun_out/r2b_text2g.err || exit 1
python bench.py --workload realtext --steps 20 --warmup 5 --no-compress > gpurun_out/r2b_bench_realtext.json 2> gpurun_out/r2b_realtext.err || exit 1
python bench.py --workload realsmall --steps 20 --warmup 5 --no-compress --no-cpu > gpurun_out/r2b_bench_realsmall.json 2> gpurun_out/r2b_realsmall.err || exit 1
B="--workload realtext --steps 2 --warmup 1 --no-cpu --sustain 0 --no-compress"
ncu --set full --clock-control none --import-source on -k regex:"k_zseq1|k_zseq2" -c 2 -f -o gpurun_out/r2b_zseq python bench.py $B > gpurun_out/ncu_h.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_realtext.csv python bench.py $B > gpurun_out/ncu_i.log 2>&1
for f in text2g_quick realtext realsmall; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/r2b_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "cpu", (d.get("cpu_baseline") or {}).get("value"), "launches", d["gpu_launches"], d["roofline"].get("kernels"))
PY
done
