# The round's bench lines (run under gpurun, one GPU): default metric line + reference arm as the driver runs them, the
# secondary workloads, the archive pipeline table and the compress sweep.
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_text2g.json 2> gpurun_out/r2_bench_text2g.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_text2g_reference.json 2>> gpurun_out/r2_bench_text2g.err
for w in small100k mixed realtext realsmall; do
  python bench.py --workload $w --steps 20 --warmup 5 --no-compress > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err
done
python bench.py --impl reference --workload realtext --steps 5 --warmup 2 > gpurun_out/r2_bench_realtext_reference.json 2>/dev/null
python tools/bench_archive.py > gpurun_out/r2_archive_pipeline.jsonl 2> gpurun_out/r2_archive_pipeline.err
python tools/bench_compress.py > gpurun_out/r2_compress_sweep.jsonl 2> gpurun_out/r2_compress_sweep.err
python - <<'PY'
import json
for f in ("text2g","text2g_reference","small100k","mixed","realtext","realsmall","realtext_reference"):
    try:
        d=json.loads(open(f"gpurun_out/r2_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "x", d.get("e2e_extract",{}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "frac", d.get("roofline",{}).get("frac"), "alu", d.get("roofline",{}).get("alu",{}).get("frac"), "sus", d.get("sustained",{}).get("value"), (d.get("sustained",{}).get("clocks") or {}).get("sm_mhz"))
    except Exception as e: print(f, "ERR", e)
PY
tail -6 gpurun_out/r2_archive_pipeline.jsonl | cut -c1-330
tail -3 gpurun_out/r2_archive_pipeline.err
cat gpurun_out/r2_compress_sweep.jsonl | cut -c1-300
