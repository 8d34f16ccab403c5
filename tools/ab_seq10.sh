python -m pytest tests -m gpu -x -q 2>&1 | tail -2
B="--steps 5 --warmup 2 --no-cpu --sustain 0 --no-compress"
for w in realtext realsmall small100k; do
python bench.py --workload $w $B 2>gpurun_out/ab10_$w.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', d['value'], d['ms_per_step'], 'launches', d['gpu_launches'], 'fallbacks', d['pipeline_fallbacks'], d['roofline'].get('kernels'))"
done
