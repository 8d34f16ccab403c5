# development: GPU tests, three runs of the metric workload, one of the small-file corpus
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for i in 1 2 3; do python bench.py --steps 100 --no-cpu --sustain 0 --no-compress 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernels'], d['gpu_launches'])"; done
python bench.py --workload small100k --steps 50 --no-cpu --sustain 0 --no-compress 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernels'], d['e2e']['value'])"
