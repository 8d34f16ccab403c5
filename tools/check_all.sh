# development: the GPU tests, then one line per bench workload (value, ms/step, e2e, per-kernel ms)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in text2g small100k mixed realtext realsmall; do python bench.py --workload $w --steps 10 --no-cpu --sustain 0 --no-compress 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'x', d.get('e2e_extract',{}).get('value'), {k:v['ms'] for k,v in d['roofline']['kernels'].items()})"; done
python bench.py --workload multirepo --gib 8 --steps 5 --no-cpu --sustain 0 --extract 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('multirepo8', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'x', d.get('e2e_extract',{}).get('value'))"
