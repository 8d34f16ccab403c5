for x in 1 0; do for w in small100k realsmall text2g; do ZN_XXH=$x python bench.py --workload $w --steps 5 --no-cpu --sustain 0 --no-compress 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('xxh=$x', '$w', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], {k:v['ms'] for k,v in d['roofline']['kernels'].items()})"; done; done
