# development: BASELINE configs[4] at N GPUs (strong scaling): bash tools/scale_multirepo.sh N [extra bench flags]
N=$1; shift
if [ "$N" = 1 ]; then python bench.py --gpus 1 --workload multirepo --steps 5 --warmup 3 --sustain 0 "$@" > gpurun_out/mr64_n$N.json 2> gpurun_out/mr64_n$N.err
else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload multirepo --steps 5 --warmup 3 --sustain 0 --no-cpu "$@" > gpurun_out/mr64_n$N.json 2> gpurun_out/mr64_n$N.err; fi
python -c "
import json
d=json.loads(open('gpurun_out/mr64_n$N.json').read().strip().splitlines()[-1])
print('N=$N', 'device', d['value'], 'GB/s', d['ms_per_step'], 'ms | e2e verify', d['e2e']['value'], d['e2e']['ms_per_step'], '| extract', d.get('e2e_extract',{}).get('value'), '| cpu', d.get('cpu_baseline'))"
tail -2 gpurun_out/mr64_n$N.err
