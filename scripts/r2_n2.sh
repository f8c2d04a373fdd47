mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -5 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
print('n_gpus', d['n_gpus'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'trees', round(d['e2e']['trees_mode']['value']), d['e2e']['matches_device_arm'])
print('strong', json.dumps(d['strong_scaling']))
print('dubins', {k:(round(v['plans_per_s']) if isinstance(v,dict) and 'plans_per_s' in v else None) for k,v in d.get('dubins_bench',{}).items()})
PY
