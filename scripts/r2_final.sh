# the round-end sequence the driver runs: GPU tests, smoke, the two bench arms
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err ) 2>&1 | grep real
tail -3 gpurun_out/r2_final_bench.err
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/r2_final_ref.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ref', round(r['value'],2), 'cores', r['cpu_baseline']['cores'], 'ratio e2e', round(d['e2e']['value']/r['value']))
print('frac', round(d['roofline']['frac'],3), 'matches_oracle', d['matches_oracle'], 'launches', d['gpu_launches'], 'clocks', d['clocks'])
for k in ('collision_microbench','informed_bench','class_api_bench'):
    print(k, {kk: d[k].get(kk) for kk in ('ms_per_launch','plans_per_s','s_per_plan','matches_oracle')}, d[k].get('roofline',{}).get('frac'))
print('dubins', {k:(round(v['plans_per_s']), v.get('matches_oracle')) for k,v in d['dubins_bench'].items() if isinstance(v,dict)})
print('strong', round(d['strong_scaling']['value']))
PY
