#!/bin/bash
# round 2, call B: new bench (default line + train config), N=1
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2b_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1])
    print('value',d['value'],'e2e',d['e2e']['value'],'repeats',d['repeats'],'parity',d['parity'])
    print('fp32',d['fp32'])
    print('train',json.dumps(d['stages'].get('training'))[:1500])
    print('roof',d['roofline']['frac'])
    bf=d['stages'].get('boundary_forms',{})
    print({k:(v.get('ms'),v.get('reference_gpu_ms')) for k,v in bf.items() if isinstance(v,dict)})
    print('vox',d['stages']['voxelize_mean']['ms'],d['stages']['voxelize_mean']['frac'],'pool',d['stages']['bev_pool_fused']['ms'],d['stages']['bev_pool_fused']['frac'], d['stages'].get('bev_pool_fused_batch4'))
    print('hostlink', d['e2e']['hostlink'])
except Exception as e: print('parse failed',e)
PY
echo skip-train

