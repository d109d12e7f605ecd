#!/bin/bash
# full GPU suite + smoke + default bench; tag = $1
tag=${1:-r2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${tag}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
    print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'repeats',d['repeats'])
    print('parity',d['parity']['lidar_bev']['max_rel'],d['parity']['camera_bev']['max_rel'],'fp32',d['fp32']['value'],d['fp32'].get('parity',{}).get('lidar_bev',{}).get('max_rel'))
    t=d['stages'].get('training',{}); print('train fps',t.get('frames_per_s'),'ms',t.get('ms_per_step'),'ratio',t.get('fwd_bwd_over_fwd'))
    print('roof',d['roofline']['frac'],'gemm_ms',d['stages']['sparse_encoder']['gemm_ms'])
    print('vox',d['stages']['voxelize_mean']['ms'],d['stages']['voxelize_mean']['frac'],'pool',d['stages']['bev_pool_fused']['ms'],d['stages']['bev_pool_fused']['frac'],'pool4',d['stages'].get('bev_pool_fused_batch4',{}).get('ms'),d['stages'].get('bev_pool_fused_batch4',{}).get('frac'))
except Exception as e: print('parse failed',e)
PY
