#!/bin/bash
# table builder + depth prep parity, then the default bench (each under its own timeout; stop at the first failure)
set -o pipefail
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bev_pool_gpu.py tests/test_depth_prep_gpu.py tests/test_static_gpu.py -x -q 2>&1 | tail -15 || exit 1
timeout 600 python bench.py > gpurun_out/bench_next.json 2> gpurun_out/bench_next.err || { tail -20 gpurun_out/bench_next.err; exit 1; }
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_next.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"])
st = d.get("stages") or d.get("config", {}).get("stages") or {}
for k in ("depth_prep", "pool_tables", "bev_pool_fused"):
    print(k, json.dumps(st.get(k)))
PY
