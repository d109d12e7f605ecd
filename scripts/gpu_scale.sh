#!/bin/bash
# multi-GPU: default bench + compact e2e + train config at N = $1 GPUs
N=$1
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:3}" > gpurun_out/r2s_$2_n$N.json 2> gpurun_out/r2s_$2_n$N.err; echo "$2 rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/r2s_$2_n$N.err | tail -3; }
run 29601 default --steps 20 --warmup 5
run 29602 compact --steps 20 --warmup 5 --compact-output --no-fp32 --no-train-stage
run 29603 train --config train --steps 8 --warmup 3
python - <<PY
import json
for n in ("default","compact","train"):
    try:
        d=json.loads(open("gpurun_out/r2s_%s_n$N.json"%n).read().strip().splitlines()[-1])
        t=d["stages"].get("training") or {}
        print(n,"N=$N value",round(d["value"],1),"ms/step",round(d["ms_per_step"],3),"e2e",d["e2e"]["value"] and round(d["e2e"]["value"],1),"hostlink",d["e2e"].get("hostlink",{}).get("gbs_all_ranks"),"train",t.get("frames_per_s"),t.get("ms_per_step"),t.get("ms_per_step_without_allreduce"))
    except Exception as e: print(n,"parse failed",e)
PY
