#!/bin/bash
# multi-GPU (round 2, final code): default bench (dense fp32 output; e2e_rows block inside) + e2e with the lossless rows
# output + train config at N = $1 GPUs; tag = $2
N=$1; tag=${2:-r2x}
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:3}" > gpurun_out/${tag}_$2_n$N.json 2> gpurun_out/${tag}_$2_n$N.err; echo "$2 rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/${tag}_$2_n$N.err | tail -3; }
run 29611 default --steps 20 --warmup 5 --no-cpu-baseline --no-fp32 --no-train-stage
run 29612 rows --steps 20 --warmup 5 --output rows --no-cpu-baseline --no-fp32 --no-train-stage
run 29613 train --config train --steps 8 --warmup 3
python - <<PY
import json
for n in ("default","rows","train"):
    try:
        d=json.loads(open("gpurun_out/${tag}_%s_n$N.json"%n).read().strip().splitlines()[-1])
        t=d["stages"].get("training") or {}
        r=d.get("e2e_rows") or {}
        print(n,"N=$N value",round(d["value"],1),"ms/step",round(d["ms_per_step"],3),"e2e",d["e2e"]["value"] and round(d["e2e"]["value"],1),"d2h",d["e2e"].get("d2h_bytes_per_step"),"hostlink",d["e2e"].get("hostlink",{}).get("gbs_all_ranks"),"e2e_rows",r.get("value"),"train",t.get("frames_per_s"),t.get("ms_per_step"),t.get("ms_per_step_without_allreduce"))
    except Exception as e: print(n,"parse failed",e)
PY
