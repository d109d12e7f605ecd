mkdir -p gpurun_out


timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_bf16.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_bf16.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value']); print(json.dumps(d['stages']['bev_pool_fused'],indent=1))"

