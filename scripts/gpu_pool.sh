mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bev_pool_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/t_pool.log 2>&1; echo "pool rc=$?"
tail -5 gpurun_out/t_pool.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_bf16.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_bf16.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value']); print(json.dumps(d['stages']['bev_pool_fused'],indent=1))"
python scripts/profile_small.py > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 200 --csv --log-file gpurun_out/launches_small.csv python scripts/profile_small.py > gpurun_out/ncu_small.log 2>&1; echo "ncu rc=$?"
