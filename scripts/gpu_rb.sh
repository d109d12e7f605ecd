mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spconv_gpu.py tests/test_static_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/t_rb.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t_rb.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_graph.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_graph.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step']); print(d['stages']['sparse_encoder'])" || tail -20 gpurun_out/bench_graph.log
bash scripts/gpu_ncu_static.sh
