mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "all rc=$?"; tail -6 gpurun_out/t_all.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_graph.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_graph.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])" || tail -20 gpurun_out/bench_graph.log
