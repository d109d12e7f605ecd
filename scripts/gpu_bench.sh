mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_static_gpu.py tests/test_spconv_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/t_static.log 2>&1; echo "static+spconv rc=$?"; tail -15 gpurun_out/t_static.log
for mode in graph; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --mode $mode > gpurun_out/bench_$mode.log 2>&1; echo "bench $mode rc=$?"; tail -1 gpurun_out/bench_$mode.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'launches', d['gpu_launches']); print({k:(v['ms']) for k,v in d['stages'].items()}); print(d['stages']['sparse_encoder'])" || tail -20 gpurun_out/bench_$mode.log
done
