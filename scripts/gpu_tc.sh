mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_spconv_gpu.py -m gpu -q -x -k "bf16 or encoder or epilogue" -p no:cacheprovider > gpurun_out/t_sp_tc.log 2>&1; rc=$?; echo "tests rc=$rc"
tail -3 gpurun_out/t_sp_tc.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 180 python -m pytest tests/test_static_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/t_static.log 2>&1; echo "static rc=$?"; tail -2 gpurun_out/t_static.log
timeout 120 python scripts/profile_layers.py bf16 2>&1 | grep -v "^\[" | awk 'NR==2||NR==6||NR==7||NR==11||NR==12||NR==16||NR==17||NR==21||NR==22'
timeout 300 python bench.py --steps 40 --warmup 6 --no-cpu-baseline > gpurun_out/bench_graph.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_graph.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step']); print(d['stages']['sparse_encoder'])" || tail -20 gpurun_out/bench_graph.log
