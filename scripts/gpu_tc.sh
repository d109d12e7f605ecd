mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_spconv_gpu.py -m gpu -q -x -k "bf16 or encoder or epilogue" -p no:cacheprovider > gpurun_out/t_sp_tc.log 2>&1; echo "sp tc rc=$?"
tail -3 gpurun_out/t_sp_tc.log
for cfg in "0 0" "1 0" "2 0"; do set -- $cfg; echo "=== MT=$1 NB=$2"; BEVFRONT_TC_MT=$1 BEVFRONT_TC_NB=$2 timeout 300 python scripts/profile_layers.py bf16 2>&1 | grep -v "^\[" | awk 'NR==2||NR==6||NR==7||NR==11||NR==12||NR==16||NR==17||NR==21||NR==22'; done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_bf16.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_bf16.log | cut -c1-300
