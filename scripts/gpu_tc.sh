mkdir -p gpurun_out
set -x
timeout 300 python -m pytest tests/test_spconv_gpu.py -m gpu -q -x -k "bf16" -p no:cacheprovider > gpurun_out/t_sp_tc.log 2>&1; echo "sp tc rc=$?"
tail -25 gpurun_out/t_sp_tc.log
timeout 300 python -m pytest tests/test_spconv_gpu.py -m gpu -q -k "encoder or epilogue or rulebook" -p no:cacheprovider > gpurun_out/t_sp_enc.log 2>&1; echo "sp enc rc=$?"
tail -5 gpurun_out/t_sp_enc.log
for mt in 0 1 2; do BEVFRONT_TC_MT=$mt timeout 300 python scripts/profile_layers.py bf16 > gpurun_out/layers_bf16_mt$mt.log 2>&1; echo "layers mt=$mt rc=$?"; grep -v "^\[" gpurun_out/layers_bf16_mt$mt.log | tail -23; done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_bf16.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_bf16.log
