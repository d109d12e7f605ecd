mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
nproc
timeout 600 python -m pytest tests -m gpu -q -x --deselect tests/test_spconv_gpu.py -p no:cacheprovider > gpurun_out/t_core.log 2>&1; echo "core rc=$?"
timeout 300 python -m pytest tests/test_spconv_gpu.py -m gpu -q -k "rulebook or index_flags or fp32 or training" -p no:cacheprovider > gpurun_out/t_sp_f32.log 2>&1; echo "sp f32 rc=$?"
timeout 300 python -m pytest tests/test_spconv_gpu.py -m gpu -q -k "bf16" -p no:cacheprovider > gpurun_out/t_sp_bf16.log 2>&1; echo "sp bf16 rc=$?"
timeout 300 python -m pytest tests/test_spconv_gpu.py -m gpu -q -k "epilogue or encoder" -p no:cacheprovider > gpurun_out/t_sp_enc.log 2>&1; echo "sp enc rc=$?"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --precision fp32 > gpurun_out/bench_fp32.log 2>&1; echo "bench fp32 rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16.log 2>&1; echo "bench bf16 rc=$?"
tail -3 gpurun_out/t_*.log gpurun_out/smoke.log gpurun_out/bench_*.log
