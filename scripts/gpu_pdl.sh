mkdir -p gpurun_out
export BEVFRONT_TC_PDL=1
timeout 600 python -m pytest tests/test_spconv_gpu.py tests/test_static_gpu.py -m gpu -q -p no:cacheprovider --timeout 200 > gpurun_out/t_pdl.log 2>&1; rc=$?
echo "pdl tests rc=$rc"; tail -4 gpurun_out/t_pdl.log | cut -c1-300
if [ $rc -ne 0 ]; then exit 1; fi
for p in 1 0; do
BEVFRONT_TC_PDL=$p timeout 600 python bench.py --no-cpu-baseline 2>gpurun_out/bench_pdl$p.err | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('pdl', $p, d['value'], d['e2e']['value'], d['stages']['sparse_encoder']['gemm_ms'])"
done
