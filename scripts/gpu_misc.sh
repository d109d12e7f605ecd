mkdir -p gpurun_out
for inf in 3 2; do
timeout 300 python bench.py --steps 60 --warmup 6 --no-cpu-baseline --inflight $inf > gpurun_out/bench_inf$inf.log 2>&1; echo "bench inflight=$inf rc=$?"; tail -3 gpurun_out/bench_inf$inf.log | grep -v "^{" ; tail -1 gpurun_out/bench_inf$inf.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])" || tail -20 gpurun_out/bench_inf$inf.log
done
