mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_spconv_gpu.py -m gpu -q -x -k "bf16 or encoder" -p no:cacheprovider > gpurun_out/t_sp_tc.log 2>&1; rc=$?; echo "tests rc=$rc"; tail -2 gpurun_out/t_sp_tc.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 120 python scripts/profile_layers.py bf16 2>&1 | grep -v "^\[" | awk 'NR==2||NR==6||NR==7||NR==11||NR==12||NR==16||NR==17||NR==21||NR==22'
