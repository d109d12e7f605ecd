#!/bin/bash
# round 2, call A: full GPU test suite (incl. the reference-CUDA differential tests) + config-A parity measurement
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
python scripts/parity_config_a.py 0 > gpurun_out/r2a_parity.log 2>&1; echo "parity rc=$?"
tail -5 gpurun_out/r2a_parity.log
