mkdir -p gpurun_out
for cfg in "1 220" "0 220" "1 128" "1 96"; do set -- $cfg; echo "=== L1=$1 SMEM_KB=$2"; BEVFRONT_TC_L1=$1 BEVFRONT_TC_SMEM_KB=$2 timeout 300 python scripts/profile_layers.py bf16 2>&1 | grep -v "^\[" | awk 'NR==2||NR==7||NR==8||NR==12||NR==13||NR==16||NR==17||NR==18||NR==22'; done
