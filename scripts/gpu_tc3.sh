for st in 2 3 4 6; do echo "=== STAGES=$st"; BEVFRONT_TC_STAGES=$st timeout 300 python scripts/profile_layers.py bf16 2>&1 | grep -v "^\[" | awk 'NR==2||NR==7||NR==12||NR==17'; done
echo "=== NO_HALO"; BEVFRONT_TC_NO_HALO=1 timeout 300 python scripts/profile_layers.py bf16 2>&1 | grep -v "^\[" | awk 'NR==2||NR==7||NR==12||NR==17'
