for cfg in "4" "0"; do echo "=== MT=$cfg"; BEVFRONT_TC_MT=$cfg timeout 120 python scripts/profile_layers.py bf16 2>&1 | grep -v "^\[" | awk 'NR==2||NR==6||NR==7||NR==11||NR==22'; done
