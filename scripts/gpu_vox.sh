#!/bin/bash
# voxelizer: parity tests, in-graph timing (second vs first generation), ncu launch list of the kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_voxelize_gpu.py tests/test_reference_gpu.py -m gpu -q -x -p no:cacheprovider --timeout 300 2>&1 | tail -3
python scripts/vox_times.py; BEVFRONT_VOX_V1=1 python scripts/vox_times.py
ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum --clock-control none --csv --log-file gpurun_out/vox2_ncu.csv python scripts/profile_vox.py > gpurun_out/vox2_ncu.log 2>&1
python - <<'P'
import csv
rows=list(csv.reader(open('gpurun_out/vox2_ncu.csv')))
h=next(i for i,r in enumerate(rows) if "Kernel Name" in r)
hdr=rows[h]; kn=hdr.index("Kernel Name"); mn=hdr.index("Metric Name"); mv=hdr.index("Metric Value"); idc=hdr.index("ID")
d={}
for r in rows[h+1:]:
    if len(r)>mv: d.setdefault((int(r[idc]),r[kn][:44]),{})[r[mn]]=r[mv]
for k,v in sorted(d.items()):
    if k[0] in (6,7,8,15,16,17): print(k[0],k[1],' | '.join(f"{a.split('.')[0][-20:]}={b}" for a,b in v.items()))
P
