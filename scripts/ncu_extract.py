"""Condense an `ncu --page raw --csv` export to the columns the profiles/ summaries keep, and write the traffic JSON
bench.py reads (`roofline.traffic`).  usage: ncu_extract.py <raw.csv> <out.csv> [<traffic.json> "<kernel description>"]"""
import csv
import json
import sys

KEEP = ["ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "sm__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = [hdr.index(k) for k in KEEP if k in hdr]


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for d in data:
        w.writerow([d[i][:80] for i in idx])
if len(sys.argv) > 4:
    ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    # ncu picks a unit per VALUE in --page raw --csv only when units differ per column; the unit row applies to all rows
    rd = sum(to_bytes(d[ir], units[ir]) for d in data)
    wr = sum(to_bytes(d[iw], units[iw]) for d in data)
    tm = sum(float(d[it].replace(",", "")) for d in data)
    json.dump(dict(kernel=sys.argv[4], dram_read_bytes=rd, dram_write_bytes=wr, ncu_time_us=tm,
                   source="profiles/" + sys.argv[2].split("/")[-1], launches=len(data)), open(sys.argv[3], "w"), indent=1)
    print("launches", len(data), "dram read MB", rd / 1e6, "write MB", wr / 1e6, "time us", tm)
