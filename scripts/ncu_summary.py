"""Key metrics per kernel from an `ncu --page raw --csv` export.  usage: ncu_summary.py <raw.csv> [more metric substrings]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum"]
keys += [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h]
keys += [h for h in hdr for s in sys.argv[2:] if s in h]
ki = hdr.index("Kernel Name")
for d in data:
    print("=====", d[ki][:110])
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            v = d[i]
            try:
                if float(v.replace(",", "")) == 0 and "stalled" in k:
                    continue
            except ValueError:
                pass
            print("  %-95s %s %s" % (k.replace("smsp__average_warps_issue_stalled_", "stall:").replace("_per_issue_active.ratio", ""), v, units[i]))
