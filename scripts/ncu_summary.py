#!/usr/bin/env python
"""Print the metrics that matter from an .ncu-rep (raw page): duration, tensor pipe, DRAM/L2 traffic, occupancy, stalls."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform.sum",
        "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.ratio", "smsp__sass_average_data_bytes_per_sector_mem_global_op_st.ratio",
        "l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio",
        "l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_st.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum"]
extra = [h for h in hdr if "pcsamp" in h or "issue_stalled" in h and "pct" in h]
for r in data:
    print("== ", r[hdr.index("Kernel Name")][:60], " id", r[0])
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print("  %-85s %s %s" % (k, r[i], units[i]))
    st = [(h, r[hdr.index(h)]) for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    st = sorted(((float(v), h) for h, v in st if v not in ("", "n/a")), reverse=True)[:8]
    for v, h in st:
        print("  stall %-60s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
