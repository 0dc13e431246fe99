"""Key metrics of an .ncu-rep (read offline): python scripts/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv, subprocess, sys, io
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_op_global_red.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2:]
for v in vals:
    d = dict(zip(hdr, v)); u = dict(zip(hdr, units))
    print("kernel:", d.get("Kernel Name", "")[:60], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
    for k in hdr:
        if k in KEYS or (len(sys.argv) > 2 and sys.argv[2] in k):
            print(f"  {k:75s} {d[k]:>18s} {u[k]}")
    stalls = sorted(((float(d[k].replace(',', '')), k) for k in hdr if k.startswith("smsp__average_warp") and "issue_stalled" in k and k.endswith("_per_warp_active.pct") is False and d[k] not in ("", "n/a")), reverse=True)[:8]
    for s, k in stalls:
        print(f"  stall {k.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_','')[:60]:62s} {s:10.3f}")
