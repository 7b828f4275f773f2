"""Prints the handful of ncu metrics the DESIGN/profiles notes quote, from a .ncu-rep (via ncu --page raw --csv)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__inst_executed.avg.per_cycle_active"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "")[:70])
    for k in hdr:
        if k in want or ("issue_stalled" in k and "per_issue_active" in k):
            i = hdr.index(k)
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                v = r[i]
            if "issue_stalled" in k and isinstance(v, float) and v < 0.3:
                continue
            print("  %-85s %-12s %s" % (k, units[i], v))
