"""Per-launch key metrics from an .ncu-rep: python tools/ncu_summary.py rep [metric-substring ...]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.avg",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"] + sys.argv[2:]
idx = [i for i, h in enumerate(hdr) if h in want]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    print("---", r[ki][:110])
    for i in idx:
        print(f"  {hdr[i]:70s} {r[i]:>16s} {rows[1][i]}")
