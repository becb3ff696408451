"""Print the key metrics of every kernel instance in an .ncu-rep (raw page)."""
import csv, subprocess, sys
WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__inst_executed_pipe_tensor", "sm__pipe_tensor_cycles_active",
        "sm__pipe_tensor_op", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warp_issue_stalled"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:80], "grid", r[hdr.index("Grid Size")] if "Grid Size" in hdr else "")
    for h, u, v in zip(hdr, units, r):
        if any(h.startswith(w) for w in WANT) and "peak_sustained" not in h.replace("pct_of_peak_sustained", "") :
            if h.endswith(".per_second") or ".pct_of_peak_sustained_elapsed" in h and not h.startswith(("sm__throughput", "gpu__dram", "lts__throughput", "l1tex__throughput", "sm__pipe_tensor", "sm__inst_executed_pipe_tensor")):
                continue
            print(f"   {h:90s} {v} {u}")
