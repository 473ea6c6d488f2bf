"""ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv) -> one line per kernel launch with the numbers the roofline
discussion uses: duration, DRAM bytes and GB/s, L2 hit rate, SM / tensor-pipe / issue utilisation, registers, occupancy."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], newline="")))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__inst_executed.avg.per_cycle_elapsed", "launch__grid_size", "launch__block_size"]
units = rows[1]
print("%-64s %9s %10s %8s %6s %6s %6s %6s %6s %5s %5s %8s" % ("kernel (grid x block)", "us", "dram MB", "GB/s", "L2hit", "SM%", "tens%", "issue%", "warps%", "regs", "IPC", ""))
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    def g(name):
        i = col.get(name)
        if i is None or r[i] in ("", "n/a"):
            return None
        v = float(r[i].replace(",", ""))
        u = units[i]
        if name == "gpu__time_duration.sum":
            v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)         # -> us
        if name.startswith("dram__bytes"):
            v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        return v
    us = g("gpu__time_duration.sum")
    db = (g("dram__bytes_read.sum") or 0) + (g("dram__bytes_write.sum") or 0)
    tens = g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") or g("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active")
    issue = g("sm__issue_active.avg.pct_of_peak_sustained_active") or g("smsp__issue_active.avg.pct_of_peak_sustained_active")
    name = r[col["Kernel Name"]][:44]
    f = lambda v, fmt: (fmt % v) if v is not None else "-"
    print("%-64s %9s %10s %8s %6s %6s %6s %6s %6s %5s %5s" % (
        "%s (%sx%s)" % (name, f(g("launch__grid_size"), "%d"), f(g("launch__block_size"), "%d")), f(us, "%.1f"), f(db / 1e6, "%.2f"),
        f(db / us / 1e3 if us else None, "%.0f"), f(g("lts__t_sector_hit_rate.pct"), "%.0f"),
        f(g("sm__throughput.avg.pct_of_peak_sustained_elapsed"), "%.0f"), f(tens, "%.0f"), f(issue, "%.0f"),
        f(g("sm__warps_active.avg.pct_of_peak_sustained_active"), "%.0f"), f(g("launch__registers_per_thread"), "%d"),
        f(g("sm__inst_executed.avg.per_cycle_elapsed"), "%.2f")))
