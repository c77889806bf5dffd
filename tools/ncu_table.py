"""Developer script: markdown table of the per-kernel metrics of an `ncu --page raw --csv` export (tools/ncu_full.sh)."""
import csv, sys, re
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
def g(r, name, scale=1.0, fmt="%.3g"):
    v = r[col[name]].replace(",", "")
    try: return fmt % (float(v) * scale)
    except ValueError: return v
def unit(name): return units[col[name]]
print("| kernel | time us | warp instr | issue active % | warps active % | threads/instr | L1 hit % | L2 hit % | DRAM read | DRAM write | regs | long-scoreboard stall/issue | grid | block | cluster |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in data:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("osep::", "").replace("void ", "")
    t = float(r[col["gpu__time_duration.sum"]].replace(",", "")); tu = unit("gpu__time_duration.sum")
    t = t / 1000 if tu == "ns" else (t * 1000 if tu == "ms" else t)
    print("| `%s` | %.1f | %s | %s | %s | %s | %s | %s | %s %s | %s %s | %s | %s | %s | %s | %s |" % (
        name, t, g(r, "smsp__inst_executed.sum"), g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        g(r, "smsp__thread_inst_executed_per_inst_executed.ratio"), g(r, "l1tex__t_sector_hit_rate.pct"), g(r, "lts__t_sector_hit_rate.pct"),
        g(r, "dram__bytes_read.sum"), unit("dram__bytes_read.sum"), g(r, "dram__bytes_write.sum"), unit("dram__bytes_write.sum"),
        g(r, "launch__registers_per_thread", fmt="%.0f"), g(r, "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        r[col["Grid Size"]], r[col["Block Size"]], g(r, "launch__cluster_size", fmt="%.0f")))
