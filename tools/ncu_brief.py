"""Brief per-launch table from an .ncu-rep (ncu -i ... --page raw --csv): time, DRAM bytes, occupancy, issue rate, top stalls.
    python tools/ncu_brief.py gpurun_out/x.ncu-rep [kernel-substring]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]


def g(r, k, d=float("nan")):
    try:
        return float(r[col[k]].replace(",", ""))
    except Exception:
        return d


for r in rows[2:]:
    name = r[col["Kernel Name"]]
    if flt and flt not in name:
        continue
    t = g(r, "gpu__time_duration.sum")
    unit = rows[1][col["gpu__time_duration.sum"]]
    rd, wr = g(r, "dram__bytes_read.sum"), g(r, "dram__bytes_write.sum")
    bu = rows[1][col["dram__bytes_read.sum"]]
    st = sorted(((g(r, s, 0.0), s.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for s in stalls), reverse=True)[:5]
    print("%s\n   grid %s regs %s  time %.3f %s  dram rd %.4f wr %.4f %s  dram_thr %.1f%%  warps_active %.1f%%  issue_active %.1f%%  inst %.3g  bank_conf %.3g / wavefronts %.3g\n   stalls: %s" % (
        name[:110], r[col["launch__grid_size"]], r[col["launch__registers_per_thread"]], t, unit, rd, wr, bu,
        g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), g(r, "smsp__inst_executed.sum"),
        g(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"), g(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        ", ".join("%s %.2f" % (n, v) for v, n in st)))
