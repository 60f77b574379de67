"""Turn the ncu artefacts of one bench command into the markdown summary kept under profiles/.

    python tools/profile_summary.py LAUNCHES.csv [FULL.ncu-rep] > profiles/<name>.md

LAUNCHES.csv: ncu --metrics gpu__time_duration.sum --clock-control none --csv ...   (launch list)
FULL.ncu-rep: ncu --set full --clock-control none --import-source on ...            (read with ncu -i --page raw --csv)
"""
import csv
import re
import subprocess
import sys


def short(name):
    name = re.sub(r"\(.*", "", name)
    for junk in ("void ", "calz::", "(anonymous namespace)::", "<unnamed>::"):
        name = name.replace(junk, "")
    return name


def launch_table(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    seq = []
    for r in rows[1:]:
        try:
            seq.append((short(r[ki]), float(r[vi].replace(",", "")) / 1000.0))
        except ValueError:
            pass
    spmv = [i for i, (k, _) in enumerate(seq) if "spmv" in k]
    blk = seq[spmv[-8]:]                      # one steady-state block: from its first MPK step on
    agg, order = {}, []
    for k, t in blk:
        if k not in agg:
            agg[k] = [0, 0.0]
            order.append(k)
        agg[k][0] += 1
        agg[k][1] += t
    tot = sum(t for _, t in blk)
    out = ["| kernel | launches/block | us/block | share |", "|---|---:|---:|---:|"]
    for k in order:
        out.append("| %s | %d | %.1f | %.1f%% |" % (k, agg[k][0], agg[k][1], 100 * agg[k][1] / tot))
    out.append("| total | %d | %.1f | 100%% |" % (len(blk), tot))
    mpk = sum(v[1] for k, v in agg.items() if "spmv" in k)
    out.append("")
    out.append("MPK share under ncu: %.1f%%." % (100 * mpk / tot))
    return "\n".join(out)


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def full_tables(rep, skip_repeats=True):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units = r[0], r[1]
    out, seen = [], set()
    for row in r[2:]:
        name = short(row[hdr.index("Kernel Name")])
        if skip_repeats and name in seen:
            continue
        seen.add(name)
        out += ["### " + name, "| metric | value | unit |", "|---|---:|---|"]
        for w in WANT:
            if w in hdr:
                out.append("| %s | %s | %s |" % (w, row[hdr.index(w)], units[hdr.index(w)]))
        out.append("")
    return "\n".join(out)


if __name__ == "__main__":
    print(launch_table(sys.argv[1]))
    if len(sys.argv) > 2:
        print()
        print(full_tables(sys.argv[2]))
