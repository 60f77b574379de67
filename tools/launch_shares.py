"""Shares of one steady-state block from an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`) of `bench.py`.
The list also holds the set-up (matrix upload fills, the 2s-step shift Lanczos, the first block) and the untimed checks
(orthogonality of the whole run: wide k_tsmm_tn launches); only the kernels of the timed blocks are counted here.
    python tools/launch_shares.py profiles/r2_launches_bench_steps3.csv"""
import collections
import csv
import re
import sys

BLOCK = [r"k_spmv_selr<1", r"k_spmv_selp<1", r"k_spmv_selld<1", r"k_tile<", r"k_tile_finalize<", r"k_chol_pan", r"k_halo_", r"k_allreduce_p2p"]
STEADY_END = r"k_tile<\d+, \d+, 3>"          # the fused update + solve pass closes a steady-state block


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").replace("calz::", "").replace("<unnamed>::", "")


def main(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            h, start = r, i
            break
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    seq = []
    for r in rows[start + 2:]:
        if len(r) <= vi:
            continue
        try:
            seq.append((short(r[ki]), float(r[vi].replace(",", ""))))
        except ValueError:
            continue
    # a steady-state block = a run of s Newton SpMV launches (with their halo kernels) followed by block kernels up to and including
    # the update + solve pass, with nothing else in between.  The set-up (shift Lanczos, first block) and the untimed checks use
    # the same tile kernels on other shapes since the panel path exists; they are not counted.
    agg = collections.OrderedDict()
    blocks = 0
    i = 0
    while i < len(seq):
        if not re.match(r"k_spmv_sel[prld]+<1", seq[i][0]) and not seq[i][0].startswith("k_halo_"):
            i += 1
            continue
        j = i
        while j < len(seq) and any(re.search(p, seq[j][0]) for p in BLOCK):
            if re.match(STEADY_END, seq[j][0]):
                break
            j += 1
        if j < len(seq) and re.match(STEADY_END, seq[j][0]) and sum(1 for k, _ in seq[i:j + 1] if k.startswith("k_spmv")) >= 1:
            blocks += 1
            for k, v in seq[i:j + 1]:
                a = agg.setdefault(k, [0, 0.0])
                a[0] += 1
                a[1] += v
            i = j + 1
        else:
            i = max(j, i + 1)
    blocks = max(1, blocks)
    tot = sum(t for c, t in agg.values())
    print("| kernel | launches / block | avg us | us / block | share |\n|---|---|---|---|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %.1f | %.1f | %.1f | %.1f %% |" % (k, c / blocks, t / c / 1e3, t / blocks / 1e3, 100 * t / tot))
    print("| total | | | %.1f | (%d steady-state blocks in the list; ncu serialises the launches and reads cold caches) |" % (tot / blocks / 1e3, blocks))


if __name__ == "__main__":
    main(sys.argv[1])
