"""Shares of one steady-state block from an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`) of `bench.py`.
The list also holds the set-up (matrix upload fills, the 2s-step shift Lanczos, the first block) and the untimed checks
(orthogonality of the whole run: wide k_tsmm_tn launches); only the kernels of the timed blocks are counted here.
    python tools/launch_shares.py profiles/r2_launches_bench_steps3.csv"""
import collections
import csv
import re
import sys

BLOCK = [r"k_spmv_selr<1", r"k_spmv_selp<1", r"k_spmv_selld<1", r"k_tile<", r"k_tile_finalize<", r"k_chol_pan", r"k_halo_", r"k_allreduce_p2p"]


def main(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            h, start = r, i
            break
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[start + 2:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki]
        if not any(re.search(p, name) for p in BLOCK):
            continue
        key = re.sub(r"\(.*", "", name).replace("void ", "").replace("calz::", "").replace("<unnamed>::", "")
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
    # launches per block: SpMV count / s identifies the number of blocks in the list
    nspmv = sum(c for k, (c, t) in agg.items() if k.startswith("k_spmv"))
    blocks = max(1, nspmv // 8)
    tot = sum(t for c, t in agg.values())
    print("| kernel | launches / block | avg us | us / block | share |\n|---|---|---|---|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %.1f | %.1f | %.1f | %.1f %% |" % (k, c / blocks, t / c / 1e3, t / blocks / 1e3, 100 * t / tot))
    print("| total | | | %.1f | (%d blocks in the list; ncu serialises the launches and reads cold caches) |" % (tot / blocks / 1e3, blocks))


if __name__ == "__main__":
    main(sys.argv[1])
