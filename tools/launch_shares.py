#!/usr/bin/env python
"""Summarises an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel of this repository and compares the
shares with the live CUDA-event stage times of a bench line.   tools/launch_shares.py launches.csv [bench.json]"""
import collections
import csv
import json
import re
import sys

STAGE = {  # kernel -> stage of scde_b200_stats
    "contract_i8_kernel": "contract", "contract_mma_kernel": "contract", "contract_generic_kernel": "contract",
    "softmax_i8_warp_kernel": "softmax", "softmax_i8_reduce_kernel": "softmax", "softmax_avg_kernel": "contract",
    "lp_rows_q_kernel": "lp_table", "lp_rows_fast_kernel": "lp_table", "lp_rows_kernel": "lp_table", "row_const_kernel": "lp_table",
    "cell_prep_kernel": "lp_table", "zero_rows_kernel": "lp_table", "based_flags_kernel": "lp_table", "row_cell_kernel": "lp_table",
    "quantize_rows_kernel": "lp_table",
    "dedup_bitmap_count_kernel": "dedup", "dedup_bitmap_emit_kernel": "dedup", "dedup_count_kernel": "dedup",
    "dedup_emit_kernel": "dedup", "exclusive_scan_kernel": "dedup",
    "build_w_kernel": "other", "build_lists_kernel": "other", "order_genes_kernel": "other", "iota_kernel": "other",
    "base_sum_partial_kernel": "other", "base_sum_reduce_kernel": "other", "w_to_i8_kernel": "other",
    "sentinel_range_kernel": "other",
    "ratio_summary_kernel": "ratio",
}


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
    hdr = next(r for r in rows if "Kernel Name" in r)
    agg = collections.OrderedDict()
    n_all, t_all = 0, 0.0
    for r in rows:
        if len(r) != len(hdr) or r == hdr:
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6}.get(d["Metric Unit"], 1e-6)
        n_all += 1
        t_all += v
        kn = d["Kernel Name"].replace("<unnamed>::", "").replace("void ", "")
        full = re.sub(r"\(.*", "", kn).split("::")[-1].strip()   # name with its template arguments
        name = re.sub(r"<.*", "", full)
        if name not in STAGE:
            continue
        a = agg.setdefault(full, [0, 0.0, STAGE[name]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    for k, (n, t, st) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%10.3f ms %5d launches %5.1f%%  %s" % (t, n, 100 * t / tot, k))
    print("%d launches in the csv, sum %.3f ms; kernels of this repository: %.3f ms" % (n_all, t_all, tot))
    by_stage = collections.defaultdict(float)
    for k, (n, t, st) in agg.items():
        by_stage[st] += t
    live = None
    if len(sys.argv) > 2:
        live = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])["roofline"]["stage_ms"]
    names = {"contract": "contraction", "lp_table": "lp table", "softmax": "soft-max", "other": "lists/W/base/ranges",
             "dedup": "dedup", "ratio": "ratio + summary"}
    print("# shares among these kernels%s" % (" vs live CUDA-event stage times of the plain run" if live else ""))
    lt = sum(live.values()) if live else 0
    for st in ("contract", "lp_table", "softmax", "other", "dedup", "ratio"):
        line = "#   %-22s %5.1f %%" % (names[st], 100 * by_stage[st] / tot)
        if live:
            line += "   live %5.1f %% (%.1f ms)" % (100 * live[st] / lt, live[st])
        print(line)


if __name__ == "__main__":
    main()
