#!/usr/bin/env python
"""Turn an ncu launch list (tools/gpu_launches.sh: gpu__time_duration + dram bytes per launch) of bench.py into
  profiles/<tag>_launches.md        per-kernel table of ONE job
  profiles/<tag>_traffic.json       DRAM bytes per job for every profiling scope of bench.py (what bench.py reports as roofline.traffic)
usage: summarize_launches.py gpurun_out/launches.csv <tag>
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    kn, mn, mv, idc = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    d = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        d.setdefault(r[idc], {"name": r[kn]})[r[mn]] = float(r[mv].replace(",", ""))
    return list(d.values())


def short(n):
    return n.split("(")[0].replace("void ", "")


SCOPE_OF = {  # kernels that are a profiling scope of their own
    "lq_pack_k": "pack", "lq_sketch_pk_k<5, 12, 8>": "sketch", "lq_sketch_pk_k<5, 15, 8>": "sketch", "lq_sketch_roll_k<5>": "sketch", "lq_sketch_roll_k<10>": "sketch",
    "lq_count_k": "idx_count", "lq_rs_hist_k": "radix_hist", "lq_rs_scatter_k": "radix_scatter", "lq_lookup_k": "seed_lookup",
    "lq_filter_count_k": "seed_filter", "lq_fill_k": "seed_fill", "lq_fill_masked_k": "seed_fill_filtered", "lq_gather_k": "seed_gather",
    "lq_runs_k": "runs", "lq_runs_emit_k": "runs", "lq_runs_q_k": "runs", "lq_chain_small_k": "chain_small", "lq_chain_k": "chain",
}


def main():
    path, tag = sys.argv[1], sys.argv[2]
    L = load(path)
    packs = [i for i, x in enumerate(L) if short(x["name"]) == "lq_pack_k"]
    job = L[packs[0]:packs[2]] if len(packs) > 2 else L          # a job packs twice (targets, queries)
    agg, scope = collections.OrderedDict(), collections.defaultdict(lambda: [0.0, 0.0, 0])
    level = -1
    for x in job:
        nm = short(x["name"])
        t = x.get("gpu__time_duration.sum", 0.0) / 1e6
        b = x.get("dram__bytes_read.sum", 0.0) + x.get("dram__bytes_write.sum", 0.0)
        a = agg.setdefault(nm, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += t; a[2] += x.get("dram__bytes_read.sum", 0.0); a[3] += x.get("dram__bytes_write.sum", 0.0)
        sc = SCOPE_OF.get(nm)
        if nm in ("lq_af_big_k<0>", "lq_af_big_k"):
            level += 1                                           # first launch of a sort level (shift 56, 48, ...)
        if nm in ("lq_af_big_k<0>", "lq_af_big_k", "lq_af_level_k"):
            sc = "seed_sort_s%d" % (56 - 8 * (level % 8))
        if nm in ("lq_af_walk_k", "lq_af_big_k<1>", "lq_af_walk3_k", "lq_af_walkf_k"):
            sc = "seed_walk_s%d" % (56 - 8 * (level % 8))
        if nm == "lq_af_place_k":
            sc = "seed_place_s%d" % (56 - 8 * (level % 8))
        if nm == "lq_af_walk_small_k":
            sc = "seed_walksmall_s%d" % (56 - 8 * (level % 8))
        if sc:
            s = scope[sc]; s[0] += t; s[1] += b; s[2] += 1
    tot = sum(a[1] for a in agg.values())
    out_md = os.path.join(ROOT, "profiles", tag + "_launches.md")
    with open(out_md, "w") as f:
        f.write("# %s -- ncu launch list of ONE job of `python bench.py --steps 1 --warmup 0 --no-cpu-baseline`\n\n" % tag)
        f.write("`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none` (tools/gpu_launches.sh);\n")
        f.write("per-launch times are cold-cache and serialised: the SHARES are what compares with bench.py's CUDA-event shares.\n\n")
        f.write("%d launches, %.1f ms of kernel time.\n\n| kernel | launches | ms | share | DRAM read GB | DRAM write GB |\n|---|---|---|---|---|---|\n" % (len(job), tot))
        for nm, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.3f | %.1f%% | %.2f | %.2f |\n" % (nm, a[0], a[1], 100 * a[1] / tot, a[2] / 1e9, a[3] / 1e9))
        f.write("\n## by profiling scope of bench.py\n\n| scope | launches | ms | DRAM bytes (read+write) GB |\n|---|---|---|---|\n")
        for sc, s in sorted(scope.items(), key=lambda kv: -kv[1][0]):
            f.write("| %s | %d | %.3f | %.2f |\n" % (sc, s[2], s[0], s[1] / 1e9))
    json.dump({sc: {"dram_bytes_per_job": s[1], "ms_under_ncu": s[0], "launches": s[2]} for sc, s in scope.items()},
              open(os.path.join(ROOT, "profiles", tag + "_traffic.json"), "w"), indent=1, sort_keys=True)
    print("wrote", out_md)


if __name__ == "__main__":
    main()
