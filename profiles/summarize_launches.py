"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`)
of tests/gpu_profile_step.py: per-kernel time shares of the LAST guided step and the DRAM traffic of the tcgen05 GEMM
launches (bench.py's roofline.traffic).  Usage: python profiles/summarize_launches.py launches.csv out.md out.json"""
import collections
import csv
import gzip
import json
import sys

src, out_md, out_json = sys.argv[1:4]
opener = gzip.open if src.endswith(".gz") else open
rows = list(csv.reader(l for l in opener(src, "rt") if l.startswith('"')))
h = rows[0]
ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
launch = collections.OrderedDict()
for r in rows[1:]:
    d = launch.setdefault(int(r[ii]), {"name": r[ki].split("(")[0].replace("void ", "").replace("mdc::", "")})
    d[r[mi]] = float(r[vi].replace(",", ""))
L = list(launch.values())
starts = [i for i, d in enumerate(L) if d["name"].endswith("begin_step_kernel")]
# the last COMPLETE step: between the last two markers (the final one may be cut off by a launch-count / time limit)
step = L[starts[-2]:starts[-1]] if len(starts) > 1 else L[starts[-1]:]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in step:
    a = agg[d["name"]]
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0) / 1e3          # ns -> us
    a[2] += d.get("dram__bytes_read.sum", 0.0)
    a[3] += d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
with open(out_md, "w") as f:
    f.write(f"# ncu launch list, one guided step ({len(step)} launches, {tot/1e3:.2f} ms serialised, cold caches: compare shares)\n\n")
    f.write("| kernel | launches | us | share | DRAM read MB | DRAM write MB |\n|---|---|---|---|---|---|\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| {k} | {a[0]} | {a[1]:.0f} | {100*a[1]/tot:.1f}% | {a[2]/1e6:.1f} | {a[3]/1e6:.1f} |\n")
g = [a for k, a in agg.items() if k.startswith("umma_gemm")]
n = sum(a[0] for a in g)
json.dump({"source": src, "kernels": [k for k in agg if k.startswith("umma_gemm")], "launches_per_step": n,
           "dram_bytes_per_step": sum(a[2] + a[3] for a in g), "dram_bytes_per_launch": sum(a[2] + a[3] for a in g) / max(n, 1),
           "us_per_step_under_ncu": sum(a[1] for a in g)}, open(out_json, "w"), indent=1)
print(open(out_md).read())
