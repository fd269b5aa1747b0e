"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST step
(a step starts at the marker kernel, default ccv_cdf).  python tools/agg_launches.py file.csv [marker] [--list pattern]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
marker = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "ccv_cdf"
pat = sys.argv[sys.argv.index("--list") + 1] if "--list" in sys.argv else None
with open(path) as f:
    rows = list(csv.DictReader(l for l in f if not l.startswith("==")))
idx = [i for i, r in enumerate(rows) if marker in r["Kernel Name"]]
# the loop prefetches the next batch's synthesis at the end of a step: a full step is the span between the last two markers
step = rows[idx[-2]:idx[-1]] if len(idx) >= 2 else (rows[idx[-1]:] if idx else rows)


def us(r):
    v = float(r["Metric Value"].replace(",", ""))
    return v / 1e3 if r["Metric Unit"] == "ns" else (v * 1e3 if r["Metric Unit"] == "ms" else v)


def short(n):
    return re.sub(r"\(.*", "", n).replace("void ", "").replace("at::native::", "")[:80]


if pat:
    for r in step:
        if re.search(pat, r["Kernel Name"]):
            print(f"{us(r):9.1f} us grid {r['Grid Size']:>18} block {r['Block Size']:>14} {short(r['Kernel Name'])}")
    sys.exit(0)
agg = collections.OrderedDict()
for r in step:
    a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
    a[0] += 1
    a[1] += us(r)
tot = sum(a[1] for a in agg.values())
print(f"total {tot:.1f} us over {len(step)} launches")
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| `{k}` | {a[0]} | {a[1]:.1f} | {a[1] / tot:.3f} |")
