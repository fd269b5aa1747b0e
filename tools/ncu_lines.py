"""Per-source-line totals of an ncu report: `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > f.csv;
python tools/ncu_lines.py f.csv [top]` -> instructions executed, share, stall samples and lane occupancy per CUDA line."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
        ci, ti, si = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    elif hdr and r[0] not in ("", "Function Name") and len(r) > ci:
        try:
            out.append((int(r[ci]), int(r[si]), int(r[ti]), cur_file, r[0], r[1][:100]))
        except ValueError:
            pass
tot, tots = sum(o[0] for o in out), sum(o[1] for o in out)
print(f"total warp instructions {tot / 1e6:.2f} M, samples {tots}")
for n, s, t, f, l, src in sorted(out, reverse=True)[:top]:
    print(f"{n / 1e6:8.2f}M {100 * n / tot:5.1f}%  smp {100 * s / max(tots, 1):5.1f}%  lanes {t / max(n, 1):5.1f} | {f}:{l} {src}")
if len(sys.argv) > 3:   # line-range totals: a,b,c,... boundaries within the first file
    bounds = [int(x) for x in sys.argv[3].split(",")]
    f0 = out[0][3] if out else None
    acc = {}
    for n, s, t, f, l, src in out:
        if f != "raster.cu":
            key = f
        else:
            li = int(l)
            key = "raster.cu:" + str(max([b for b in bounds if b <= li], default=0))
        a = acc.setdefault(key, [0, 0, 0])
        a[0] += n; a[1] += s; a[2] += t
    for k, a in sorted(acc.items(), key=lambda kv: -kv[1][0]):
        print(f"{k:40s} {a[0] / 1e6:8.2f}M {100 * a[0] / tot:5.1f}%  smp {100 * a[1] / max(tots, 1):5.1f}% lanes {a[2] / max(a[0], 1):5.1f}")
