"""ncu launch list (gpu__time_duration per launch) -> per-kernel share table (markdown)."""
import collections, csv, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[vi]) / 1e3
tot = sum(a[1] for a in agg.values())
print(f"launches captured: {len(rows)}, total {tot:.1f} us (cold-cache, serialised under ncu: compare SHARES)\n")
print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {t:.1f} | {t / n:.2f} | {t / tot:.3f} |")
