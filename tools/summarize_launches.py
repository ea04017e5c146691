"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total time and share per kernel."""
import collections, csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt, mx = collections.defaultdict(float), collections.Counter(), collections.defaultdict(float)
for r in rows[1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    if r[ui] == "us": v *= 1e3
    if r[ui] == "ms": v *= 1e6
    name = r[ki].split("(")[0].replace("void ", "")
    tot[name] += v; cnt[name] += 1; mx[name] = max(mx[name], v)
s = sum(tot.values())
print(f"{'kernel':28s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'max ms':>9s}")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"{k:28s} {cnt[k]:8d} {v/1e6:10.3f} {v/s*100:6.1f}% {mx[k]/1e6:9.3f}")
print(f"{'total':28s} {sum(cnt.values()):8d} {s/1e6:10.3f}")
