"""summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share"""
import collections, csv, re, sys
path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))[skip:]
agg = collections.OrderedDict()
for r in rows:
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    v = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
    n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")[:64]
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"# {path}: {len(rows)} launches, {tot:.1f} ms total (ncu serialised, cold cache: compare SHARES)")
print("| kernel | launches | ms | share |\n|---|---:|---:|---:|")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:22]:
    print(f"| `{n}` | {c} | {t:.2f} | {100*t/tot:.1f}% |")
