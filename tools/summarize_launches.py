"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (markdown)."""
import collections, csv, re, sys
path, steps = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") == "gpu__time_duration.sum":
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("unnamed>::", "")
        agg[name][0] += 1
        agg[name][1] += float(row["Metric Value"].replace(",", ""))
tot = sum(v for _, v in agg.values())
print("| kernel | launches/step | ms/step | share |\n|---|---:|---:|---:|")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %.1f | %.3f | %.1f%% |" % (k, n / steps, v / 1e6 / steps, 100 * v / tot))
print("| **total** | %.1f | %.3f | 100%% |" % (sum(n for n, _ in agg.values()) / steps, tot / 1e6 / steps))
