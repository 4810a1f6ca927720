"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown)."""
import csv, re, sys, collections
path, steps = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
rows = []
with open(path) as fh:
    lines = [l for l in fh if l.startswith('"')]
rd = csv.DictReader(lines)
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name).replace("nlps::<unnamed>::", "").replace("nlps::", "")
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
    tot[name][0] += 1; tot[name][1] += ms
total = sum(v[1] for v in tot.values())
print("| kernel | launches/step | ms/step | share |\n|---|---:|---:|---:|")
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %.1f | %.3f | %.1f%% |" % (k, n / steps, ms / steps, 100 * ms / total))
print("| **total** | %.1f | %.3f | |" % (sum(v[0] for v in tot.values()) / steps, total / steps))
