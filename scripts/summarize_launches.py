"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share."""
import csv, sys, collections, re
path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"): v *= 1e3
        if r["Metric Unit"] in ("ms", "msecond"): v *= 1e6
        rows.append((name, v, r["Grid Size"], r["Block Size"]))
agg = collections.OrderedDict()
for n, v, g, b in rows:
    a = agg.setdefault(n, [0, 0.0, g, b]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
own = sum(a[1] for n, a in agg.items() if n.startswith("egn_"))
print(f"| kernel | launches | grid | block | total ms | mean us | share of all | share of egn_* |\n|---|---|---|---|---|---|---|---|")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{n[:70]}` | {a[0]} | {a[2]} | {a[3]} | {a[1]/1e6:.3f} | {a[1]/a[0]/1e3:.1f} | {100*a[1]/tot:.1f}% | {(100*a[1]/own if n.startswith('egn_') else 0):.1f}% |")
