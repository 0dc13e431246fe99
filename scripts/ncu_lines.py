"""Per-source-line aggregation of an .ncu-rep (needs -lineinfo + --import-source on): executed warp instructions and stall
samples per CUDA line, grouped per file, top lines first.  Usage: ncu_lines.py x.ncu-rep [top]"""
import csv, io, subprocess, sys, collections
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, ix = "?", None, None
lines = []          # (file, line, src, samples, inst, stall dict)
def num(r, k):
    try: return float(r[ix[k]].replace(',', ''))
    except Exception: return 0.0
for r in rows:
    if len(r) >= 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]; continue
    if "# Samples" in r:
        hdr = r; ix = {}
        for i, k in enumerate(hdr):
            ix.setdefault(k, i)
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    st = {k[6:]: num(r, k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
    lines.append((fname, int(r[0]), r[1].strip(), num(r, "# Samples"), num(r, "Instructions Executed"), st))
tot_s = sum(l[3] for l in lines); tot_i = sum(l[4] for l in lines)
print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.0f}")
agg = collections.defaultdict(float)
for l in lines:
    for k, v in l[5].items(): agg[k] += v
print("stall reasons (all lines):", ", ".join(f"{k} {v / max(tot_s,1) * 100:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
print(f"{'file:line':28s} {'samples%':>8s} {'inst%':>7s} {'top stall':>22s}  source")
for l in sorted(lines, key=lambda l: -l[3])[:top]:
    ts = max(l[5].items(), key=lambda kv: kv[1]) if l[5] else ("-", 0)
    print(f"{l[0] + ':' + str(l[1]):28s} {l[3] / tot_s * 100:8.2f} {l[4] / max(tot_i,1) * 100:7.2f} {ts[0] + ' ' + format(ts[1] / max(l[3],1) * 100, '.0f') + '%':>22s}  {l[2][:90]}")
