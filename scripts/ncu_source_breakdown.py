"""Per-instruction-class breakdown of an .ncu-rep source page (SASS view): executed warp instructions, L1 tag requests, shared
wavefronts and stall samples, grouped by opcode and by code region.  Usage: ncu_source_breakdown.py x.ncu-rep [top]"""
import csv, io, subprocess, sys, collections
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {k: i for i, k in enumerate(hdr)}
def num(r, k):
    try: return float(r[ix[k]].replace(',', ''))
    except Exception: return 0.0
agg = collections.defaultdict(lambda: [0.0] * 6)
tot = [0.0] * 6
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    op = src.split()[0] if src else "?"
    if op.startswith("@"): op = src.split()[1]
    op = op.rstrip(";")
    key = op if op.startswith(("LDG", "STG", "LDS", "STS", "LDL", "STL", "RED", "ATOM", "LDTM", "STTM", "UTC", "SYNCS", "BAR", "SHFL", "MUFU")) else op.split(".")[0]
    vals = [num(r, "Instructions Executed"), num(r, "L1 Tag Requests Global"), num(r, "L1 Wavefronts Shared"), num(r, "# Samples"),
            num(r, "L2 Theoretical Sectors Global"), num(r, "stall_long_sb")]
    for i, v in enumerate(vals):
        agg[key][i] += v; tot[i] += v
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
print(f"{'opcode':28s} {'warp inst':>14s} {'L1 tag req':>14s} {'smem wavefr':>14s} {'samples':>10s} {'L2 sectors':>14s} {'long_sb':>9s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k:28s} {v[0]:14.0f} {v[1]:14.0f} {v[2]:14.0f} {v[3]:10.0f} {v[4]:14.0f} {v[5]:9.0f}")
print(f"{'TOTAL':28s} {tot[0]:14.0f} {tot[1]:14.0f} {tot[2]:14.0f} {tot[3]:10.0f} {tot[4]:14.0f} {tot[5]:9.0f}")
