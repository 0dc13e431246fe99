"""Warp-group split of a fused-kernel .ncu-rep: stall samples and executed instructions of the gather code (before the first
UTCHMMA of the SASS) and of the MLP code (after), plus the mbarrier wait sites with their samples.
Usage: python scripts/ncu_groups.py x.ncu-rep [tiles]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
tiles = float(sys.argv[2]) if len(sys.argv) > 2 else 131072.0
rows = list(csv.reader(io.StringIO(out)))
for i, r in enumerate(rows):
    if '# Samples' in r:
        hdr, start = r, i + 1
        break
ix = {k: i for i, k in enumerate(hdr)}
ins = [r for r in rows[start:] if len(r) >= len(hdr)]
def num(r, k):
    try: return float(r[ix[k]].replace(',', ''))
    except Exception: return 0.0
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
first = min(j for j, r in enumerate(ins) if 'UTCHMMA' in r[ix['Source']]) - 60
tot = sum(num(r, '# Samples') for r in ins)
def summarize(name, a, b):
    s = sum(num(r, '# Samples') for r in ins[a:b]); n = sum(num(r, 'Instructions Executed') for r in ins[a:b])
    st = {k: sum(num(r, k) for r in ins[a:b]) for k in stalls}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:7]
    print(f"{name:10s} samples {int(s):7d} ({s / tot * 100:4.1f}%)  inst/tile {n / tiles:8.1f}  " + ' '.join(f"{k[6:]} {v / max(s, 1) * 100:.0f}%" for k, v in top))
summarize('gather', 0, first)
summarize('mlp', first, len(ins))
print("wait sites (NANOSLEEP.SYNCS):")
for j, r in enumerate(ins):
    if 'NANOSLEEP' in r[ix['Source']] and num(r, '# Samples') > 0:
        # samples of the whole poll loop: +-3 instructions around
        s = sum(num(x, '# Samples') for x in ins[max(0, j - 3):j + 4])
        print(f"  sass #{j} ({'gather' if j < first else 'mlp'}): {int(s)} samples ({s / tot * 100:.1f}%), {num(r, 'Instructions Executed') / tiles:.0f} polls/tile")
for j, r in enumerate(ins):
    if 'BAR.SYNC' in r[ix['Source']] and num(r, '# Samples') > 200:
        print(f"  sass #{j} BAR.SYNC: {int(num(r, '# Samples'))} samples ({num(r, '# Samples') / tot * 100:.1f}%)")
