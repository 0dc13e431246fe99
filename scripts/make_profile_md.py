"""profiles/<tag>_end.md from one GPU session's gpurun_out/ (scripts/gpu_r02a.sh): bench lines, launch lists, ncu summaries.
Usage: python scripts/make_profile_md.py r02"""
import json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
out = []
def line(name):
    return json.loads(open(os.path.join(G, name)).read().strip().splitlines()[-1])
def run(*cmd):
    return subprocess.run([sys.executable, *cmd], capture_output=True, text=True, cwd=ROOT).stdout
d = line("bench_default.json")
tr = line("bench_train.json")
ref = line("bench_reference.json")
out.append(f"# {tag} — end-of-round measurements (B200, one GPU; produced by scripts/gpu_r02a.sh + scripts/make_profile_md.py)\n")
out.append(f"`pytest tests -m gpu`: {open(os.path.join(G, 'pytest.log')).read().strip().splitlines()[-1]}\n")
out.append("## bench.py lines (device-timed CUDA events, clocks sampled during the timed region)\n")
out.append("| run | value | ms / step | e2e | notes |\n|---|---|---|---|---|")
st = d["roofline"]["stage_ms"]
out.append(f"| `python bench.py` (cfg2 render, 65 536 rays, default mode tc_f16) | {d['value']/1e6:.2f} M rays/s | {d['ms_per_step']:.3f} | {d['e2e']['value']/1e6:.2f} M rays/s (with alpha to the host: {d['e2e']['with_alpha']['value']/1e6:.2f} M) | stages: sampler {list(st.values())[0]} ms, fused fine pass {list(st.values())[1]} ms; clocks {d['clocks']['sm_mhz']}/{d['clocks']['sm_max_mhz']} MHz {d['clocks']['reasons']} |")
pm = d.get("parity_mode")
if pm:
    out.append(f"| … `parity_mode` (fp32-equivalent, tc_split) | {pm['value']/1e6:.2f} M rays/s | | | stages {pm['stage_ms']} |")
for k, t in (d.get("train") or {}).items():
    out.append(f"| … `train.{k}` (16 384 rays, fwd + bwd + exchange + TableAdam) | {t['value']/1e6:.3f} M rays/s | {t['ms_per_step']:.3f} | | exchange {t['allreduce_ms']:.3f} ms ({t['allreduce']['what'][:60]}) |")
e = d.get("erp_frame")
if e:
    out.append(f"| … `erp_frame` (cfg5: 256-row tile, 256 + 512 samples) | {e['value']/1e6:.2f} M rays/s | {e['ms_per_step']:.1f} | | full 4096 x 2048 frame at this rate: {e['full_frame_ms_at_this_rate']:.0f} ms on one GPU |")
out.append(f"| `python bench.py --mode train --rays 16384` | {tr['value']/1e6:.3f} M rays/s | {tr['ms_per_step']:.3f} | {tr['e2e']['value']/1e6:.3f} M rays/s | |")
out.append(f"| `python bench.py --impl reference` (CPU port, {ref['cpu_baseline']['cores']} host threads, 4 096 rays / step) | {ref['value']:.0f} rays/s | {ref['ms_per_step']:.0f} | | in-line leg of the default line: {d['cpu_baseline']['value']:.0f} rays/s |")
rf = d["roofline"]
out.append(f"\nroofline of the default line: kernel `{rf['kernel']}`, achieved {rf['achieved']:.0f} GB/s of tap-model bytes vs peak {rf['peak']:.0f} "
           f"({rf['peak_source']}) = {rf['frac']:.2f}; traffic (ncu dram bytes per launch) {rf['traffic']}; issue: {json.dumps(rf['issue'])}\n")
out.append("Full default line:\n\n```json\n" + json.dumps(d) + "\n```\n")
for kind in ("render", "train"):
    src = os.path.join(G, f"launches_{kind}.csv")
    dst = os.path.join(ROOT, "profiles", f"{tag}_end_launches_{kind}.csv")
    shutil.copy(src, dst)
    out.append(f"## ncu launch list, {kind} (`ncu --metrics gpu__time_duration.sum --clock-control none`, `profiles/{tag}_end_launches_{kind}.csv`)\n")
    out.append(run("scripts/summarize_launches.py", dst))
for rep, title in (("fused_full.ncu-rep", "egn_fused_fine_kernel<COMP>"), ("coarse_full.ncu-rep", "egn_coarse_kernel")):
    if os.path.isfile(os.path.join(G, rep)):
        out.append(f"## ncu --set full: {title} (cfg2, 65 536 rays)\n\n```\n" + run("scripts/ncu_summary.py", os.path.join(G, rep)) + "```\n")
sass = subprocess.run("cuobjdump -sass egonerf_b200/libegn_b200.so | grep -oE '\\b(UTCHMMA|LDTM|STTM|UTCBAR|UBLKCP|UTMALDG|HMMA|REDG|SYNCS|HFMA2|UTCATOMSWS)[A-Z0-9_.]*' | sed 's/\\..*//' | sort | uniq -c",
                      shell=True, capture_output=True, text=True, cwd=ROOT).stdout
out.append("## SASS mnemonics of the shipped libegn_b200.so (`cuobjdump -sass`)\n\n```\n" + sass + "```\n")
open(os.path.join(ROOT, "profiles", f"{tag}_end.md"), "w").write("\n".join(out))
print("wrote", f"profiles/{tag}_end.md")
