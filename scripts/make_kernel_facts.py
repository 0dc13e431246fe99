"""profiles/kernel_facts.json from an `ncu --set full` capture of the dominant kernel: the numbers only a profiler sees
(dram bytes per launch, issue-slot / L1-data-pipe utilisation, warp instructions), tagged with the hash of the kernel sources
they were captured from -- bench.py reports them only when that hash equals the build it runs (never a stale constant).
Usage: python scripts/make_kernel_facts.py gpurun_out/fused_full.ncu-rep cfg2 tc_f16 "fused fine pass" 65536"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import source_sha          # noqa: E402
rep, workload, mlp, stage, rays = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
d = dict(zip(rows[0], rows[2]))
u = dict(zip(rows[0], rows[1]))
def val(k):
    v = float(d[k].replace(",", ""))
    unit = u[k].lower()
    return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3}.get(unit, 1.0)
samples = rays * 256
inst = val("smsp__inst_executed.sum")
rec = {"rays": rays, "kernel": d["Kernel Name"].split("(")[0],
       "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
       "issue": {"issue_slots_busy_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                 "l1_data_pipe_wavefronts_pct": val("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                 "tensor_pipe_pct": val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                 "l2_hit_pct": val("lts__t_sector_hit_rate.pct"), "l1_hit_pct": val("l1tex__t_sector_hit_rate.pct"),
                 "warp_instructions": inst, "warp_instructions_per_sample": inst / samples,
                 "duration_ms_under_ncu": val("gpu__time_duration.sum") / 1e6 if u["gpu__time_duration.sum"] in ("ns", "nsecond") else val("gpu__time_duration.sum") * {"us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}.get(u["gpu__time_duration.sum"], 1.0),
                 "source": os.path.basename(rep) + " (ncu --set full --clock-control none)"}}
path = os.path.join(ROOT, "profiles", "kernel_facts.json")
facts = json.load(open(path)) if os.path.isfile(path) else {}
if facts.get("source_sha") != source_sha():
    facts = {"source_sha": source_sha(), "kernels": {}}
facts["kernels"][f"{workload}:{mlp}:{stage}"] = rec
json.dump(facts, open(path, "w"), indent=1)
print(json.dumps(facts, indent=1))
