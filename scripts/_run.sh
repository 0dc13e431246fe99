timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
python bench.py --workload cfg3 --mode train --steps 5 --warmup 3 --no-cpu-baseline --no-parity-line > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
python bench.py --workload cfg1 --steps 50 --no-cpu-baseline > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err
for f in cfg5 cfg3 cfg1; do python -c "
import json
d=json.loads(open('gpurun_out/bench_$f.json').read().strip().splitlines()[-1]); print('$f', d['metric'], round(d['value']), d['ms_per_step'], d['roofline']['stage_ms'], round(d['e2e']['value']), d.get('parity_mode',{}).get('value'))"; tail -2 gpurun_out/bench_$f.err; done
