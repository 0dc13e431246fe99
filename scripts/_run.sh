timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline --mlp tc_bf16 --tables bf16 > gpurun_out/b1.json 2> gpurun_out/b1.err
python -c "
import json
d=json.loads(open('gpurun_out/b1.json').read().strip().splitlines()[-1]); print(round(d['value']), d['ms_per_step'], d['roofline']['stage_ms'], round(d['e2e']['value']))"
