timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline --no-parity-line --mlp tc_split --tables f32 > gpurun_out/bp.json 2> gpurun_out/bp.err
python -c "
import json
d=json.loads(open('gpurun_out/bp.json').read().strip().splitlines()[-1]); print(round(d['value']), d['ms_per_step'], d['roofline']['stage_ms'])"
tail -1 gpurun_out/bp.err
