timeout 300 python -m pytest tests/test_gpu_tc.py -x -q -k "fused" 2>&1 | tail -2
python bench.py --no-cpu-baseline --no-parity-line > gpurun_out/bd.json 2> gpurun_out/bd.err
python -c "
import json
d=json.loads(open('gpurun_out/bd.json').read().strip().splitlines()[-1]); print(round(d['value']), d['ms_per_step'], d['roofline']['stage_ms'], round(d['e2e']['value']))"
tail -1 gpurun_out/bd.err
