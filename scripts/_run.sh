timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline --no-parity-line > gpurun_out/bd.json 2> gpurun_out/bd.err
python -c "
import json
d=json.loads(open('gpurun_out/bd.json').read().strip().splitlines()[-1]); print(round(d['value']), d['ms_per_step'], d['roofline']['stage_ms'])"
python bench.py --no-cpu-baseline --no-parity-line --mode train --rays 16384 --steps 5 > gpurun_out/bt.json 2> gpurun_out/bt.err
python -c "
import json
d=json.loads(open('gpurun_out/bt.json').read().strip().splitlines()[-1]); print('train', round(d['value']), d['ms_per_step'])"
