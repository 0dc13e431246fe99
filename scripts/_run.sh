timeout 300 python -m pytest tests/test_gpu_tc.py -x -q -k "backward" 2>&1 | tail -2
python bench.py --mode train --rays 16384 --steps 20 --no-cpu-baseline --no-parity-line > gpurun_out/bt.json 2> gpurun_out/bt.err
python -c "
import json
d=json.loads(open('gpurun_out/bt.json').read().strip().splitlines()[-1]); print('train 16384', round(d['value']), d['ms_per_step'])"
tail -1 gpurun_out/bt.err
