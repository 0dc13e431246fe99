timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -k "backward" 2>&1 | tail -3
EGN_TC_BACKWARD=1 python bench.py --mode train --rays 16384 --steps 10 --no-cpu-baseline --no-parity-line --mlp tc_split --tables f32 > gpurun_out/bt.json 2> gpurun_out/bt.err
python -c "
import json
d=json.loads(open('gpurun_out/bt.json').read().strip().splitlines()[-1]); print('train parity-fwd + tc-bwd 16384', round(d['value']), d['ms_per_step'])"
tail -1 gpurun_out/bt.err
