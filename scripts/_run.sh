timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_train_psnr.py tests/test_gpu_parity.py -x -q -s -k "backward or psnr_matches or checkpoint" 2>&1 | grep -E "passed|failed|tc backward|PSNR on|Error|error|assert" | tail -8
python bench.py --mode train --rays 16384 --steps 10 --no-cpu-baseline --no-parity-line > gpurun_out/bt.json 2> gpurun_out/bt.err
python -c "
import json
d=json.loads(open('gpurun_out/bt.json').read().strip().splitlines()[-1]); print('train', round(d['value']), d['ms_per_step'])"
tail -1 gpurun_out/bt.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 75 --csv --log-file gpurun_out/launches_train.csv python bench.py --mode train --rays 16384 --steps 3 --warmup 3 --no-cpu-baseline --no-parity-line > gpurun_out/ncu_train.log 2>&1
