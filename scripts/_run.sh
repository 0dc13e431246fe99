timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_train_psnr.py -x -q -s -k "backward or psnr_matches" 2>&1 | grep -E "passed|failed|tc backward|PSNR on|Error|error|assert|\{" | tail -8
python bench.py --mode train --rays 16384 --steps 5 --no-cpu-baseline --no-parity-line > gpurun_out/bt.json 2> gpurun_out/bt.err
python -c "
import json
d=json.loads(open('gpurun_out/bt.json').read().strip().splitlines()[-1]); print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']))"
tail -2 gpurun_out/bt.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/launches_train3.csv python bench.py --mode train --steps 2 --warmup 3 --no-cpu-baseline --no-parity-line --rays 16384 > gpurun_out/ncu_bench.log 2>&1
