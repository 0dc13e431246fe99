#!/bin/bash
# One GPU-box session: parity tests, bench lines, ncu launch lists.  Usage: gpu_round.sh [quick]
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
if [ "$1" != "quick" ]; then
( time python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2> gpurun_out/bench_default.time
python bench.py --mode train --rays 16384 --steps 10 --no-cpu-baseline --no-parity-line > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
python bench.py --mode train --steps 5 --no-cpu-baseline --no-parity-line > gpurun_out/bench_train64k.json 2> gpurun_out/bench_train64k.err
python bench.py --workload cfg1 --steps 50 --no-cpu-baseline > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 40 --csv --log-file gpurun_out/launches_render.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-parity-line > gpurun_out/ncu_render.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/launches_train.csv \
    python bench.py --mode train --rays 16384 --steps 3 --warmup 3 --no-cpu-baseline --no-parity-line > gpurun_out/ncu_train.log 2>&1
for f in default train train64k cfg1 reference; do python -c "
import json
d=json.loads(open('gpurun_out/bench_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d.get('roofline',{}).get('stage_ms'), (d.get('parity_mode') or {}).get('value'))"; tail -1 gpurun_out/bench_$f.err; done
cat gpurun_out/bench_default.time
fi
