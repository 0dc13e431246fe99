#!/bin/bash
# One GPU-box session: parity tests, bench lines, ncu launch list.  Usage: gpu_round.sh [quick]
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest.log
tail -15 gpurun_out/pytest.log
if [ "$1" != "quick" ]; then
python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
python bench.py --mode train --rays 16384 --steps 5 --no-cpu-baseline > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_train.csv \
    python bench.py --mode train --steps 1 --warmup 3 --no-cpu-baseline --rays 8192 > gpurun_out/ncu_bench.log 2>&1
cat gpurun_out/bench_cfg2.json; tail -3 gpurun_out/bench_cfg2.err; cat gpurun_out/bench_train.json; tail -3 gpurun_out/bench_train.err
fi
