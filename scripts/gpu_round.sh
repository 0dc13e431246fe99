#!/bin/bash
# One GPU-box session: parity tests, bench lines, ncu launch list (+ optional full capture of the top kernel).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest.log
python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
python bench.py --workload cfg1 --steps 50 --no-cpu-baseline > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --rays 16384 > gpurun_out/ncu_bench.log 2>&1
tail -5 gpurun_out/pytest.log; cat gpurun_out/bench_cfg2.json; tail -3 gpurun_out/bench_cfg2.err
