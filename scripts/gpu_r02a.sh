#!/bin/bash
# r02 session A: parity tests, default bench line (+ reference arm), launch lists, ncu --set full of the fused fine pass
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
( time python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2> gpurun_out/bench_default.time
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py --mode train --rays 16384 --steps 10 --no-cpu-baseline --no-parity-line --no-extras > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 40 --csv --log-file gpurun_out/launches_render.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-parity-line --no-extras > gpurun_out/ncu_render.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/launches_train.csv \
    python bench.py --mode train --rays 16384 --steps 3 --warmup 3 --no-cpu-baseline --no-parity-line --no-extras > gpurun_out/ncu_train.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:egn_fused_fine -s 3 -c 1 -f -o gpurun_out/fused_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-line --no-extras > gpurun_out/ncu_fused.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:egn_coarse -s 3 -c 1 -f -o gpurun_out/coarse_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-line --no-extras > gpurun_out/ncu_coarse.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:egn_gather_bwd_tc -s 3 -c 1 -f -o gpurun_out/gbwd_full \
    python bench.py --mode train --rays 16384 --steps 2 --warmup 3 --no-cpu-baseline --no-parity-line --no-extras > gpurun_out/ncu_gbwd.log 2>&1
timeout 200 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_peer.py "tests/test_gpu_grad.py::test_gradients_with_reference_depths" -x -q > gpurun_out/san_memcheck_r02b.log 2>&1
tail -4 gpurun_out/san_memcheck_r02b.log
for f in default train reference; do python -c "
import json
d=json.loads(open('gpurun_out/bench_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d.get('roofline',{}).get('stage_ms'), (d.get('parity_mode') or {}).get('value'))"; tail -1 gpurun_out/bench_$f.err; done
cat gpurun_out/bench_default.time
