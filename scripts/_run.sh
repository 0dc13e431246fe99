timeout 300 python -m pytest tests/test_gpu_tc.py -x -q -s 2>&1 | grep -E "passed|failed|tc_|Error|error" | tail -20
python bench.py --no-cpu-baseline > gpurun_out/bench_split.json 2> gpurun_out/bench_split.err
python bench.py --no-cpu-baseline --mlp tc_bf16 > gpurun_out/bench_bf16.json 2>> gpurun_out/bench_split.err
cat gpurun_out/bench_split.json gpurun_out/bench_bf16.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['mlp'], round(d['value']), d['ms_per_step'], d['roofline']['stage_ms'], d['e2e']['value'])
"
tail -3 gpurun_out/bench_split.err
