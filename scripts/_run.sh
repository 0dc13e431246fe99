timeout 300 python -m pytest tests/test_gpu_tc.py -x -q -s 2>&1 | grep -E "passed|failed|fused|Error|error|assert" | tail -12
for t in f32 bf16; do
python bench.py --no-cpu-baseline --mlp tc_bf16 --tables $t > gpurun_out/bench_fused_$t.json 2> gpurun_out/bench_fused_$t.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_fused_$t.json').read().strip().splitlines()[-1]); print('$t', round(d['value']), d['ms_per_step'], d['roofline']['stage_ms'], round(d['e2e']['value']))"
tail -2 gpurun_out/bench_fused_$t.err
done
