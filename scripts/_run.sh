python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_render.json 2> gpurun_out/bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --mode train --rays 16384 > gpurun_out/bench_2gpu_train.json 2>> gpurun_out/bench_2gpu.err
python bench.py --steps 5 --warmup 3 --mode train --rays 16384 --no-cpu-baseline > gpurun_out/bench_1gpu_train.json 2>> gpurun_out/bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2>> gpurun_out/bench_2gpu.err
for f in bench_2gpu_render bench_2gpu_train bench_1gpu_train bench_2gpu_ref; do python -c "
import json,sys
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', d.get('n_gpus'), round(d['value']), d['ms_per_step'], d.get('e2e'))"; done
grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/bench_2gpu.err | tail -5
