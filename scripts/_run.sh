for n in 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_render_$n.json 2> gpurun_out/scale_$n.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 5 --warmup 3 --mode train --workload cfg3 > gpurun_out/scale_train_$n.json 2>> gpurun_out/scale_$n.err
done
for f in scale_render_4 scale_train_4 scale_render_8 scale_train_8; do python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', d.get('n_gpus'), round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['clocks'])"; done
grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/scale_8.err | tail -3
