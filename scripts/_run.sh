for r in 4096 16384; do
python bench.py --mode train --rays $r --steps 30 --no-cpu-baseline --no-parity-line > gpurun_out/bt.json 2> gpurun_out/bt.err
python -c "
import json
d=json.loads(open('gpurun_out/bt.json').read().strip().splitlines()[-1]); print('train $r', round(d['value']), d['ms_per_step'], round(d['e2e']['value']))"
tail -1 gpurun_out/bt.err
done
