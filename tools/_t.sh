timeout 500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for t in sbstaged=1; do
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --tune $t > gpurun_out/b_$t.json 2> gpurun_out/b.err; tail -c 600 gpurun_out/b.err
python -c "
import json; d=json.loads(open('gpurun_out/b_$t.json').read().strip().splitlines()[-1]); print('$t', d['value'], d['ms_per_step'], d['e2e']['value']); print({k:round(v['ms_per_step'],3) for k,v in d['kernel_ms_per_step'].items()})"
done
