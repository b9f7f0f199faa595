timeout 500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for t in gradjw=0 gradjw=8; do
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --tune $t > gpurun_out/b_$t.json 2> gpurun_out/b.err; tail -c 600 gpurun_out/b.err
python -c "
import json; d=json.loads(open('gpurun_out/b_$t.json').read().strip().splitlines()[-1]); print('$t', d['value'], d['ms_per_step'], d['e2e']['value'], d['kernel_ms_per_step']['grad'], d['kernel_ms_per_step']['du_reduce'])"
done
