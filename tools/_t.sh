timeout 500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for r in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/b_$r.json 2> gpurun_out/b.err; tail -c 300 gpurun_out/b.err
python -c "
import json; d=json.loads(open('gpurun_out/b_$r.json').read().strip().splitlines()[-1]); print($r, d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']); print({k:round(v['ms_per_step'],3) for k,v in d['kernel_ms_per_step'].items()})"
done
