for d in 0 4 6; do
timeout 120 python bench.py --steps 3 --warmup 2 --no-cpu --tune tcdbg=$d > gpurun_out/o_$d.log 2>gpurun_out/o_$d.err; tail -c 300 gpurun_out/o_$d.err; python -c "
import json,sys; d=json.loads(open('gpurun_out/o_$d.log').read().strip().splitlines()[-1]); print($d, {k:round(v['ms_per_step'],3) for k,v in d['kernel_ms_per_step'].items()})"
done
