set -x
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo bench=$?; tail -c 300 gpurun_out/final_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo ref=$?
timeout 120 ./tools/probe_overlap > gpurun_out/probe_overlap.txt 2>&1; echo probe=$?
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1; echo plain=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_l.log 2>&1; echo ncul=$?
