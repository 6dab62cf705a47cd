python tools/gpu_check.py 2>&1 | grep -E "FAIL|EXCEPTION|====|checks took" 
python bench.py --steps 20 --warmup 3 --no-cpu-baseline | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1300 --csv --log-file gpurun_out/launches_warm.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu_warm.log 2>&1
