python tools/gpu_check.py 2>&1 | grep -E "FAIL|EXCEPTION|====|checks took" 
python tools/perf_layer.py --op hc_bwd --iters 20
python tools/perf_layer.py --op hc_bwd --L 180 --C 512 --iters 20
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:hc_post_bwd --csv python tools/perf_layer.py --op hc_bwd --iters 3 --warmup 1 2>&1 | grep hc_post | awk -F'","' '{print $5, $NF}' | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:hc_post_bwd --csv python tools/perf_layer.py --op hc_bwd --L 180 --C 512 --iters 3 --warmup 1 2>&1 | grep hc_post | awk -F'","' '{print $5, $NF}' | tail -3
python bench.py --steps 20 --warmup 3 --no-cpu-baseline
