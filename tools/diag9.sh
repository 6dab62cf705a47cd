python tools/gpu_check.py 2>&1 | grep -E "FAIL|EXCEPTION|====|checks took" 
for f in 1024 1280 1536 2560; do
echo "flags $f"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:hc_post_bwd --csv python tools/perf_layer.py --op hc_bwd --iters 3 --warmup 1 --dbg $f 2>&1 | grep hc_post | awk -F'","' '{print $NF}' | tail -2 | tr '\n' ' '
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:hc_post_bwd --csv python tools/perf_layer.py --op hc_bwd --L 180 --C 512 --iters 3 --warmup 1 --dbg $f 2>&1 | grep hc_post | awk -F'","' '{print $NF}' | tail -2 | tr '\n' ' '
echo
done
python bench.py --steps 20 --warmup 3 --no-cpu-baseline | cut -c1-400
