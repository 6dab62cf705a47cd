python tools/gpu_check.py 2>&1 | grep -E "FAIL|EXCEPTION|====|checks took" 
for f in 0 32; do
for op in hc_fwd hc_dgrad; do python tools/perf_layer.py --op $op --iters 20 --dbg $f; done
python tools/perf_layer.py --op hc_fwd --L 180 --C 512 --iters 20 --dbg $f
python tools/perf_layer.py --op hc_dgrad --L 180 --C 512 --iters 20 --dbg $f
done
python tools/perf_layer.py --op conv_fwd --iters 20
python bench.py --steps 20 --warmup 3 --no-cpu-baseline | cut -c1-400
