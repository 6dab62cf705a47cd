python tools/gpu_check.py 2>&1 | grep -E "FAIL|EXCEPTION|====|checks took" | head -10
for f in 0 8192 4096; do
echo "flags $f"
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:hc_post_fwd --csv python tools/perf_layer.py --op hc_fwd --iters 3 --warmup 1 --dbg $f 2>&1 | grep hc_post | awk -F'","' '{print $NF}' | tail -2 | tr '\n' ' '
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:hc_post_fwd --csv python tools/perf_layer.py --op hc_fwd --L 180 --C 512 --iters 3 --warmup 1 --dbg $f 2>&1 | grep hc_post | awk -F'","' '{print $NF}' | tail -2 | tr '\n' ' '
echo
done
python bench.py --steps 20 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['achieved'])"
OPH_DEBUG_FLAGS=4096 python bench.py --steps 20 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['achieved'])"
