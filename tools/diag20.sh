python tools/gpu_check.py 2>&1 | grep -E "FAIL|EXCEPTION|====|checks took" | head -10
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/perf_layer.py --op hc_fwd --L 180 --C 512 --iters 20 | head -3
python tools/perf_layer.py --op hc_fwd --L 180 --C 512 --iters 20 --dbg 16 | head -3
python bench.py --steps 20 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['achieved'])"
OPH_DEBUG_FLAGS=16 python bench.py --steps 20 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['achieved'])"
