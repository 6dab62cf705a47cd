python tools/gpu_check.py 2>&1 | grep -E "FAIL|EXCEPTION|====|checks took" 
python tools/perf_layer.py --op attn_fwd --iters 20 | head -3
python bench.py --steps 20 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['gemm_breakdown'])"
