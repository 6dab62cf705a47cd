python tools/gpu_check.py 2>&1 | grep -E "FAIL|EXCEPTION|====|checks took|attention" | head -40
python tools/perf_layer.py --op attn_fwd --iters 20 2>/dev/null | head -4
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['gemm_breakdown'])"
