python tools/gpu_check.py 2>&1 | grep -E "FAIL|EXCEPTION|====|checks took" | head -40
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['gemm_breakdown'])"
python bench.py --workload synth --full-dim 513 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:(round(v['wall_s'],4), round(v['rtf'],5)) for k,v in d['routes'].items()})"
