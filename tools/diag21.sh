python bench.py --workload ssrn_train --steps 6 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ssrn513', d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['roofline']['achieved'], d['gemm_breakdown'])"
python bench.py --workload ssrn_train --full-dim 1025 --steps 6 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ssrn1025', d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['roofline']['achieved'], d['gemm_breakdown'])"
python bench.py --batch 64 --steps 10 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('t2m B64', d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['roofline']['achieved'], d['gemm_breakdown'])"
