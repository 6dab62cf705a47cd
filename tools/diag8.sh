python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 --no-cpu-baseline
python bench.py --workload synth --full-dim 513
python bench.py --workload ssrn_train --steps 5 --warmup 3 --no-cpu-baseline
