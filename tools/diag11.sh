python tools/gpu_check.py 2>&1 | grep -E "FAIL|EXCEPTION|====|checks took" 
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 --no-cpu-baseline | cut -c1-400
OPH_DEBUG_FLAGS=64 python bench.py --steps 20 --warmup 3 --no-cpu-baseline | cut -c1-400
python bench.py --steps 20 --warmup 3 --no-cpu-baseline | cut -c1-400
