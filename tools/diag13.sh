python -m pytest tests -m gpu -x -q 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:pack_batch -c 3 --csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph 2>&1 | grep pack_batch | awk -F'","' '{print "pack_batch us:", $NF}'
python bench.py --steps 20 --warmup 3 --no-cpu-baseline | cut -c1-400
