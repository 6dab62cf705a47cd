for f in 0 2 128 130; do
python tools/perf_layer.py --op hc_fwd --iters 20 --dbg $f | head -3
python tools/perf_layer.py --op hc_dgrad --iters 20 --dbg $f | head -3
done
