python tools/perf_layer.py --op attn_fwd --iters 20 --dbg 0 2>/dev/null | head -4
python tools/perf_layer.py --op attn_fwd --iters 20 --dbg 1 2>/dev/null | head -4
python tools/perf_layer.py --op attn_fwd --iters 20 --B 8 2>/dev/null | head -4
python tools/perf_layer.py --op attn_fwd --iters 20 --B 64 2>/dev/null | head -4
