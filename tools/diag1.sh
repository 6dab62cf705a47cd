for op in hc_fwd hc_dgrad hc_bwd conv_fwd; do python tools/perf_layer.py --op $op --iters 20; done
python tools/perf_layer.py --op hc_fwd --L 180 --C 512 --iters 20
python tools/perf_layer.py --op hc_bwd --L 180 --C 512 --iters 20
python tools/perf_layer.py --op attn_fwd --iters 20
