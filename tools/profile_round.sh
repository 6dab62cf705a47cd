#!/bin/bash
# Run ON THE GPU BOX (via gpurun): ncu launch list of the bench command + full captures of the dominant kernels.
# Outputs land in gpurun_out/; summarise them afterwards with tools/ncu_summary.py / tools/launch_shares.py into profiles/.
set -u
R=${1:-r02}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_${R}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-sub > gpurun_out/bench_under_ncu_${R}.log 2>&1
# highway-conv forward: the plain conv launch + separate tail at B=32 (109 row tiles: not fused) and the one-launch form at B=64
for op in hc_fwd hc_dgrad hc_bwd; do
  skip=3; [ $op = hc_bwd ] && skip=7
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s $skip -c 1 -f \
      -o gpurun_out/${R}_gemm_${op} python tools/perf_layer.py --op $op --iters 2 > gpurun_out/ncu_${R}_${op}.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s 3 -c 1 -f \
    -o gpurun_out/${R}_gemm_hc_fused_fwd python tools/perf_layer.py --op hc_fwd --B 64 --iters 2 > gpurun_out/ncu_${R}_hc_fused.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hc_post_bwd_wide -s 2 -c 1 -f -o gpurun_out/${R}_hc_post_bwd \
    python tools/perf_layer.py --op hc_bwd --iters 2 > gpurun_out/ncu_${R}_post.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hc_post_fwd_wide -s 2 -c 1 -f -o gpurun_out/${R}_hc_post_fwd \
    python tools/perf_layer.py --op hc_fwd --iters 2 > gpurun_out/ncu_${R}_postf.log 2>&1
# the one-kernel attention forward (32 items x 870 queries x 180 keys)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fused -s 2 -c 1 -f -o gpurun_out/${R}_attn_fused \
    python tools/perf_layer.py --op attn_fwd --iters 2 > gpurun_out/ncu_${R}_attn.log 2>&1
# incremental autoregressive route: launch list of ~2 frame steps (eager) and per-frame timing of the graph replays
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3000 -c 70 --csv \
    --log-file gpurun_out/launches_ar_${R}.csv python tools/ar_probe.py --frames 120 > gpurun_out/ar_probe_ncu_${R}.log 2>&1
timeout 120 python tools/ar_probe.py --graph > gpurun_out/ar_probe_${R}.log 2>&1
ls -la gpurun_out/ | tail -12
