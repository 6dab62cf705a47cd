#!/bin/bash
# Run ON THE GPU BOX (via gpurun): ncu launch list of the bench command + full captures of the dominant kernels.
# Every .ncu-rep is summarised on the box (tools/ncu_summary.py -> gpurun_out/<R>_<name>_ncu.txt) and then deleted: the
# reports are ~10 MB each and gpurun_out/ is only copied back below 64 MiB.  Copy the summaries into profiles/ afterwards.
set -u
R=${1:-r02}
cap() {   # cap <name> <kernel regex> <launch skip> <command...>
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/${R}_${name} "$@" > gpurun_out/ncu_${R}_${name}.log 2>&1
  python tools/ncu_summary.py gpurun_out/${R}_${name}.ncu-rep > gpurun_out/${R}_${name}_ncu.txt 2>&1
  rm -f gpurun_out/${R}_${name}.ncu-rep
}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_${R}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-sub > gpurun_out/bench_under_ncu_${R}.log 2>&1
python tools/launch_shares.py gpurun_out/launches_${R}.csv > gpurun_out/${R}_step_kernel_shares.txt 2>&1
# highway-conv forward: the plain conv launch + separate tail at B=32 (109 row tiles: not fused) and the one-launch form at B=64
cap gemm_hc_fwd gemm_bf16x3 3 python tools/perf_layer.py --op hc_fwd --iters 2
cap gemm_hc_dgrad gemm_bf16x3 3 python tools/perf_layer.py --op hc_dgrad --iters 2
cap gemm_hc_bwd gemm_bf16x3 7 python tools/perf_layer.py --op hc_bwd --iters 2
cap gemm_hc_fused_fwd gemm_bf16x3 3 python tools/perf_layer.py --op hc_fwd --B 64 --iters 2
cap hc_post_bwd hc_post_bwd_wide 2 python tools/perf_layer.py --op hc_bwd --iters 2
cap hc_post_fwd hc_post_fwd_wide 2 python tools/perf_layer.py --op hc_fwd --iters 2
# the one-kernel attention forward (32 items x 870 queries x 180 keys)
cap attn_fused attn_fused 2 python tools/perf_layer.py --op attn_fwd --iters 2
# the TextEnc-sized row kernels (5760 rows x 512 channels): short launches, fixed costs show
cap hc_post_bwd_text hc_post_bwd_wide 2 python tools/perf_layer.py --op hc_bwd --L 180 --C 512 --iters 2
cap hc_post_fwd_text hc_post_fwd_wide 2 python tools/perf_layer.py --op hc_fwd --L 180 --C 512 --iters 2
# incremental autoregressive route: launch list of ~2 frame steps (eager) and per-frame timing of the graph replays
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3000 -c 70 --csv \
    --log-file gpurun_out/launches_ar_${R}.csv python tools/ar_probe.py --frames 120 > gpurun_out/ar_probe_ncu_${R}.log 2>&1
timeout 120 python tools/ar_probe.py --graph > gpurun_out/ar_probe_${R}.log 2>&1
# in-kernel time stamps of every GEMM launch of one graph replay + the kernel timeline of that replay
timeout 200 python tools/gemm_ring.py --out ${R}_gemm_ring.txt > /dev/null 2>&1
timeout 200 python tools/timeline.py --tag ${R} > /dev/null 2>&1
# per-launch table of one eager step (CUDA events around every launch, shapes from the launch code)
OPH_PROF_DUMP=gpurun_out/prof_${R}.txt python bench.py --no-graph --no-sub --no-cpu-baseline --steps 1 --warmup 3 > /dev/null 2>&1
python tools/launch_table.py gpurun_out/prof_${R}.txt > gpurun_out/${R}_launch_table.txt 2>&1; rm -f gpurun_out/prof_${R}.txt
du -sh gpurun_out; ls gpurun_out | tail -30
