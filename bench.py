#!/usr/bin/env python
"""Benchmark of the dc_tts hot path on B200.  Headline: Text2Mel training (BASELINE.json configs[1]: batch 32 per GPU,
180 phonemes, 870 mel frames, d=256, guided-attention loss, dropout 0.05, synthetic LJ-shape data).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload t2m_train|ssrn_train|synth] [--batch B] [--no-sub]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

One JSON line on rank 0.  `value` = whole-job mel-frames/s with inputs resident in HBM; `e2e` = same metric through
the reference-facing Session.run call with pinned host batches (H2D + loss D2H inside the timed region);
`roofline` = the tcgen05 GEMM core timed per launch with CUDA events inside the timed region;
`cpu_baseline` = the torch-CPU fp32 restatement of the reference's path on this host (N=1 only).
At N=1 the default run also measures BASELINE.json's other single-GPU configurations in sub-processes and attaches
their lines (each with its own roofline / cpu_baseline / e2e) under `sub_results`: the north star's target shape
(batch 64), SSRN training (configs[2]) and autoregressive synthesis (configs[3]).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = "mel-frames/s"
GEMM_TAGS = (0, 1, 2, 3, 4, 7)          # OPH_TAG_*: other, conv_fwd, dgrad, wgrad, attention, hc_fwd
ROW_TAGS = (5, 6, 8)                    # row_fwd, row_bwd, hc_row_fwd
TAG_NAMES = {0: "other", 1: "conv_fwd", 2: "dgrad", 3: "wgrad", 4: "attention", 7: "hc_fwd"}
NUM_TAGS = 10


def metric_name(args):
    if args.workload == "t2m_train":
        return "Text2Mel train mel-frames/sec (B=%d per GPU, N=%d, T=%d)" % (args.batch, args.N, args.T)
    if args.workload == "ssrn_train":
        return "SSRN train coarse mel-frames/sec (B=%d per GPU, T=%d -> %d frames)" % (args.batch, args.T, 4 * args.T)
    return "synthesis RTF (10 sentences, max_N=150, max_T=200, monotonic attention, Griffin-Lim off)"


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tensor_burst=d["bf16_tflops"], tensor_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, src="fallback")


def hc_stack_work(model, B, N, T, c=512, d=256):
    """Algorithmic FLOPs and bytes (per layer 2*k*C*2C*B*L and 8*C*B*L + 4*(k*C*2C + 6C)) of the forward pass
    through all highway-conv layers of a model."""
    if model == "t2m":
        layers = [(2 * d, N, 3)] * 10 + [(2 * d, N, 1)] * 2 + [(d, T, 3)] * 10 + [(d, T, 3)] * 6
    else:
        layers = [(c, T, 3)] * 2 + [(c, 2 * T, 3)] * 2 + [(c, 4 * T, 3)] * 2 + [(2 * c, 4 * T, 3)] * 2
    flops = sum(2.0 * k * C * 2 * C * B * L for C, L, k in layers)
    byts = sum(8.0 * C * B * L + 4.0 * (k * C * 2 * C + 6 * C) for C, L, k in layers)
    return flops, byts, len(layers)


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed regions run."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
def _cpu_setup():
    """All host threads for the CPU arm: torchrun exports OMP_NUM_THREADS=1 to its workers, which would silently pin the
    port to one core."""
    n = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ["MKL_NUM_THREADS"] = str(n)
    import torch
    torch.set_num_threads(n)
    return torch.get_num_threads()


def cpu_train_step(args, model, batch):
    """Builds the torch-CPU fp32 restatement (oracle/dctts_torch.py) of one training step at the benchmark's N / T and
    full_dim with `batch` utterances; returns a zero-argument step function."""
    import numpy as np
    import torch
    from oracle import dctts_torch as ot
    from oracle.params import HP, init_params, ssrn_specs, synthetic_batch, text2mel_specs
    N, T = args.N, args.T
    hp = HP(max_N=N, max_T=T, full_dim=args.full_dim)
    gen = torch.Generator().manual_seed(0)
    if model == "t2m":
        P = ot.to_torch(init_params(text2mel_specs(hp), 0), torch.float32, requires_grad=True)
        b = synthetic_batch(hp, batch, N, T, ragged=True)
        L, mels = torch.tensor(b["L"].astype(np.int64)), torch.tensor(b["mels"])
        opt = ot.TFAdam(hp, P)
        return lambda: ot.text2mel_train_step(hp, P, opt, L, mels, gen)
    P = ot.to_torch(init_params(ssrn_specs(hp), 0), torch.float32, requires_grad=True)
    b = synthetic_batch(hp, batch, 8, T, with_mags=True)
    mels, mags = torch.tensor(b["mels"]), torch.tensor(b["mags"])
    opt = ot.TFAdam(hp, P)
    return lambda: ot.ssrn_train_step(hp, P, opt, mels, mags, gen)


def cpu_reference(args, model, batch, steps, warmup, budget_s=None):
    """The reference's path on host cores: one training step of the port per timed step.  With budget_s the batch is halved
    until warmup + steps fit the budget (judged by the first step).  Returns (frames/s, s per step, threads, sample text,
    batch used)."""
    cores = _cpu_setup()
    while True:
        step = cpu_train_step(args, model, batch)
        t0 = time.perf_counter()
        step()
        first = time.perf_counter() - t0
        if budget_s is None or batch == 1 or first * (steps + max(warmup - 1, 0)) <= budget_s:
            break
        batch = max(1, batch // 2)
    for _ in range(max(warmup - 1, 0)):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    sample = "torch-CPU fp32 restatement, %s train step (fwd+bwd+clip+TF-Adam, dropout 0.05), batch %d x N=%d x T=%d, " \
             "%d timed step(s), %d threads" % (model, batch, args.N, args.T, steps, cores)
    return batch * args.T / dt, dt, cores, sample, batch


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path.  The reference is Python-2 / TensorFlow-1.12
    code that cannot run in this image, so this arm times the CPU restatement (oracle/, kind "port") with every host thread
    at the labelled batch (reduced, and labelled so, only if K steps would not fit a few minutes)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "synth":
        res = cpu_synth_baseline(args, frames=16)
        line = {"impl": "reference", "metric": metric_name(args), "value": res["value"], "unit": "wall s per audio s",
                "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": res["s_per_frame"] * 200 * 1e3,
                "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "encode_text + autoregressive Text2Mel loop + SSRN, B=10, max_N=150, max_T=200"},
                "cpu_baseline": res, "e2e": {"value": res["value"], "unit": "wall s per audio s", "h2d_bytes_per_step": 0,
                                             "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    model = "t2m" if args.workload == "t2m_train" else "ssrn"
    batch = args.ref_batch if args.ref_batch > 0 else (args.batch if model == "t2m" else min(args.batch, 2))
    value, dt, cores, sample, used = cpu_reference(args, model, batch, max(1, args.steps), max(0, min(args.warmup, 2)),
                                                   budget_s=170.0)
    cfg_args = argparse.Namespace(**vars(args))
    cfg_args.batch = used
    line = {"impl": "reference", "metric": metric_name(cfg_args), "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(cfg_args, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference = TensorFlow-1.12/Python-2.7 code that cannot run here; this arm times the CPU "
                    "restatement of the same graph (oracle/, kind=port), batch %d per step%s" %
                    (used, "" if used == args.batch else " (reduced from %d to fit the time budget)" % args.batch)}
    print(json.dumps(line), flush=True)


def cpu_c1_forward():
    """Text2Mel forward at config C1 (B=2, N=60, T=200), median of 20 runs after 3 warm-ups."""
    import numpy as np
    import torch
    from oracle import dctts_torch as ot
    from oracle.params import HP, init_params, synthetic_batch, text2mel_specs
    cores = _cpu_setup()
    hp = HP(max_N=60, max_T=200)
    P = ot.to_torch(init_params(text2mel_specs(hp), 0), torch.float32)
    b = synthetic_batch(hp, 2, 60, 200, text_len=50)
    L, mels = torch.tensor(b["L"].astype(np.int64)), torch.tensor(b["mels"])
    ts = []
    with torch.no_grad():
        for i in range(23):
            t0 = time.perf_counter()
            ot.text2mel_forward(hp, P, L, mels, "generate_attention")
            if i >= 3:
                ts.append(time.perf_counter() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    return {"value": 2 * 200 / med, "unit": "mel-frames/s (forward only)", "seconds": med, "cores": cores, "kind": "port",
            "sample": "BASELINE.md CPU-1: Text2Mel forward, B=2, N=60, T=200, median of 20 runs"}


def cpu_ssrn_forward(full_dim):
    """SSRN forward, B=2, T=200 -> 800."""
    import torch
    from oracle import dctts_torch as ot
    from oracle.params import HP, init_params, ssrn_specs, synthetic_batch
    cores = _cpu_setup()
    hp = HP(full_dim=full_dim)
    P = ot.to_torch(init_params(ssrn_specs(hp), 0), torch.float32)
    mels = torch.tensor(synthetic_batch(hp, 2, 8, 200)["mels"])
    ts = []
    with torch.no_grad():
        for i in range(6):
            t0 = time.perf_counter()
            ot.SSRN(hp, P, mels)
            if i >= 1:
                ts.append(time.perf_counter() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    return {"value": 2 * 200 / med, "unit": "coarse mel-frames/s (forward only)", "seconds": med, "cores": cores,
            "kind": "port", "sample": "BASELINE.md CPU-3: SSRN forward, B=2, T=200 -> 800, F=%d, median of 5 runs" % full_dim}


def cpu_synth_baseline(args, frames=16):
    """the reference-style autoregressive loop (the WHOLE graph over all max_T rows per generated frame,
    synthesize.py:181-228) on the host, bounded to the first `frames` frames of the 10 x N=150 x T=200 workload; the cost
    per frame does not depend on the frame index, so RTF = (encode + 200 * seconds per frame + SSRN) / audio seconds."""
    import numpy as np
    import torch
    from oracle import dctts_torch as ot
    from oracle.params import HP, init_params, ssrn_specs, text2mel_specs
    cores = _cpu_setup()
    nsent, maxN, maxT = 10, 150, 200
    hp = HP(max_N=maxN, max_T=maxT, full_dim=args.full_dim)
    P = ot.to_torch(init_params(text2mel_specs(hp), 0), torch.float32)
    P2 = ot.to_torch(init_params(ssrn_specs(hp), 1), torch.float32)
    rng = np.random.default_rng(1234)
    L = np.zeros((nsent, maxN), np.int64)
    for i in range(nsent):
        n = int(rng.integers(60, maxN - 1))
        L[i, :n] = rng.integers(1, len(hp.vocab), n)
    with torch.no_grad():
        t0 = time.perf_counter()
        K, V = ot.TextEnc(hp, P, torch.tensor(L))
        t_enc = time.perf_counter() - t0
        hp_s = HP(max_N=maxN, max_T=maxT, full_dim=args.full_dim)
        Y = torch.zeros(nsent, maxT, hp.n_mels)
        prev = torch.zeros(nsent, dtype=torch.long)
        t0 = time.perf_counter()
        for j in range(frames):
            out = ot.text2mel_forward(hp_s, P, None, Y, "synthesize", prev, K=K, V=V)
            Y[:, j] = out["Y"][:, j]
            prev = out["max_attentions"][:, j]
        t_frame = (time.perf_counter() - t0) / frames
        t0 = time.perf_counter()
        ot.SSRN(hp, P2, Y[:2])
        t_ssrn = (time.perf_counter() - t0) * nsent / 2
    audio = nsent * maxT * hp.r * 275 / 22050.0
    wall = t_enc + t_frame * maxT + t_ssrn
    return {"value": wall / audio, "unit": "wall s per audio s", "cores": cores, "kind": "port", "s_per_frame": t_frame,
            "sample": "BASELINE.md CPU-4: reference-style loop (whole graph per frame), 10 sentences, max_N=150, max_T=200; "
                      "%d of 200 frames timed (%.3f s each, constant cost per frame), encode %.2f s, SSRN on 2 of 10 "
                      "sentences scaled x5 (%.2f s)" % (frames, t_frame, t_enc, t_ssrn)}


def workload_config(args, world):
    if args.workload == "t2m_train":
        wl = "Text2Mel training step (TextEnc+AudioEnc+Attention+AudioDec fwd, L1+BD+guided-attention loss, bwd, " \
             "clip+Adam): batch %d per GPU, %d phonemes, %d mel frames, d=256, dropout 0.05" % (args.batch, args.N, args.T)
    else:
        wl = "SSRN training step: batch %d per GPU, %d->%d frames, 80 mels -> %d bins" % (args.batch, args.T, 4 * args.T, args.full_dim)
    return {"workload": wl, "global_batch": args.batch * world, "parallelism": "dp%d" % world,
            "l2": "per-step working set (saved activations, several GB) exceeds the 126 MB L2; no explicit flush",
            "precision": "fp32 I/O; GEMMs as 3-term split-bf16 on tcgen05 (fp32-grade, 3 tensor passes per algorithmic FLOP)"}


# ------------------------------------------------------------------------------------------------ GPU arm: training
def run_ours(args):
    import ctypes
    import torch
    import torch.distributed as dist
    import __graft_entry__
    from ophelia_b200 import _lib
    from ophelia_b200.architectures import SSRNGraph, Text2MelGraph
    from ophelia_b200.configuration import default_hparams
    from ophelia_b200.data import SyntheticBatches
    from ophelia_b200.parallel import init_from_env
    from ophelia_b200.session import Session
    from ophelia_b200.variables import VariableStore

    rank, world, local = init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    lib = _lib.load()
    group = dist.group.WORLD if world > 1 else None
    t2m = args.workload == "t2m_train"
    hp = default_hparams(max_N=args.N, max_T=args.T, full_dim=args.full_dim, seed=0)
    hp.overlap_allreduce = int(args.overlap) if world > 1 else 0
    hp.allreduce_chunks = args.ar_chunks
    src = SyntheticBatches(hp, "t2m" if t2m else "ssrn", args.batch, N=args.N, T=args.T, seed=1234 + rank)
    store = VariableStore(dev, seed=0)
    Graph = Text2MelGraph if t2m else SSRNGraph
    g = Graph(hp, mode="train", store=store, data=src, device=dev, process_group=group)
    sess = Session()
    frames_per_step = args.batch * args.T * world

    b0 = src.batches[0]
    if t2m:
        dev_in = (b0["text"].to(dev), b0["mel"].to(dev))
    else:
        dev_in = (b0["mel"].to(dev), b0["mag"].to(dev))

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        sync_all()
        return e0.elapsed_time(e1) / steps, out

    for _ in range(max(args.warmup, 3)):
        g.train_step_device(*dev_in)
    for _ in range(5):      # eager twice, then the step is captured into a CUDA graph and replayed
        sess.run([g.global_step, g.loss_components, g.train_op])
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()

    # ---- timed region 1: device-resident inputs, eager launches on ONE stream with CUDA events around every GEMM
    #      launch (the per-kernel roofline numbers; side streams would make the event intervals overlap)
    hp.use_side_streams = False
    saved_overlap, hp.overlap_allreduce = hp.overlap_allreduce, False
    for _ in range(2):
        g.train_step_device(*dev_in)
    sync_all()
    launches0 = lib.oph_launch_count()
    lib.oph_profile_begin()
    ms, comps = timed(lambda: g.train_step_device(*dev_in), args.steps)
    prof = (ctypes.c_double * (NUM_TAGS * 3))()
    lib.oph_profile_end(prof)
    hp.use_side_streams = True
    hp.overlap_allreduce = saved_overlap
    launches = (lib.oph_launch_count() - launches0) // args.steps
    last_loss = [float(c) for c in comps.cpu().numpy()]

    # ---- timed region 1b: the same steps replayed from one CUDA graph (kernel-for-kernel identical work)
    ms_graph = None
    ms_nocomm = None
    if not args.no_graph:
        step = g.capture_train_step(*dev_in)
        for _ in range(3):
            step(*dev_in)
        ms_graph, comps = timed(lambda: step(*dev_in), args.steps)
        last_loss = [float(c) for c in comps.cpu().numpy()]
        if world > 1:
            # the same captured step without the gradient exchange: the difference is the collective time that the
            # backward pass does not hide
            pg, g.process_group = g.process_group, None
            g.__dict__.pop("_buckets", None)
            step_nc = g.capture_train_step(*dev_in)
            for _ in range(3):
                step_nc(*dev_in)
            ms_nocomm, _ = timed(lambda: step_nc(*dev_in), args.steps)
            g.process_group = pg
            g.__dict__.pop("_buckets", None)
            del step_nc

    # ---- timed region 2: end to end through Session.run with pinned host batches
    ms_e2e, fetched = timed(lambda: sess.run([g.global_step, g.loss_components, g.train_op]), args.steps)
    gs = fetched[0]
    clk = clocks.stop() if rank == 0 else None

    if world > 1:
        t = torch.tensor([ms, ms_e2e, ms_graph if ms_graph is not None else 0.0, ms_nocomm if ms_nocomm is not None else 0.0],
                         device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        ms_graph = float(t[2]) if ms_graph is not None else None
        ms_nocomm = float(t[3]) if ms_nocomm is not None else None
    ms_eager = ms
    if ms_graph is not None and ms_graph < ms:
        ms = ms_graph                      # headline: the faster of the two launch modes over the same kernels
    if rank != 0:
        if world > 1:
            dist.barrier()
        return

    pk = peaks()
    P = lambda tag, j: prof[tag * 3 + j]   # noqa: E731
    tot_n = sum(P(t, 0) for t in GEMM_TAGS); tot_ms = sum(P(t, 1) for t in GEMM_TAGS); tot_fl = sum(P(t, 2) for t in GEMM_TAGS)
    achieved = tot_fl / (tot_ms * 1e-3) / 1e12 if tot_ms > 0 else 0.0
    row_n = sum(P(t, 0) for t in ROW_TAGS); row_ms = sum(P(t, 1) for t in ROW_TAGS); row_by = sum(P(t, 2) for t in ROW_TAGS)
    row_gbs = row_by / (row_ms * 1e-3) / 1e9 if row_ms > 0 else 0.0
    fwd_ms, fwd_by = P(5, 1) + P(8, 1), P(5, 2) + P(8, 2)
    breakdown = {TAG_NAMES[t]: {"launches_per_step": P(t, 0) / args.steps, "ms_per_step": P(t, 1) / args.steps,
                                "tflops": (P(t, 2) / (P(t, 1) * 1e-3) / 1e12) if P(t, 1) > 0 else 0.0}
                 for t in GEMM_TAGS if P(t, 0) > 0}
    if "attention" in breakdown:       # the attention block is HBM-bound (AI 58-75 FLOP/B); report that fraction too
        att_bytes = 3.0 * 4.0 * args.batch * (2.0 * args.T * 256 + 2.0 * args.N * 256)      # fwd + bwd, A never counted
        att = breakdown["attention"]
        att["algorithmic_gbs"] = att_bytes / (att["ms_per_step"] * 1e-3) / 1e9
        att["frac_of_hbm_peak"] = att["algorithmic_gbs"] / pk["hbm"]
        att["frac_of_tensor_peak"] = att["tflops"] / pk["tensor_sustained"]
    # the highway-conv stack of the forward pass (north star: fused HighwayConv1d stack against the HBM roofline): GEMM
    # launches of the highway layers + their tail launches (none where the tail runs in the GEMM's epilogue)
    hc_fl, hc_by, hc_layers = hc_stack_work("t2m" if t2m else "ssrn", args.batch, args.N, args.T)
    hc_ms = (P(7, 1) + P(8, 1)) / args.steps
    hc_stack = None
    if hc_ms > 0:
        hc_stack = {"layers": hc_layers, "launches_per_step": (P(7, 0) + P(8, 0)) / args.steps, "ms_forward": hc_ms,
                    "tail_launches_per_step": P(8, 0) / args.steps, "tail_ms": P(8, 1) / args.steps,
                    "algorithmic_gflop": hc_fl / 1e9, "algorithmic_mb": hc_by / 1e6,
                    "hbm": {"achieved_gbs": hc_by / (hc_ms * 1e-3) / 1e9, "frac": hc_by / (hc_ms * 1e-3) / 1e9 / pk["hbm"],
                            "ceiling_3pass": (hc_by / (pk["hbm"] * 1e9)) / (3.0 * hc_fl / (pk["tensor_burst"] * 1e12))},
                    "tensor": {"achieved_tflops": hc_fl / (hc_ms * 1e-3) / 1e12,
                               "frac": hc_fl / (hc_ms * 1e-3) / 1e12 / pk["tensor_sustained"], "ceiling_3pass": 1.0 / 3.0},
                    "note": "forward pass of a training step (z and the row statistics are saved); the stack is tensor-bound: "
                            "its HBM fraction cannot exceed the 3-pass ceiling (time at the tensor peak x 3 passes)"}
    traffic = ncu_traffic(args) if t2m else {}
    line = {
        "metric": metric_name(args), "value": frames_per_step / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
        "per_gpu": frames_per_step / world / (ms * 1e-3),
        "launch_mode": {"headline": "cuda_graph" if ms is ms_graph else "eager", "ms_per_step_eager": ms_eager,
                        "ms_per_step_cuda_graph": ms_graph,
                        "gemm_ms_over_graph_step": (tot_ms / args.steps) / ms_graph if ms_graph else None,
                        "note": "roofline/gemm_breakdown are CUDA-event timings of every GEMM launch in the eager single-stream "
                                "steps; the CUDA-graph steps run the same kernels with TextEnc and the weight-gradient GEMMs "
                                "on side streams"},
        "e2e": {"value": frames_per_step / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": src.bytes_per_batch(), "d2h_bytes_per_step": 4 * len(last_loss) + 8},
        "gpu_launches": int(launches) * args.steps,
        "gpu_launches_per_step": int(launches),
        "roofline": {"kernel": "gemm_bf16x3_kernel (tcgen05 implicit GEMM: conv fwd / dgrad / wgrad / attention)",
                     "bound": "tensor", "achieved": achieved, "peak": pk["tensor_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / pk["tensor_sustained"], "traffic": traffic.get("gemm"),
                     "traffic_detail": traffic.get("detail"),
                     "peak_source": pk["src"] + " bf16 cuBLAS sustained",
                     "launches_per_step": tot_n / args.steps, "avg_launch_ms": tot_ms / max(tot_n, 1),
                     "share_of_step": (tot_ms / args.steps) / ms_eager,
                     "frac_of_3pass_ceiling": 3.0 * achieved / pk["tensor_sustained"],
                     "note": "achieved counts ALGORITHMIC FLOPs once; the split-bf16 scheme issues 3 tensor passes per "
                             "FLOP, so the ceiling of this number is peak/3"},
        "roofline_hbm": {"kernel": "row-wise LayerNorm / highway tails (hc_post_fwd/bwd_wide, ln_act_fwd/bwd)", "bound": "hbm",
                         "achieved": row_gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": row_gbs / pk["hbm"], "traffic": traffic.get("rowwise"),
                         "launches_per_step": row_n / args.steps, "ms_per_step": row_ms / args.steps,
                         "forward_gbs": (fwd_by / (fwd_ms * 1e-3) / 1e9) if fwd_ms > 0 else 0.0,
                         "backward_gbs": (P(6, 2) / (P(6, 1) * 1e-3) / 1e9) if P(6, 1) > 0 else 0.0,
                         "note": "algorithmic bytes (fp32 in/out + operand planes) over CUDA-event time, eager single stream"},
        "gemm_breakdown": breakdown, "hc_stack": hc_stack,
        "clocks": clk, "loss_components": last_loss, "global_step": int(gs),
    }
    if world > 1 and ms_nocomm is not None and ms_graph is not None:
        line["collective"] = {"ms_per_step_without_exchange": ms_nocomm, "exposed_ms_per_step": ms_graph - ms_nocomm,
                              "overlap": int(hp.overlap_allreduce), "chunks": int(hp.allreduce_chunks), "bytes": int(store.numel) * 4,
                              "note": "same captured step with and without the gradient all-reduce, max over ranks"}
    if world == 1 and not args.no_cpu_baseline:
        # ~10-20 s of host work: timed steps of the port at a bounded batch (one SSRN step is already ~10x a Text2Mel one)
        nb = args.ref_batch if args.ref_batch > 0 else 4
        v, dt, cores, sample, _ = cpu_reference(args, "t2m" if t2m else "ssrn", nb if t2m else min(nb, 2), 20 if t2m else 2, 1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "s_per_step": dt}
        if not t2m:
            line["cpu_forward"] = cpu_ssrn_forward(args.full_dim)
        elif not args.sub:
            line["cpu_forward"] = cpu_c1_forward()
    if world == 1 and not args.sub and not args.no_sub and t2m and args.batch == 32:
        del g, sess, store, src
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        line["sub_results"] = run_subs(args)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()


def run_subs(args):
    """BASELINE.json's other single-GPU configurations, one sub-process each (fresh CUDA context, bounded steps)."""
    base = [sys.executable, os.path.abspath(__file__), "--sub", "--gpus", "1", "--warmup", "3", "--full-dim", str(args.full_dim)]
    jobs = {
        "t2m_b64": base + ["--workload", "t2m_train", "--batch", "64", "--steps", "10"],
        "ssrn_train": base + ["--workload", "ssrn_train", "--batch", "32", "--steps", "5"],
        "synth": base + ["--workload", "synth"],
    }
    out = {}
    for name, cmd in jobs.items():
        t0 = time.time()
        try:
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=420)
            lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            out[name] = json.loads(lines[-1]) if lines else {"error": "no JSON line (rc %d): %s" % (r.returncode, r.stderr[-400:])}
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": repr(e)}
        out[name]["wall_s"] = time.time() - t0
    return out


def ncu_traffic(args):
    """DRAM bytes per launch of the dominant kernels from the committed `ncu --set full` captures (profiles/*_traffic.json,
    the newest round wins).  The captures were taken at the headline shape (B=32, N=180, T=870): other shapes report null."""
    if (args.batch, args.N, args.T) != (32, 180, 870):
        return {}
    pdir = os.path.join(ROOT, "profiles")
    cands = sorted(f for f in os.listdir(pdir) if f.endswith("_traffic.json")) if os.path.isdir(pdir) else []
    if not cands:
        return {}
    path = os.path.join(pdir, cands[-1])
    with open(path) as f:
        cap = json.load(f)["captures"]
    tot = lambda k: cap[k]["dram_read"] + cap[k]["dram_write"]      # noqa: E731
    return {"gemm": tot("gemm_hc_fwd"), "rowwise": tot("hc_post_bwd") if "hc_post_bwd" in cap else None,
            "detail": {"unit": "bytes per launch, one ncu --set full capture each (cold caches)", "file": "profiles/" + cands[-1],
                       "launches": {k: {"dram": tot(k), "algorithmic": v["algorithmic"], "launch": v["launch"]}
                                    for k, v in cap.items()},
                       "note": "roofline.traffic is the highway-conv forward launch (the most frequent launch); "
                               "roofline_hbm.traffic is the highway-tail backward kernel"}}


# ------------------------------------------------------------------------------------------------ GPU arm: synthesis
def run_synth(args):
    """BASELINE.json configs[3]: synthesize.py autoregressive inference, 10 sentences, max_N=150, max_T=200, monotonic
    attention on, Griffin-Lim off: encode_text + AR loop + SSRN.  RTF = wall seconds / audio seconds."""
    import ctypes
    import numpy as np
    import torch
    import __graft_entry__
    from ophelia_b200 import _lib
    from ophelia_b200 import synthesize as syn
    from ophelia_b200.architectures import SSRNGraph, Text2MelGraph
    from ophelia_b200.configuration import default_hparams
    from ophelia_b200.session import Session
    from ophelia_b200.variables import VariableStore
    __graft_entry__.build()
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    nsent, maxN, maxT = 10, 150, 200
    hp = default_hparams(max_N=maxN, max_T=maxT, full_dim=args.full_dim, seed=0)
    rng = np.random.default_rng(1234)
    L = np.zeros((nsent, maxN), np.int32)
    for i in range(nsent):
        n = int(rng.integers(60, maxN - 1))
        L[i, :n] = rng.integers(1, len(hp.vocab), n)
    g1 = Text2MelGraph(hp, mode="synthesize", store=VariableStore(dev, seed=0), device=dev)
    g2 = SSRNGraph(hp, mode="synthesize", store=VariableStore(dev, seed=1), device=dev)
    sess = Session()
    ends = syn.get_text_lengths(L)
    sec_per_frame = hp.r * hp.hop_length / float(hp.sr)

    def route(kind):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        K, V = syn.encode_text(hp, L, g1, sess)
        if kind == "session":
            Y, t_ends, _ = syn.synth_codedtext2mel(hp, K, V, ends, g1, sess)
        elif kind == "incremental":
            Y, t_ends, _ = syn.synth_codedtext2mel_incremental(hp, K, V, ends, g1)
        else:
            Y, t_ends, _ = syn.synth_codedtext2mel_device(hp, K, V, ends, g1, use_cuda_graph=(kind == "device_graph"))
        t1 = time.perf_counter()
        Z = syn.synth_mel2mag(hp, Y, g2, sess)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        audio = sum(t_ends) * sec_per_frame
        return {"wall_s": t2 - t0, "text2mel_s": t1 - t0, "ssrn_s": t2 - t1, "audio_s": audio, "rtf": (t2 - t0) / audio,
                "frames": int(sum(t_ends)), "mag_shape": list(Z.shape)}
    route("device_graph")                                # warm-up (packing, graph pools)
    route("incremental")
    clocks = ClockSampler(0)
    clocks.start()
    launches0 = lib.oph_launch_count()
    res = {k: route(k) for k in ("session", "device", "device_graph", "incremental")}
    launches = lib.oph_launch_count() - launches0
    clk = clocks.stop()
    best = min((res["device_graph"], res["incremental"]), key=lambda r: r["rtf"])
    # per-launch timing of one eager pass of the incremental route (profile tags; a graph replay cannot be timed per launch)
    K, V = syn.encode_text(hp, L, g1, sess)
    lib.oph_profile_begin()
    syn.synth_codedtext2mel_incremental(hp, K, V, ends, g1, use_cuda_graph=False)
    prof = (ctypes.c_double * (NUM_TAGS * 3))()
    lib.oph_profile_end(prof)
    pk = peaks()
    P = lambda tag, j: prof[tag * 3 + j]   # noqa: E731
    gemm_n = sum(P(t, 0) for t in GEMM_TAGS); gemm_ms = sum(P(t, 1) for t in GEMM_TAGS); gemm_fl = sum(P(t, 2) for t in GEMM_TAGS)
    enc_gbs = P(9, 2) / (P(9, 1) * 1e-3) / 1e9 if P(9, 1) > 0 else 0.0
    gemm_tf = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    line = {"metric": metric_name(args),
            "value": best["rtf"], "unit": "wall s per audio s", "n_gpus": 1, "steps": 1, "warmup": 1,
            "ms_per_step": best["wall_s"] * 1e3, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic (random-init weights: attention never reaches the sentence end, so every "
                                    "sentence runs all max_T frames)",
            "config": {"workload": "encode_text + autoregressive Text2Mel loop (synthesize.py:150-230) + SSRN, B=10, "
                                   "max_N=150, max_T=200, F=%d" % args.full_dim,
                       "headline_route": "incremental" if best is res["incremental"] else "device_graph",
                       "routes": "session / device / device_graph re-run the full graph per frame like the reference; "
                                 "incremental caches the AudioEnc rows and re-runs Attention + AudioDec over the "
                                 "decoder's 84-frame causal reach (same results up to fp32 rounding)"},
            "routes": res,
            "e2e": {"value": res["session"]["rtf"], "unit": "wall s per audio s",
                    "h2d_bytes_per_step": int(2 * nsent * maxN * hp.d * 4 + nsent * maxT * hp.n_mels * 4 + nsent * 4),
                    "d2h_bytes_per_step": int(nsent * maxT * hp.n_mels * 4 + nsent * maxT * 4 + nsent * maxN * maxT * 4),
                    "note": "Session.run route: K, V, mels, prev_max_attentions fed and Y, max_attentions, alignments fetched "
                            "as numpy arrays every frame like the reference (synthesize.py:181-183); bytes are per frame"},
            "gpu_launches": int(launches), "clocks": clk,
            "roofline": {"kernel": "ar_encoder_kernel (13 AudioEnc layers of one frame step, fp32 weight streaming, one "
                                   "8-CTA cluster per sentence)", "bound": "hbm", "achieved": enc_gbs, "peak": pk["hbm"],
                         "unit": "GB/s", "frac": enc_gbs / pk["hbm"], "traffic": None,
                         "launches": P(9, 0), "avg_launch_ms": P(9, 1) / max(P(9, 0), 1),
                         "note": "algorithmic bytes = the fp32 kernels of the 13 layers (16.4 MB) per launch, counted once "
                                 "although each of the 10 sentence clusters streams them (L2-resident after the first)"},
            "roofline_window_gemm": {"kernel": "gemm_bf16x3_kernel on the 85-row Attention + AudioDec window", "bound": "tensor",
                                     "achieved": gemm_tf, "peak": pk["tensor_sustained"], "unit": "TFLOP/s",
                                     "frac": gemm_tf / pk["tensor_sustained"], "launches_per_frame": gemm_n / maxT,
                                     "ms_per_frame": gemm_ms / maxT,
                                     "note": "850 rows per launch = 4 work units on 74 CTA pairs: launch-latency bound"}}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_synth_baseline(args, frames=12)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="t2m_train", choices=["t2m_train", "ssrn_train", "synth"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--N", type=int, default=180)
    ap.add_argument("--T", type=int, default=870)
    ap.add_argument("--full-dim", dest="full_dim", type=int, default=513)
    ap.add_argument("--ref-batch", dest="ref_batch", type=int, default=0,
                    help="batch of the CPU port (0: the labelled batch for --impl reference, 4 for the cpu_baseline sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="skip the CUDA-graph replay of the training step")
    ap.add_argument("--no-sub", action="store_true", help="skip the sub-process runs of the other BASELINE configurations")
    ap.add_argument("--sub", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--overlap", type=int, default=0,
                    help="data parallel: 0 = one all-reduce after the backward pass (default), 1 = gradient buckets are all-reduced "
                         "on a communication stream while the backward pass still runs, 2 = 'late' buckets: the highway layers + "
                         "decoder (98 %% of the bytes) are exchanged under the small launches that end the backward pass "
                         "(measured at 2 GPUs: 6.82 vs 6.88 ms per step)")
    ap.add_argument("--ar-chunks", dest="ar_chunks", type=int, default=1,
                    help="data parallel: the flat gradient is all-reduced in this many slices after the backward pass, the fused "
                         "clip + Adam kernel of slice i overlapping the exchange of slice i+1 (1 = one collective, then Adam; measured at 2 GPUs: "
                         "6.98 / 7.05 / 7.01 / 7.13 ms per step with 1 / 2 / 4 / 8 slices, so 1 is the default)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "synth":
        return run_synth(args)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29411", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
