"""Kernel timeline of ONE replay of the captured training step (torch.profiler / CUPTI activity records; diagnostics
only, numbers taken under the profiler are never bench values).  Writes gpurun_out/timeline_<tag>.txt: one line per
kernel (start us, duration us, stream, grid, name) and a summary (span, union of busy time, idle gaps, per-stream sums).

    python tools/timeline.py [--batch 32] [--tag r02] [--workload t2m_train]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--N", type=int, default=180)
    ap.add_argument("--T", type=int, default=870)
    ap.add_argument("--tag", default="r02")
    ap.add_argument("--workload", default="t2m_train")
    ap.add_argument("--flags", type=int, default=0, help="oph_gemm_debug_flags value for A/B runs")
    ap.add_argument("--cache", type=int, default=0, help="oph_cache_config mode for A/B runs")
    args = ap.parse_args()
    import torch
    from torch.profiler import ProfilerActivity, profile
    import __graft_entry__
    __graft_entry__.build()
    from ophelia_b200 import _lib
    from ophelia_b200.architectures import SSRNGraph, Text2MelGraph
    from ophelia_b200.configuration import default_hparams
    from ophelia_b200.data import SyntheticBatches
    from ophelia_b200.variables import VariableStore
    lib = _lib.load()
    if args.flags:
        _lib.set_debug_flags(args.flags)
    dev = torch.device("cuda", 0)
    torch.cuda.init()
    if args.cache:
        lib.oph_cache_config(args.cache)
    t2m = args.workload == "t2m_train"
    hp = default_hparams(max_N=args.N, max_T=args.T, seed=0)
    src = SyntheticBatches(hp, "t2m" if t2m else "ssrn", args.batch, N=args.N, T=args.T, seed=1234)
    store = VariableStore(dev, seed=0)
    g = (Text2MelGraph if t2m else SSRNGraph)(hp, mode="train", store=store, data=src, device=dev)
    b0 = src.batches[0]
    dev_in = (b0["text"].to(dev), b0["mel"].to(dev)) if t2m else (b0["mel"].to(dev), b0["mag"].to(dev))
    for _ in range(3):
        g.train_step_device(*dev_in)
    step = g.capture_train_step(*dev_in)
    for _ in range(5):
        step(*dev_in)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step(*dev_in)
    e1.record()
    torch.cuda.synchronize()
    ms_plain = e0.elapsed_time(e1) / 10
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            step(*dev_in)
        torch.cuda.synchronize()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    trace = os.path.join(ROOT, "gpurun_out", "timeline_%s_trace.json" % args.tag)
    prof.export_chrome_trace(trace)
    ev = [e for e in json.load(open(trace))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
    os.remove(trace)
    ev.sort(key=lambda e: e["ts"])
    # split into replays: a gap after the last kernel of a step (adam / pack / step_inc) -- use the graph id if present
    starts = [i for i, e in enumerate(ev) if "step_inc" in e["name"]]
    # one replay = events between the 1st and the 2nd step_inc kernel (step_inc is the last launch of a step)
    if len(starts) >= 2:
        ev = ev[starts[0] + 1:starts[1] + 1]
    t0 = ev[0]["ts"]
    out = os.path.join(ROOT, "gpurun_out", "timeline_%s.txt" % args.tag)
    with open(out, "w") as f:
        f.write("# ms per step without the profiler: %.3f ; kernels in one replay: %d\n" % (ms_plain, len(ev)))
        busy_end, busy, gaps = t0, 0.0, []
        per_stream = {}
        for e in ev:
            s, d = e["ts"] - t0, e["dur"]
            a = e.get("args", {})
            name = e["name"].replace("void ", "").replace("oph::", "")[:60]
            f.write("%9.1f %8.1f s%-3s g%-6s %s\n" % (s, d, a.get("stream", "?"), str(a.get("grid", ["?"])[0]), name))
            per_stream[a.get("stream")] = per_stream.get(a.get("stream"), 0.0) + d
            if e["ts"] > busy_end:
                gaps.append((busy_end - t0, e["ts"] - busy_end))
                busy_end = e["ts"]
            if e["ts"] + d > busy_end:
                busy += e["ts"] + d - busy_end
                busy_end = e["ts"] + d
        span = busy_end - t0
        f.write("# span %.1f us, union of kernel time %.1f us, idle %.1f us in %d gaps\n" % (span, busy, span - busy, len(gaps)))
        f.write("# per-stream kernel time (us): %s\n" % json.dumps(per_stream))
        gaps.sort(key=lambda x: -x[1])
        f.write("# largest gaps (at us, length us): %s\n" % json.dumps([(round(a, 1), round(b, 1)) for a, b in gaps[:15]]))
    txt = open(out).read()
    print(txt[:txt.index(chr(10)) + 1] + txt[-1200:])


if __name__ == "__main__":
    main()
