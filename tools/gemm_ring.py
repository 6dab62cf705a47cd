"""In-kernel time stamps of EVERY GEMM launch of one replay of the captured training step (oph_gemm_debug_ring: each
launch of the capture owns a slot [74][16] of globaltimer stamps).  Prints per launch: start (us after the first GEMM of the
step), span, pairs, mean ns to: first copy issued / first A landed / first operands landed / first accumulator complete /
MMA loop end / last epilogue end / exit, and the SM-time accounting (busy = sum over pairs of entry->exit).

    python tools/gemm_ring.py [--batch 32] [--eager 1]
"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--N", type=int, default=180)
    ap.add_argument("--T", type=int, default=870)
    ap.add_argument("--eager", type=int, default=0, help="1: stamps of an eager single-stream step instead of a graph replay")
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--out", default="gemm_ring.txt")
    args = ap.parse_args()
    import torch
    import __graft_entry__
    __graft_entry__.build()
    from ophelia_b200 import _lib
    from ophelia_b200.architectures import Text2MelGraph
    from ophelia_b200.configuration import default_hparams
    from ophelia_b200.data import SyntheticBatches
    from ophelia_b200.variables import VariableStore
    lib = _lib.load()
    if args.flags:
        _lib.set_debug_flags(args.flags)
    dev = torch.device("cuda", 0)
    hp = default_hparams(max_N=args.N, max_T=args.T, seed=0)
    src = SyntheticBatches(hp, "t2m", args.batch, N=args.N, T=args.T, seed=1234)
    store = VariableStore(dev, seed=0)
    g = Text2MelGraph(hp, mode="train", store=store, data=src, device=dev)
    b0 = src.batches[0]
    dev_in = (b0["text"].to(dev), b0["mel"].to(dev))
    for _ in range(3):
        g.train_step_device(*dev_in)
    torch.cuda.synchronize()
    SLOTS = 512
    ring = torch.zeros(SLOTS, 74, 24, dtype=torch.int64, device=dev)
    if args.eager:
        hp.use_side_streams = False
        g.train_step_device(*dev_in)
        torch.cuda.synchronize()
        lib.oph_gemm_debug_ring(ctypes.c_void_p(ring.data_ptr()), SLOTS)
        g.train_step_device(*dev_in)
        torch.cuda.synchronize()
    else:
        # the warm-up steps inside capture_train_step take slots too; after zeroing the ring only the captured launches
        # (replayed below) write stamps again
        lib.oph_gemm_debug_ring(ctypes.c_void_p(ring.data_ptr()), SLOTS)
        step = g.capture_train_step(*dev_in)
        for _ in range(3):
            step(*dev_in)
        torch.cuda.synchronize()
        ring.zero_()
        torch.cuda.synchronize()
        step(*dev_in)
        torch.cuda.synchronize()
    d = ring.cpu().double()
    buf = ctypes.create_string_buffer(256)
    rows = []
    for s in range(SLOTS):
        x = d[s]
        x = x[x[:, 8] > 0]
        if len(x) == 0:
            continue
        desc = ""
        if lib.oph_gemm_debug_ring_desc(s, buf, 256) == 0:
            desc = buf.value.decode()
        rows.append((float(x[:, 8].min()), s, x, desc))
    rows.sort(key=lambda r: r[0])
    T0 = rows[0][0]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = os.path.join(ROOT, "gpurun_out", args.out)
    tot_busy = tot_mma = tot_start = tot_tail = 0.0
    with open(out, "w") as f:
        f.write("# start_us span_us pairs | ns mean(max): bar_init tmem_alloc prologue_end pre_setmaxnreg loader_in decoded | copy1 A1 ops1 acc1 mma_end epi_end exit | skew_max | desc\n")
        for t0, s, x, desc in rows:
            span = x[:, 13].max() - t0
            skew = x[:, 8] - t0
            m = lambda c: "%5.1f(%5.1f)" % (x[:, c].mean() / 1e3, x[:, c].max() / 1e3)  # noqa: E731
            f.write("%8.1f %6.1f %3d | %s %s %s %s %s %s | %s %s %s %s %s %s %s | %5.1f | %s\n" % ((t0 - T0) / 1e3, span / 1e3, len(x), m(17), m(16), m(9), m(18), m(19), m(20), m(14), m(15), m(10), m(11),
                                                                           m(5), m(12), m(7), skew.max() / 1e3, desc))
            busy = float((x[:, 13] - x[:, 8]).sum())
            tot_busy += busy
            tot_start += float(x[:, 10].sum())
            tot_mma += float((x[:, 5] - x[:, 10]).sum())
            tot_tail += float((x[:, 7] - x[:, 5]).sum())
        end = max(float(r[2][:, 13].max()) for r in rows)
        f.write("# %d GEMM launches; first entry -> last exit %.1f us; pair-time (sum over pairs, /74 = us of the whole chip): busy %.1f, "
                "of which before the first operands %.1f, MMA loop %.1f, after the MMA loop %.1f\n" %
                (len(rows), (end - T0) / 1e3, tot_busy / 74e3, tot_start / 74e3, tot_mma / 74e3, tot_tail / 74e3))
    print(open(out).read()[-6000:])


if __name__ == "__main__":
    main()
