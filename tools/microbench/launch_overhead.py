import os, sys, torch
sys.path.insert(0, "/root/repo" if os.path.exists("/root/repo/ophelia_b200") else os.getcwd())
from ophelia_b200 import ops, _lib
dev = torch.device("cuda:0")
A = torch.randn(256, 256, device=dev); Bm = torch.randn(256, 256, device=dev)
y = torch.randn(27840, 256, device=dev)
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
print("tiny gemm back-to-back: %.1f us" % timeit(lambda: ops.gemm_nt(A, Bm)))
print("torch elementwise only: %.1f us" % timeit(lambda: y.mul_(1.0001)))
def alt():
    ops.gemm_nt(A, Bm); y.mul_(1.0001)
print("alternating gemm + elementwise: %.1f us" % timeit(alt))
C = torch.zeros(1, 256, 256, device=dev)
def raw():
    _lib.call("oph_gemm_nt", A.data_ptr(), 256, Bm.data_ptr(), 256, C.data_ptr(), 256, None, 256, 256, 256, 1, 1.0, 1, 0, 0, 0, torch.cuda.current_stream().cuda_stream)
print("raw tiny gemm (no allocs): %.1f us" % timeit(raw))

dbg = torch.zeros(74, 24, dtype=torch.int64, device=dev)
_lib.call("oph_gemm_debug_buffer", dbg.data_ptr())
raw(); torch.cuda.synchronize()
print("with timers: %.1f us" % timeit(raw))
d = dbg.cpu()[0].tolist()
print("pair 0: mma-loop cycles %d (wait acc %d, A %d, B %d, kblocks %d); ns from entry: mma loop end %d, producer done %d, after final cluster sync %d" % tuple(d))
