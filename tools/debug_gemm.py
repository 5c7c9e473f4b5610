"""Diagnostic (not a test): in-kernel timeline of the tcgen05 GEMM."""
import sys, os, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__; __graft_entry__.build()
from fuxictr_b200 import functional as F2, _lib
dbg = torch.zeros(8, dtype=torch.int64, device="cuda")
names = ["entry", "setup", "tma0", "full0", "mma_done", "epi_start", "epi_done", "exit"]
for (M, N, K, mode) in [(4096, 300, 624, "tf32"), (4096, 300, 300, "tf32"), (300, 300, 4096, "tf32"), (4096, 300, 624, "tf32x3")]:
    a = torch.randn(M, K, device="cuda"); b = torch.randn(N, K, device="cuda"); out = torch.empty(M, N, device="cuda")
    F2.set_matmul_precision(mode)
    for it in range(3):
        _lib.call("b2_gemm_tc_set_debug", ctypes.c_void_p(dbg.data_ptr()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); F2.gemm_nt(a, b, out); e1.record(); torch.cuda.synchronize()
        t = dbg.cpu().tolist()
        print(M, N, K, mode, "event %.1f us |" % (e0.elapsed_time(e1) * 1e3), " ".join("%s+%.1f" % (n, (x - t[0]) / 1e3) for n, x in zip(names, t)))
    _lib.call("b2_gemm_tc_set_debug", ctypes.c_void_p(0))
    F2.set_matmul_precision("fp32")
