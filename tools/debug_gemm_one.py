"""Diagnostic: a few launches of one tcgen05 GEMM shape, for `ncu --set full`."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__; __graft_entry__.build()
from fuxictr_b200 import functional as F2
M, N, K = 4096, 300, 624
a = torch.randn(M, K, device="cuda"); b = torch.randn(N, K, device="cuda"); out = torch.empty(M, N, device="cuda")
bias = torch.randn(N, device="cuda")
F2.set_matmul_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32")
for _ in range(4):
    F2.gemm_nt(a, b, out, bias=bias, act=1)
torch.cuda.synchronize()
