"""Cycle-stamped timeline of CTA 0 of one tensor-core GEMM launch (B2_GEMM_TRACE probe).
usage: python tools/gemm_trace.py [inline|aux] [M N K]"""
import os
import sys

os.environ.setdefault("B2_BUILD_PROBE", "1")     # the probes exist only in a -DB2_GEMM_PROBE build

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fuxictr_b200 import build as _build  # noqa: E402

_build.build()       # rebuilds in probe mode when the in-tree library is the product build (and vice versa later)
from fuxictr_b200 import functional as F2  # noqa: E402

inline = (sys.argv[1] if len(sys.argv) > 1 else "aux") == "inline"
kind = sys.argv[2] if len(sys.argv) > 2 else "fwd"          # fwd: bias + relu epilogue; dgrad: B MN-major + relu backward; wgrad: split-K
M, N, K = {"fwd": (4096, 300, 624), "dgrad": (4096, 624, 300), "wgrad": (300, 624, 4096)}[kind]
torch.manual_seed(0)
a_mn, b_mn = kind == "wgrad", kind != "fwd"
a = torch.randn((K, M) if a_mn else (M, K), device="cuda")
b = torch.randn((K, N) if b_mn else (N, K), device="cuda")
out = torch.zeros(M, N, device="cuda")
extra = {}
if kind == "fwd":
    extra = dict(bias=torch.randn(N, device="cuda"), act=1)
elif kind == "dgrad":
    extra = dict(ybwd=torch.randn(M, N, device="cuda"), act_bwd=1, colsum=torch.zeros(N, device="cuda"))
F2.set_x3_inline(inline)
F2.set_matmul_precision("tf32x3")
asm, bsm = F2.make_aux(a), F2.make_aux(b)
buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
for _ in range(3):
    F2.gemm_ex(a, b, out, a_mn=a_mn, b_mn=b_mn, a_small=asm, b_small=bsm, **extra)
torch.cuda.synchronize()
os.environ["B2_GEMM_TRACE"] = str(buf.data_ptr())
F2.gemm_ex(a, b, out, a_mn=a_mn, b_mn=b_mn, a_small=asm, b_small=bsm, **extra)
torch.cuda.synchronize()
os.environ.pop("B2_GEMM_TRACE")
t = buf.cpu().tolist()
t0 = t[640]
nkb = (K + 31) // 32
print("mode", "inline" if inline else "aux", kind, "shape", M, N, K, "k-blocks", nkb, "(cycles since the prologue barrier)")
print("producer: kb  loop_top  empty_ok  issued")
for i in range(min(nkb, 60)):
    if t[4 * i]:
        print("   %2d %8d %8d %8d" % (i, t[4 * i] - t0, t[4 * i + 1] - t0, t[4 * i + 2] - t0))
print("mma:      kb  top  half1_issued  half2+commit  next_ready")
for i in range(min(nkb, 60)):
    r = t[256 + 4 * i:256 + 4 * i + 4]
    if r[0]:
        print("   %2d %8d %8d %8d %8d" % (i, r[0] - t0, r[1] - t0, r[2] - t0, r[3] - t0))
print("epilogue warps: wait_start  tmem_full  done | first chunk: tmem_read  transposed+bias  stored")
for w in range(2, 10):
    r = t[512 + 4 * w:512 + 4 * w + 3]
    if r[0]:
        e = t[700 + 8 * w:700 + 8 * w + 3]
        print("   w%d %8d %8d %8d | %8d %8d %8d" % (w, r[0] - t0, r[1] - t0, r[2] - t0, e[0] - t0, e[1] - t0, e[2] - t0))
