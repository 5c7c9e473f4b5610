"""Locate a wrong / nondeterministic contraction: every GEMM shape of an MLP chain, each operand layout,
single-pass TF32 vs fp64, repeated — prints the max-norm error per (shape, layout, repetition)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__
__graft_entry__.build()
from fuxictr_b200 import functional as F2

mode = sys.argv[1] if len(sys.argv) > 1 else "tf32"
F2.set_matmul_precision(mode)
gen = torch.Generator().manual_seed(0)
B = 2048
shapes = [("fwd", B, 500, 432, False, False), ("fwd", B, 500, 500, False, False),
          ("dgrad", B, 432, 500, False, True), ("dgrad", B, 500, 500, False, True),
          ("wgrad", 500, 432, B, True, True), ("wgrad", 500, 500, B, True, True)]
for name, M, N, K, a_mn, b_mn in shapes:
    a = torch.randn(M, K, generator=gen)
    b = torch.randn(N, K, generator=gen)
    ref = a.double() @ b.double().t()
    ad = (a.t().contiguous() if a_mn else a).cuda()
    bd = (b.t().contiguous() if b_mn else b).cuda()
    errs = []
    outs = []
    for rep in range(6):
        out = torch.full((M, N), float("nan"), device="cuda")
        F2.gemm_ex(ad, bd, out, a_mn=a_mn, b_mn=b_mn, a_small=F2.make_aux(ad), b_small=F2.make_aux(bd))
        torch.cuda.synchronize()
        errs.append(float((out.double().cpu() - ref).abs().max() / ref.abs().max()))
        outs.append(out)
    same = all(torch.equal(outs[0], o) for o in outs[1:])
    bad = (outs[0].double().cpu() - ref).abs() > 0.02 * ref.abs().max()
    rows = bad.any(dim=1).nonzero().flatten()[:8].tolist()
    cols = bad.any(dim=0).nonzero().flatten()[:8].tolist()
    print("%s %s M=%d N=%d K=%d a_mn=%d b_mn=%d  err %s deterministic=%s bad=%d rows%s cols%s"
          % (mode, name, M, N, K, a_mn, b_mn, ["%.1e" % e for e in errs], same, int(bad.sum()), rows, cols), flush=True)
