"""Two launches of the fused gather on 10 GB of tables at B=524288 (run under ncu to read DRAM bytes)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__
__graft_entry__.build()
from fuxictr_b200 import functional as F2
rows, F_, D, B = 4_000_000, 39, 16, 524288
tables = [torch.empty(rows, D, device="cuda").normal_(0, 0.01) for _ in range(F_)]
plan = F2.GatherPlan([F2.GatherField("C%d" % i, i, D, padding_idx=0) for i in range(F_)])
mat = torch.randint(1, rows, (B, F_ + 1), device="cuda").double()
idx = [mat[:, i] for i in range(F_)]
with torch.no_grad():
    for _ in range(2):
        F2.embed_gather(plan, idx, tables)
torch.cuda.synchronize()
print("algorithmic bytes per launch", B * (F_ * 8 + 2 * F_ * D * 4))
