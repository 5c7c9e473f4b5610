"""Diagnostic: fused gather bandwidth on tables >> L2 (10 GB) for cudaLimitMaxL2FetchGranularity
settings (one process; the limit is changed between measurements)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from fuxictr_b200 import _lib  # noqa: E402

lib = _lib.load()
torch.zeros(1, device="cuda")
for gran in [None, 32, 64, 128, 32]:
    if gran is not None:
        rc = lib.b2_set_l2_fetch_granularity(gran)
        if rc != 0:
            print(json.dumps({"l2_fetch": gran, "error": lib.b2_last_error().decode()}), flush=True)
            continue
    pts = bench.gather_stress(bench.load_peaks())["points"]
    print(json.dumps({"l2_fetch": gran or "default",
                      "points": [(p["batch"], round(p["ms"], 4), round(p["GBps"]), round(p["frac_of_measured_hbm"], 3))
                                 for p in pts]}), flush=True)
