"""Prints bench.gather_stress() for the current environment (B2_GATHER_STREAM / B2_GATHER_UNROLL / B2_L2_FETCH)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pts = bench.gather_stress(bench.load_peaks())["points"]
print(json.dumps({"env": {k: os.environ.get(k) for k in ("B2_GATHER_STREAM", "B2_GATHER_UNROLL", "B2_L2_FETCH")},
                  "points": [(p["batch"], round(p["ms"], 4), round(p["GBps"]), round(p["frac_of_measured_hbm"], 3)) for p in pts]}))
