"""Diagnostic: fused gather bandwidth on tables >> L2 for a few launch variants."""
import os, sys, subprocess, json
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    print(json.dumps(bench.gather_stress(bench.load_peaks())["points"]))
else:
    for unroll, stream in [(4, 0), (4, 1), (8, 0), (8, 1)]:
        env = dict(os.environ, B2_GATHER_UNROLL=str(unroll), B2_GATHER_STREAM=str(stream))
        out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        try:
            pts = json.loads(out.stdout.strip().splitlines()[-1])
            print("unroll", unroll, "stream", stream, [(p["batch"], round(p["GBps"]), round(p["frac_of_measured_hbm"], 3)) for p in pts], flush=True)
        except Exception:
            print("unroll", unroll, "stream", stream, "FAILED", out.stderr[-400:], flush=True)
