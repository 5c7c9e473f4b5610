"""Times the device-resident evaluation (csrc/metrics.cu) on a Criteo_x1-size validation split and,
beside it, the arithmetic the reference runs on the host for the same arrays
(sklearn log_loss + roc_auc_score, fuxictr/metrics.py:45-48).  Prints one JSON line.

    python tools/bench_eval.py [--n 4600000] [--reps 10]
"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4_600_000)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    import __graft_entry__
    __graft_entry__.build()
    from fuxictr_b200 import metrics
    rng = np.random.default_rng(1)
    n = args.n
    y = (rng.random(n) < 0.256).astype(np.float32)
    p = (1.0 / (1.0 + np.exp(-(rng.normal(size=n) * 1.3 + 1.1 * y - 1.4)))).astype(np.float32)
    yd, pd = torch.from_numpy(y).cuda(), torch.from_numpy(p).cuda()
    for _ in range(3):
        got = metrics.evaluate_metrics(yd, pd, ["logloss", "AUC"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        metrics.evaluate_metrics(yd, pd, ["logloss", "AUC"])      # includes its 48-byte D2H + host sync
    e1.record()
    torch.cuda.synchronize()
    gpu_ms = e0.elapsed_time(e1) / args.reps
    from sklearn.metrics import log_loss, roc_auc_score
    y64, p64 = y.astype(np.float64), p.astype(np.float64)
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = {"logloss": log_loss(y64, p64), "AUC": roc_auc_score(y64, p64)}
    cpu_ms = (time.perf_counter() - t0) * 1e3
    # algorithmic bytes: read p, y (8n) + write keys (4n) + 4 radix passes x (read 2x4n_neg + write 4n_neg)
    n_neg = int((y == 0).sum())
    algo = 8 * n + 4 * n + 4 * 12 * n_neg + 8 * n
    print(json.dumps({"n": n, "gpu_ms": gpu_ms, "cpu_sklearn_ms": cpu_ms, "speedup": cpu_ms / gpu_ms,
                      "algorithmic_bytes": algo, "GBps": algo / gpu_ms / 1e6,
                      "logloss": got["logloss"], "AUC": got["AUC"],
                      "abs_err_logloss": abs(got["logloss"] - want["logloss"]),
                      "abs_err_auc": abs(got["AUC"] - want["AUC"])}))


if __name__ == "__main__":
    main()
