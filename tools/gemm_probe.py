"""Timing probe of the tensor-core GEMM on the C2 shapes: which part of a launch costs what.
B2_GEMM_DBG bits (results are WRONG when set; timing only): 1 converters idle, 2 epilogue skips its global
stores, 4 no MMA instructions, 8 two-stage ring, 16 no TMA loads, 32 helper warps do not join the epilogue.   usage: python tools/gemm_probe.py"""
import json
import os
import sys

os.environ.setdefault("B2_BUILD_PROBE", "1")     # the probes exist only in a -DB2_GEMM_PROBE build

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fuxictr_b200 import build as _build  # noqa: E402

_build.build()       # rebuilds in probe mode when the in-tree library is the product build (and vice versa later)
import bench  # noqa: E402
from fuxictr_b200 import functional as F2  # noqa: E402

torch.manual_seed(0)
B = 4096
SHAPES = {"fwd1 4096x300x624": (B, 300, 624, False, False), "fwd2 4096x300x300": (B, 300, 300, False, False),
          "dgrad1 4096x624x300": (B, 624, 300, False, True), "dgrad2 4096x300x300": (B, 300, 300, False, True),
          "wgrad1 300x624x4096": (300, 624, B, True, True), "tiny 4096x300x32": (B, 300, 32, False, False)}


def run(label, inline, env):
    for k in ("B2_GEMM_DBG", "B2_X3_BN_MAX"):
        os.environ.pop(k, None)
    os.environ.update(env)
    F2.set_x3_inline(inline)
    F2.set_matmul_precision("tf32x3")
    row = {}
    for name, (M, N, K, a_mn, b_mn) in SHAPES.items():
        a = torch.randn((K, M) if a_mn else (M, K), device="cuda")
        b = torch.randn((K, N) if b_mn else (N, K), device="cuda")
        out = torch.zeros(M, N, device="cuda")
        asm, bsm = F2.make_aux(a), F2.make_aux(b)
        row[name] = round(1e3 * bench.time_kernel(
            lambda: F2.gemm_ex(a, b, out, a_mn=a_mn, b_mn=b_mn, a_small=asm, b_small=bsm), 40), 2)
    F2.set_matmul_precision("fp32")
    print(json.dumps({"cfg": label, "us": row}), flush=True)


run("aux  (small parts from HBM)", False, {})
run("aux  4 epilogue warps", False, {"B2_GEMM_DBG": "32"})
run("aux  no epilogue stores", False, {"B2_GEMM_DBG": "2"})
run("aux  no TMA", False, {"B2_GEMM_DBG": "16"})
run("aux  no TMA, no stores", False, {"B2_GEMM_DBG": "18"})
run("aux  no MMA, no stores", False, {"B2_GEMM_DBG": "6"})
run("aux  no TMA, no MMA, no stores", False, {"B2_GEMM_DBG": "22"})
run("aux  bn<=128", False, {"B2_X3_BN_MAX": "128"})
run("inline", True, {})
run("inline 4 epilogue warps", True, {"B2_GEMM_DBG": "32"})
run("inline converters idle", True, {"B2_GEMM_DBG": "1"})
run("inline no TMA", True, {"B2_GEMM_DBG": "16"})
run("inline no TMA, converters idle", True, {"B2_GEMM_DBG": "17"})
