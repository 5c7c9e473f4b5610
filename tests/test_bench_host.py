"""Host-side pieces of bench.py (no GPU): workload definitions, id generators, clock-sample windows."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_c5_workload_is_the_200m_row_layout():
    assert len(bench.DLRM_VOCABS) == 26 and sum(bench.DLRM_VOCABS) == 200_000_000       # SURVEY 8d cardinalities
    assert bench.DLRM_GLOBAL_BATCH == 65536 and bench.DLRM_TOP == [64, 64, 64]
    assert bench.NF == 39 and bench.VOCAB == 25641 and bench.DIM == 16 and bench.HIDDEN == [300, 300, 300]   # C2
    assert bench.DEFAULT_BATCH == {"deepfm": 4096, "dcnv2": 8192, "din": 2048, "xdeepfm": 4096}


def test_zipf_ids_stay_inside_the_vocabulary_and_are_skewed():
    vs = [25641, 1000, 17]
    ids = bench.zipf_ids(20000, vs, seed=3)
    assert ids.shape == (20000, 3) and ids.dtype == torch.float64
    for j, v in enumerate(vs):
        col = ids[:, j]
        assert float(col.min()) >= 1 and float(col.max()) <= v - 1 and bool((col == col.floor()).all())
    head = float((ids[:, 0] <= 10).double().mean())        # Zipf(1.05): the ten hottest rows take a large share
    assert 0.15 < head < 0.6
    assert torch.equal(ids, bench.zipf_ids(20000, vs, seed=3))           # seed-fixed


def test_clock_sampler_keeps_only_rows_inside_the_load_windows():
    s = bench.ClockSampler(0)
    s.proc = type("P", (), {"terminate": lambda self: None, "wait": lambda self, timeout=None: 0})()
    row = lambda mhz, reasons: ["0", str(mhz), "1965", "400", "0x0"] + reasons     # noqa: E731
    idle, hot = ["Not Active"] * 4, ["Not Active", "Not Active", "Not Active", "Active"]
    s.rows = [(10.0, row(345, idle)), (11.0, row(1965, idle)), (11.5, row(1950, hot)), (12.0, row(1965, idle)),
              (20.0, row(600, idle))]
    assert s.count_between(10.5, 12.5) == 3
    out = s.stop([(10.5, 12.5)])
    assert out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0 and out["samples"] == 3
    assert out["reasons"] == ["sw_power_cap"]                            # kept and reported, not a rejection reason
    s.proc = type("P", (), {"terminate": lambda self: None, "wait": lambda self, timeout=None: 0})()
    assert s.stop(None)["samples"] == 5
