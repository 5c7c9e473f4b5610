"""SURVEY.md 8f rows 1 and 3 on the GPU: the device-resident evaluation metrics (csrc/metrics.cu)
against the reference's goldens and the numpy oracle, and the host->HBM training pipeline against
the plain eager step.  Bars: the radix sort and the Mann-Whitney numerator are integer work ->
bit-exact; logloss / AUC are fp64 reductions -> 1e-12 relative."""
import ctypes
import sys

import numpy as np
import pytest
import torch

from conftest import Golden, ROOT

sys.path.insert(0, ROOT)
from oracle import fuxictr_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__
    __graft_entry__.build()
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"


def _sort(keys_np):
    from fuxictr_b200 import _lib, functional as F2
    n = keys_np.size
    keys = torch.from_numpy(keys_np.view(np.int32).copy()).cuda()
    nbytes = ctypes.c_int64(0)
    _lib.call("b2_auc_workspace_bytes", max(n, 1), ctypes.byref(nbytes))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
    _lib.call("b2_sort_u32", F2._ptr(keys), n, F2._ptr(ws), nbytes.value, F2._stream())
    torch.cuda.synchronize()
    return keys.cpu().numpy().view(np.uint32)


@pytest.mark.parametrize("n", [1, 2, 255, 256, 257, 4097, 151_553, 1_000_003, 6_000_000])
def test_radix_sort_bit_exact(n):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    assert np.array_equal(_sort(keys), np.sort(keys))


@pytest.mark.parametrize("mask", [0x000000FF, 0xFF000000, 0x00010001, 0])
def test_radix_sort_skewed_digits(mask):
    """Few distinct values per digit (what sigmoid scores look like in the top byte) and the
    all-equal case: every key of a chunk lands in one bucket."""
    rng = np.random.default_rng(mask & 0xFFFF)
    keys = (rng.integers(0, 2 ** 32, 300_000, dtype=np.uint64).astype(np.uint32)) & np.uint32(mask)
    assert np.array_equal(_sort(keys), np.sort(keys))


def test_metrics_match_reference_golden():
    from fuxictr_b200 import metrics
    g = Golden("metrics_eval")
    y = g["in"]["y_true"].cuda()
    for name in ("smooth", "ties", "saturated"):
        r = metrics.evaluate_metrics(y, g["in"]["y_pred_" + name].cuda(), ["logloss", "AUC"])
        ll, auc = float(g["out"]["logloss_" + name]), float(g["out"]["auc_" + name])
        assert list(r.keys()) == ["logloss", "AUC"]
        assert abs(r["logloss"] - ll) <= 1e-12 * abs(ll), (name, r["logloss"], ll)
        assert abs(r["AUC"] - auc) <= 1e-12, (name, r["AUC"], auc)


@pytest.mark.parametrize("n,decimals,pos_rate", [(1, None, 1.0), (37, 1, 0.5), (200_000, 3, 0.25),
                                                 (5_000_000, None, 0.03), (300_000, 0, 0.5)])
def test_auc_numerator_is_exact(n, decimals, pos_rate):
    """{n_neg, n_pos, 2U} equal the oracle's integers (ties, negative scores, signed zeros, one class)."""
    from fuxictr_b200 import _lib, functional as F2
    rng = np.random.default_rng(n)
    y = (rng.random(n) < pos_rate).astype(np.float32)
    p = rng.normal(size=n).astype(np.float32)          # raw scores: negative values exercise the key map
    if decimals is not None:
        p = np.round(p, decimals).astype(np.float32)   # includes -0.0 and +0.0
    yp, yt = torch.from_numpy(p).cuda(), torch.from_numpy(y).cuda()
    nbytes = ctypes.c_int64(0)
    _lib.call("b2_auc_workspace_bytes", n, ctypes.byref(nbytes))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
    res = torch.full((5,), -1, dtype=torch.int64, device="cuda")
    _lib.call("b2_auc", F2._ptr(yp), F2._ptr(yt), n, F2._ptr(ws), nbytes.value, F2._ptr(res), F2._stream())
    n_neg, n_pos, n_nan, n_bad, twice_u = (int(v) for v in res.cpu())
    want_u, want_pos, want_neg = O.auc_twice_u(y, p)
    assert (n_neg, n_pos, n_nan, n_bad) == (want_neg, want_pos, 0, 0)
    assert twice_u == want_u


def test_logloss_and_auc_large_split():
    from fuxictr_b200 import metrics
    rng = np.random.default_rng(77)
    n = 4_600_000                                       # ~ Criteo_x1 validation split
    y = (rng.random(n) < 0.256).astype(np.float32)
    p = (1.0 / (1.0 + np.exp(-(rng.normal(size=n) * 1.3 + 1.1 * y - 1.4)))).astype(np.float32)
    r = metrics.evaluate_metrics(torch.from_numpy(y).cuda(), torch.from_numpy(p).cuda(), ["AUC", "logloss"])
    want = O.evaluate_metrics(y, p, ["AUC", "logloss"])
    assert list(r.keys()) == ["AUC", "logloss"]
    assert abs(r["AUC"] - want["AUC"]) <= 1e-12
    assert abs(r["logloss"] - want["logloss"]) <= 1e-12 * want["logloss"]


def test_metric_errors_mirror_the_reference():
    from fuxictr_b200 import metrics
    y = torch.tensor([0., 1., 1., 0.]).cuda()
    p = torch.tensor([0.2, 0.7, 0.4, 0.1]).cuda()
    assert metrics.evaluate_metrics(y, p, ["AUC"])["AUC"] == 1.0
    with pytest.raises(ValueError, match="Only one class"):
        metrics.evaluate_metrics(torch.ones(4).cuda(), p, ["AUC"])
    with pytest.raises(ValueError, match="NaN"):
        metrics.evaluate_metrics(y, torch.tensor([0.2, float("nan"), 0.4, 0.1]).cuda(), ["AUC"])
    with pytest.raises(ValueError, match="binary"):
        metrics.evaluate_metrics(torch.tensor([0., 2., 1., 0.]).cuda(), p, ["AUC"])
    with pytest.raises(ValueError, match="not supported"):
        metrics.evaluate_metrics(y, p, ["F1"])
    with pytest.raises(NotImplementedError):
        metrics.evaluate_metrics(y, p, ["gAUC"])
    with pytest.raises(RuntimeError, match="CUDA"):
        metrics.evaluate_metrics(y.cpu(), p.cpu(), ["AUC"])


def _small_deepfm(seed=11):
    from fuxictr_b200 import zoo
    from fuxictr_b200.schema import FeatureMap
    specs = [("C%d" % i, {"type": "categorical", "source": "", "padding_idx": 0, "vocab_size": 300 + 13 * i})
             for i in range(8)]
    fm = FeatureMap.from_specs(specs, embedding_dim=8)
    torch.manual_seed(seed)
    model = zoo.DeepFM(fm, gpu=0, embedding_dim=8, hidden_units=[32, 16])
    with torch.no_grad():
        for mod in model.modules():
            if isinstance(mod, torch.nn.Embedding):
                mod.weight[1:].normal_(0, 0.3)
    return fm, specs, model


def _host_batches(specs, nbatches, B, seed):
    gen = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(nbatches):
        ids = torch.cat([torch.randint(0, s["vocab_size"], (B, 1), generator=gen) for _, s in specs], 1)
        out.append(torch.cat([ids.double(), (torch.rand(B, 1, generator=gen) < 0.4).double()], 1))
    return out


def test_model_evaluate_and_predict_keep_predictions_on_device():
    """RankModel.evaluate / predict (rank_model.py:350-398) over a generator whose last batch is short
    and whose total exceeds the accumulator's initial capacity."""
    from fuxictr_b200 import metrics as M
    fm, specs, model = _small_deepfm()
    mats = _host_batches(specs, 5, 700, seed=3) + _host_batches(specs, 1, 123, seed=4)
    gen_batches = [fm.batch_dict(m.cuda()) for m in mats]
    old = M.DeviceMetrics.__init__.__defaults__
    M.DeviceMetrics.__init__.__defaults__ = (1024,)     # force two growth steps
    try:
        logs = model.evaluate(iter(gen_batches), metrics=["logloss", "AUC"])
        preds = model.predict(iter(gen_batches))
    finally:
        M.DeviceMetrics.__init__.__defaults__ = old
    model.eval()
    with torch.no_grad():
        want_p = np.concatenate([model.forward(b)["y_pred"].cpu().numpy().reshape(-1) for b in gen_batches])
    want_y = np.concatenate([m[:, -1].numpy() for m in mats])
    assert preds.dtype == np.float64 and preds.shape == (5 * 700 + 123,)
    assert np.allclose(preds, want_p, rtol=1e-6, atol=0)
    want = O.evaluate_metrics(want_y, want_p, ["logloss", "AUC"])      # exactness is pinned by the tests above;
    assert abs(logs["logloss"] - want["logloss"]) <= 1e-6 * want["logloss"]   # this one checks the plumbing
    assert abs(logs["AUC"] - want["AUC"]) <= 1e-6


@pytest.mark.parametrize("graph", [True, False])
def test_train_pipeline_matches_eager_steps(graph):
    """TrainPipeline.step (pinned or pageable host matrix -> copy-stream H2D -> captured step) walks
    the same trajectory as fused_train_step on device-resident batches."""
    from fuxictr_b200.pipeline import TrainPipeline
    B, warm = 256, 3
    fm, specs, ref = _small_deepfm()
    _, _, model = _small_deepfm()
    ref.use_fused_optimizer()
    model.use_fused_optimizer()
    mats = _host_batches(specs, 7, B, seed=21)
    prime = mats[0]
    pipe = TrainPipeline(model, B, len(specs) + 1, torch.float64, graph=False)
    pipe.prime(prime.cuda())
    pipe.step_device(prime.cuda())                       # an eager step BEFORE the capture (bench.py counts launches so)
    if graph:
        pipe._capture(warm)                              # 3 real warm-up steps on the primed batch
    ref.train()
    ref.fused_train_step(fm.batch_dict(prime.cuda()))
    if graph:
        for _ in range(warm):
            ref.fused_train_step(fm.batch_dict(prime.cuda()))
    want, got = [], []
    for k, m in enumerate(mats[1:]):
        want.append(float(ref.fused_train_step(fm.batch_dict(m.cuda()))))
        host = m.pin_memory() if k % 2 == 0 else m       # both input paths
        pipe.step(host)
        got.append(pipe.loss())
    pipe.wait_inputs()
    assert np.allclose(got, want, rtol=1e-5, atol=0), (got, want)
    dn, pn = dict(ref.named_parameters()), dict(model.named_parameters())
    for k in dn:
        err = float((dn[k] - pn[k]).abs().max())
        assert err <= 1e-5 * max(float(dn[k].abs().max()), 1e-3), (k, err)
    assert pipe.h2d_bytes_per_step == B * (len(specs) + 1) * 8 and pipe.d2h_bytes_per_step == 4
    with pytest.raises(ValueError):
        pipe.step(torch.zeros(B, len(specs), dtype=torch.float64))


def test_loader_feeds_pipeline_from_pinned_ring(tmp_path):
    """dataloader.NpzDataLoader(shuffle=True).matrices() -> TrainPipeline.step: the pinned ring slots
    are consumed asynchronously (copy stream) while the prefetch thread refills later slots; the
    trajectory equals eager steps over the same (cloned) batches, and model.evaluate() runs over the
    loader's dict protocol."""
    from fuxictr_b200 import dataloader as DL
    from fuxictr_b200.pipeline import TrainPipeline
    fm, specs, ref = _small_deepfm()
    _, _, model = _small_deepfm()
    rng = np.random.default_rng(0)
    n, B = 4096 + 37, 128
    cols = {name: rng.integers(0, s["vocab_size"], n).astype(np.int64) for name, s in specs}
    cols["label"] = (rng.random(n) < 0.4).astype(np.float64)
    path = str(tmp_path / "train.npz")
    np.savez(path, **cols)
    loader = DL.NpzDataLoader(fm, path, batch_size=B, shuffle=True)
    assert not loader.matrix.is_pinned() and len(loader) == 33     # shuffled: batches leave through the pinned ring slots
    torch.manual_seed(5)
    kept = [m.clone() for m in loader.matrices()]          # the epoch's batches, in order
    ref.use_fused_optimizer()
    model.use_fused_optimizer()
    ref.train()
    want = [float(ref.fused_train_step(fm.batch_dict(m.cuda()))) for m in kept[:-1]]
    pipe = TrainPipeline(model, B, len(specs) + 1, torch.float64, graph=False)
    torch.manual_seed(5)
    got = []
    for m in loader.matrices():
        if m.shape[0] != B:                                # the short last batch is not the captured shape
            break
        assert m.is_pinned()
        pipe.step(m)                                       # no host sync here: the ring must stay ahead safely
        got.append(pipe.loss_host.clone())                 # (value read after the epoch)
    last = pipe.loss()
    assert len(got) == len(want) == 32
    assert abs(last - want[-1]) <= 1e-5 * abs(want[-1])
    dn, pn = dict(ref.named_parameters()), dict(model.named_parameters())
    for k in dn:
        err = float((dn[k] - pn[k]).abs().max())
        assert err <= 1e-5 * max(float(dn[k].abs().max()), 1e-3), (k, err)
    logs = model.evaluate(loader, metrics=["logloss", "AUC"])
    assert 0.0 < logs["AUC"] < 1.0 and logs["logloss"] > 0.0
