"""Pins oracle/fuxictr_oracle.py against fixtures produced by the REAL reference
(tests/golden/make_golden.py).  CPU only.  Tolerance: the integer gather is bit-exact;
float results must agree to 1e-6 relative (same ATen ops, same order)."""
import sys
from collections import OrderedDict

import numpy as np
import pytest
import torch

from conftest import Golden, rel_err, close, ROOT

sys.path.insert(0, ROOT)
from oracle import fuxictr_oracle as O  # noqa: E402

TOL = 1e-6


def _leafs(state):
    out = OrderedDict()
    for k, v in state.items():
        t = v.clone()
        if t.is_floating_point() and "running_" not in k:
            t.requires_grad_(True)
        out[k] = t
    return out


def _batch(g, matrix=None):
    from fuxictr_b200.schema import FeatureMap
    fm = FeatureMap.from_specs(g.meta["specs"], labels=g.meta["labels"])
    mat = g["in"]["matrix"] if matrix is None else matrix
    return fm, fm.batch_dict(mat)


def _check_grads(g, state, tol=TOL):
    for k, ref in g["g"].items():
        got = state[k].grad
        assert got is not None, k
        assert close(got, ref, tol), k


def test_np_gather_bit_exact():
    g = Golden("feature_embedding_tiny_npz")
    fm, batch = _batch(g)
    specs = g.specs()
    tables = {n: g["w"]["embedding_layer.embedding_layers.%s.weight" % n].numpy() for n in specs}
    mat = g["in"]["matrix"].numpy()
    got = O.np_feature_embedding(specs, tables, mat, fm.column_index, flatten_emb=False)
    assert np.array_equal(got, g["out"]["stack"].numpy())          # bit-exact
    got = O.np_feature_embedding(specs, tables, mat, fm.column_index, flatten_emb=True)
    assert np.array_equal(got, g["out"]["flat"].numpy())


def test_feature_embedding_tiny_npz():
    g = Golden("feature_embedding_tiny_npz")
    fm, batch = _batch(g)
    X = OrderedDict((k, v) for k, v in batch.items() if k not in fm.labels)
    state = _leafs(g["w"])
    out = O.feature_embedding(g.specs(), state, "", X)
    assert torch.equal(out, g["out"]["stack"])
    out.backward(g["in"]["gout"])
    _check_grads(g, state)
    assert torch.equal(O.feature_embedding(g.specs(), state, "", X, flatten_emb=True), g["out"]["flat"])


@pytest.mark.parametrize("tag", ["plain", "avgpool", "sumpool"])
def test_feature_embedding_dict_tiny_seq(tag):
    g = Golden("feature_embedding_dict_tiny_seq_" + tag)
    fm, batch = _batch(g)
    X = OrderedDict((k, v) for k, v in batch.items() if k not in fm.labels)
    state = _leafs(g["w"])
    emb = O.feature_embedding_dict(g.specs(), state, "", X)
    assert list(emb.keys()) == list(g["out"].keys()) or set(emb.keys()) == set(g["out"].keys())
    loss = 0
    for k, ref in g["out"].items():
        assert rel_err(emb[k], ref) <= TOL, k
        loss = loss + (emb[k] * g["gout"][k]).sum()
    loss.backward()
    _check_grads(g, state)


def test_logistic_regression_tiny_seq():
    g = Golden("logistic_regression_tiny_seq")
    fm, batch = _batch(g)
    X = OrderedDict((k, v) for k, v in batch.items() if k not in fm.labels)
    state = _leafs(g["w"])
    out = O.logistic_regression(g.specs(), state, "", X)
    assert rel_err(out, g["out"]["y"]) <= TOL
    out.backward(g["in"]["gout"])
    _check_grads(g, state)


@pytest.mark.parametrize("name", ["inner_product_B16_F7_D10", "inner_product_B9_F39_D16",
                                  "inner_product_B5_F27_D16"])
def test_inner_product(name):
    g = Golden(name)
    for mode, ref in g["out"].items():
        emb = g["in"]["emb"].clone().requires_grad_(True)
        out = O.inner_product_interaction(emb, mode)
        assert rel_err(out, ref) <= TOL, mode
        out.backward(g["in"]["gout_" + mode])
        assert rel_err(emb.grad, g["gin"][mode]) <= TOL, mode


@pytest.mark.parametrize("name", ["crossnet", "crossnet_v2"])
def test_cross(name):
    g = Golden(name)
    state = _leafs(g["w"])
    x0 = g["in"]["x0"].clone().requires_grad_(True)
    fn = O.crossnet if name == "crossnet" else O.crossnet_v2
    out = fn(x0, state, "", g.meta["num_layers"])
    assert rel_err(out, g["out"]["y"]) <= TOL
    out.backward(g["in"]["gout"])
    _check_grads(g, state)
    assert rel_err(x0.grad, g["gin"]["x0"]) <= TOL


def test_cin():
    g = Golden("cin")
    state = _leafs(g["w"])
    emb = g["in"]["emb"].clone().requires_grad_(True)
    out = O.compressed_interaction_net(emb, state, "", g.meta["cin_hidden_units"])
    assert rel_err(out, g["out"]["y"]) <= TOL
    out.backward(g["in"]["gout"])
    _check_grads(g, state)
    assert rel_err(emb.grad, g["gin"]["emb"]) <= TOL


def test_mlp_relu():
    g = Golden("mlp_relu")
    state = _leafs(g["w"])
    x = g["in"]["x"].clone().requires_grad_(True)
    out = O.mlp_block(x, state, "", g.meta["layout"])
    assert rel_err(out, g["out"]["y"]) <= TOL
    out.backward(g["in"]["gout"])
    _check_grads(g, state)
    assert rel_err(x.grad, g["gin"]["x"]) <= TOL


def test_dice():
    g = Golden("dice")
    state = _leafs(g["w"])
    x = g["in"]["x"].clone().requires_grad_(True)
    out = O.dice(x, state, "", training=True)
    assert rel_err(out, g["out"]["train"]) <= TOL
    out.backward(g["in"]["gout"])
    _check_grads(g, state)
    assert rel_err(x.grad, g["gin"]["x"]) <= TOL
    # running statistics were updated in place exactly like nn.BatchNorm1d(momentum=0.01)
    assert rel_err(state["bn.running_mean"], g["w1"]["bn.running_mean"]) <= TOL
    assert rel_err(state["bn.running_var"], g["w1"]["bn.running_var"]) <= TOL
    out_eval = O.dice(x.detach(), state, "", training=False)
    assert rel_err(out_eval, g["out"]["eval"]) <= TOL


@pytest.mark.parametrize("softmax", [0, 1])
def test_din_attention(softmax):
    g = Golden("din_attention_softmax%d" % softmax)
    state = _leafs(g["w"])
    target = g["in"]["target"].clone().requires_grad_(True)
    hist = g["in"]["history"].clone().requires_grad_(True)
    out = O.din_attention(target, hist, g["in"]["mask"], state, "", g.meta["layout"], 8,
                          use_softmax=bool(softmax), training=True)
    assert rel_err(out, g["out"]["y"]) <= TOL
    out.backward(g["in"]["gout"])
    _check_grads(g, state)
    assert rel_err(target.grad, g["gin"]["target"]) <= TOL
    assert rel_err(hist.grad, g["gin"]["history"]) <= TOL


def oracle_pred_fn(name, g):
    kw = g.meta["kwargs"]
    specs = g.specs()
    if name == "DeepFM":
        return lambda s, X: torch.sigmoid(O.deepfm_logit(specs, s, X, len(kw["hidden_units"])))
    if name == "DCNv2":
        return lambda s, X: torch.sigmoid(O.dcnv2_logit(specs, s, X, kw["num_cross_layers"],
                                                         len(kw["parallel_dnn_hidden_units"])))
    if name == "DLRM":
        return lambda s, X: O.dlrm_pred(specs, s, X, len(kw["top_mlp_units"]))
    if name == "xDeepFM":
        return lambda s, X: torch.sigmoid(O.xdeepfm_logit(specs, s, X, kw["cin_hidden_units"],
                                                           len(kw["dnn_hidden_units"])))
    if name == "DIN":
        return lambda s, X: O.din_pred(specs, s, X, kw["embedding_dim"], [("item_id", "cate_id")],
                                       [("click_history", "cate_history")],
                                       len(kw["attention_hidden_units"]), len(kw["dnn_hidden_units"]),
                                       training=True, use_softmax=kw["din_use_softmax"])
    raise KeyError(name)


@pytest.mark.parametrize("name", ["DeepFM", "DCNv2", "DLRM", "xDeepFM", "DIN"])
def test_model_forward_grads_and_train_steps(name):
    g = Golden("model_" + name)
    fm, _ = _batch(g)
    B = g.meta["batch"]
    mat = g["in"]["matrix"]
    batches = [fm.batch_dict(mat[i * B:(i + 1) * B]) for i in range(3)]
    state0 = {k: v for k, v in g["w"].items() if "triu" not in k}
    # forward + loss + every parameter gradient on batch 0
    tr = O.OracleTrainer(state0, oracle_pred_fn(name, g), g.specs(), g.meta["labels"])
    y_pred, y = tr.forward(batches[0])
    assert rel_err(y_pred, g["out"]["y_pred"]) <= TOL
    loss = O.bce_mean(y_pred, y)
    assert rel_err(loss, g["out"]["loss"]) <= TOL
    loss.backward()
    for k, ref in g["g"].items():
        assert rel_err(tr.state[k].grad, ref) <= 2e-6, k
    # three optimisation steps (clip_grad_norm_ + Adam) reproduce the reference trajectory; the
    # same trainer continues (make_golden.py did the same: Dice's running statistics have
    # already seen batch 0 once).
    losses = []
    for i in range(3):
        losses.append(float(tr.train_step(batches[i])))
        if i == 0:
            for k, ref in g["w1"].items():
                if k in tr.state and tr.state[k].is_floating_point():
                    assert rel_err(tr.state[k], ref) <= 2e-6, k
    assert rel_err(torch.tensor(losses), g["out"]["step_losses"]) <= 2e-6
    for k, ref in g["w3"].items():
        if k in tr.state and tr.state[k].is_floating_point():
            assert rel_err(tr.state[k], ref) <= 5e-6, k


def test_metrics_oracle_matches_reference(golden):
    """oracle logloss / AUC (numpy restatement of sklearn's arithmetic) vs the real
    fuxictr.metrics.evaluate_metrics on smooth, tied and saturated predictions."""
    g = golden("metrics_eval")
    y = g["in"]["y_true"].numpy()
    for name in ("smooth", "ties", "saturated"):
        p = g["in"]["y_pred_" + name].numpy()
        r = O.evaluate_metrics(y, p, ["logloss", "AUC"])
        assert abs(r["logloss"] - float(g["out"]["logloss_" + name])) <= 1e-13 * abs(float(g["out"]["logloss_" + name]))
        assert abs(r["AUC"] - float(g["out"]["auc_" + name])) <= 1e-13


def test_metrics_oracle_matches_sklearn_on_random_ties():
    sk = pytest.importorskip("sklearn.metrics")
    import warnings
    rng = np.random.default_rng(5)
    for n, decimals in [(1000, 1), (20000, 3), (5000, 8)]:
        y = (rng.random(n) < 0.3).astype(np.float32)
        p = np.round(rng.random(n), decimals).astype(np.float32)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ll = sk.log_loss(y.astype(np.float64), p.astype(np.float64))
        assert abs(O.logloss(y, p) - ll) <= 1e-13 * ll
        assert abs(O.auc(y, p) - sk.roc_auc_score(y.astype(np.float64), p.astype(np.float64))) <= 1e-13


# ------------------------------------------------------------------ "next" layers (8f-4): oracle pinned before the kernels exist
def _check_next(g, fn, inputs, tol=2e-6):
    """forward output, input gradients and parameter gradients of an oracle function vs the golden."""
    state = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in g["w"].items()}
    ins = {k: (v.clone().requires_grad_(True) if k in g["gin"] else v) for k, v in inputs.items()}
    out = fn(state, ins)
    assert close(out, g["out"]["y"], tol), rel_err(out, g["out"]["y"])
    (out * g["in"]["gout"]).sum().backward()
    for k, ref in g["gin"].items():
        assert close(ins[k].grad, ref, tol), (k, rel_err(ins[k].grad, ref))
    scale = max([float(v.abs().max()) for v in g["g"].values()] + [1e-30])
    for k, ref in g["g"].items():
        assert close(state[k].grad, ref, tol, atol=tol * scale), (k, rel_err(state[k].grad, ref))


@pytest.mark.parametrize("cls_name", ["BilinearInteraction", "BilinearInteractionV2"])
@pytest.mark.parametrize("btype", ["field_all", "field_each", "field_interaction"])
def test_next_bilinear_interaction(golden, cls_name, btype):
    g = golden("next_%s_%s" % (cls_name, btype))
    _check_next(g, lambda st, i: O.bilinear_interaction(st, "", i["emb"], btype), {"emb": g["in"]["emb"]})


@pytest.mark.parametrize("act", ["ReLU", "Sigmoid"])
def test_next_squeeze_excitation(golden, act):
    g = golden("next_SqueezeExcitation_%s" % act)
    _check_next(g, lambda st, i: O.squeeze_excitation(st, "", i["emb"], act), {"emb": g["in"]["emb"]})


@pytest.mark.parametrize("name", ["h1_qkvo1", "h3_qkvo1", "h2_qkvo0"])
def test_next_multi_head_target_attention(golden, name):
    g = golden("next_MHTA_" + name)
    m = g.meta
    _check_next(g, lambda st, i: O.multi_head_target_attention(st, "", i["target"], i["history"], i["mask"],
                                                               m["heads"], m["use_scale"], m["use_qkvo"]),
                {"target": g["in"]["target"], "history": g["in"]["history"], "mask": g["in"]["mask"]})


def test_next_crossnet_mix(golden):
    g = golden("next_CrossNetMix")
    m = g.meta
    _check_next(g, lambda st, i: O.crossnet_mix(st, "", i["x"], m["layer_num"], m["num_experts"]), {"x": g["in"]["x"]})
