"""Row-sharded front on ONE GPU with `world` virtual ranks (fuxictr_b200.sharded.VirtualPeerGroup):
the push/pull kernels cannot tell a local pointer from an NVLink peer pointer, so running every
phase for all virtual ranks in lock step exercises exactly the code a torchrun job executes.
Checked against the unsharded fused front (itself pinned to the reference goldens)."""
import sys
from collections import OrderedDict

import pytest
import torch

from conftest import close, ROOT

sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu
RTOL = 1e-5


@pytest.mark.parametrize("world", [1, 2, 4])
@pytest.mark.parametrize("dim,nf,want_fm", [(16, 39, True), (8, 7, False), (40, 5, True)])
def test_virtual_ranks_match_unsharded(world, dim, nf, want_fm):
    import __graft_entry__
    __graft_entry__.build()
    from fuxictr_b200 import layers, sharded as SH
    from fuxictr_b200.schema import FeatureMap
    specs = [("C%d" % i, {"type": "categorical", "source": "", "padding_idx": 0, "vocab_size": 37 + 5 * i})
             for i in range(nf)]
    fm = FeatureMap.from_specs(specs, embedding_dim=dim)
    torch.manual_seed(world * 1000 + dim)
    emb_layer = layers.FeatureEmbedding(fm, dim, embedding_initializer="partial(nn.init.normal_, std=0.1)").cuda()
    fmach = layers.FactorizationMachine(fm).cuda()
    with torch.no_grad():
        for p in fmach.parameters():
            p.normal_(0, 0.3)
        for m in fmach.modules():
            if isinstance(m, torch.nn.Embedding):
                m.weight[0].zero_()
    B_l = 96
    B = B_l * world
    gen = torch.Generator().manual_seed(3)
    ids = torch.cat([torch.randint(0, s["vocab_size"], (B, 1), generator=gen) for _, s in specs], dim=1)
    mat = torch.cat([ids.double(), torch.zeros(B, 1, dtype=torch.float64)], dim=1).cuda()
    W = mat.shape[1]
    gx = torch.randn(B, nf, dim, generator=gen).cuda()
    gl = torch.randn(B, 1, generator=gen).cuda()

    # ---- unsharded reference on the global batch (loss = mean over ranks => grads scaled by 1/world)
    X = OrderedDict((k, v) for k, v in fm.batch_dict(mat).items() if k != "label")
    emb_ref, logit_ref = layers.fused_front(emb_layer, fmach.lr_layer, X, want_fm)
    ((emb_ref * gx).sum() + (logit_ref * gl).sum()).mul(1.0 / world).backward()
    fed = emb_layer.embedding_layer
    lfed = fmach.lr_layer.embedding_layer.embedding_layer
    names = list(fm.features.keys())

    # ---- virtual ranks
    registry = {}
    fronts, shard_tabs, shard_lr = [], [], []
    for r in range(world):
        group = SH.VirtualPeerGroup(r, world, registry)
        etabs = [SH.shard_rows(fed.embedding_layers[f].weight.detach(), r, world) for f in names]
        ltabs = [SH.shard_rows(lfed.embedding_layers[f].weight.detach(), r, world) for f in names]
        shard_tabs.append(etabs)
        shard_lr.append(ltabs)
        fronts.append(SH.ShardedFront(group, names, etabs, ltabs, [fed.embedding_layers[f].num_embeddings for f in names],
                                      [fm.get_column_index(f) for f in names], [0] * nf, dim, B_l, W, torch.float64,
                                      bias=fmach.lr_layer.bias.detach(), want_fm=want_fm))
    for r, fr in enumerate(fronts):
        fr.phase_ids(mat[r * B_l:(r + 1) * B_l])
    for fr in fronts:
        fr.phase_push()
    outs = [fr.phase_reduce() for fr in fronts]
    emb_all = torch.cat([o[0] for o in outs]).view(B, nf, dim)
    logit_all = torch.cat([o[1] for o in outs])
    assert torch.equal(emb_all, emb_ref.detach())                  # pure copies over "NVLink": bit-exact
    assert close(logit_all, logit_ref, RTOL)
    # backward
    for r, fr in enumerate(fronts):
        emb_r, _, sums_r = outs[r]
        fr.phase_gprep(gx[r * B_l:(r + 1) * B_l].reshape(B_l, -1).contiguous(), emb_r, sums_r,
                       gl[r * B_l:(r + 1) * B_l].reshape(-1).contiguous())
    egrads = [[torch.zeros_like(t) for t in shard_tabs[r]] for r in range(world)]
    lgrads = [[torch.zeros_like(t) for t in shard_lr[r]] for r in range(world)]
    for r, fr in enumerate(fronts):
        fr.phase_pull(egrads[r], lgrads[r])
    torch.cuda.synchronize()
    for i, f in enumerate(names):
        full = SH.unshard_rows([egrads[r][i] for r in range(world)], fed.embedding_layers[f].num_embeddings)
        ref = fed.embedding_layers[f].weight.grad
        assert close(full, ref, RTOL, atol=RTOL * float(ref.abs().max())), f
        assert float(full[0].abs().sum()) == 0.0                   # padding row
        full_lr = SH.unshard_rows([lgrads[r][i] for r in range(world)], lfed.embedding_layers[f].num_embeddings)
        ref_lr = lfed.embedding_layers[f].weight.grad
        assert close(full_lr, ref_lr, RTOL, atol=RTOL * float(ref_lr.abs().max()) + 1e-12), f


def test_shard_roundtrip_gpu():
    from fuxictr_b200 import sharded as SH
    w = torch.randn(101, 8, device="cuda")
    for world in (1, 2, 3, 8):
        shards = [SH.shard_rows(w, r, world) for r in range(world)]
        assert [s.shape[0] for s in shards] == [SH.local_rows(101, r, world) for r in range(world)]
        assert torch.equal(SH.unshard_rows(shards, 101), w)


@pytest.mark.parametrize("world", [2, 8])
def test_virtual_ranks_embedding_only(world):
    """DLRM-style front: embeddings only (no LR tables, no FM term) over row shards."""
    import __graft_entry__
    __graft_entry__.build()
    from fuxictr_b200 import layers, sharded as SH
    from fuxictr_b200.schema import FeatureMap
    nf, dim, B_l = 26, 16, 40
    specs = [("C%d" % i, {"type": "categorical", "source": "", "padding_idx": 0, "vocab_size": 90 + 11 * i})
             for i in range(nf)]
    fm = FeatureMap.from_specs(specs, embedding_dim=dim)
    torch.manual_seed(world)
    emb_layer = layers.FeatureEmbedding(fm, dim, embedding_initializer="partial(nn.init.normal_, std=0.1)").cuda()
    B = B_l * world
    gen = torch.Generator().manual_seed(4)
    ids = torch.cat([torch.randint(0, s["vocab_size"], (B, 1), generator=gen) for _, s in specs], dim=1)
    mat = torch.cat([ids.double(), torch.zeros(B, 1, dtype=torch.float64)], dim=1).cuda()
    gx = torch.randn(B, nf, dim, generator=gen).cuda()
    X = OrderedDict((k, v) for k, v in fm.batch_dict(mat).items() if k != "label")
    emb_ref = emb_layer(X)
    (emb_ref * gx).sum().mul(1.0 / world).backward()
    fed = emb_layer.embedding_layer
    names = list(fm.features.keys())
    registry, fronts, shard_tabs = {}, [], []
    for r in range(world):
        etabs = [SH.shard_rows(fed.embedding_layers[f].weight.detach(), r, world) for f in names]
        shard_tabs.append(etabs)
        fronts.append(SH.ShardedFront(SH.VirtualPeerGroup(r, world, registry), names, etabs, None,
                                      [fed.embedding_layers[f].num_embeddings for f in names],
                                      [fm.get_column_index(f) for f in names], [0] * nf, dim, B_l, mat.shape[1],
                                      torch.float64, bias=None, want_fm=False))
    for r, fr in enumerate(fronts):
        fr.phase_ids(mat[r * B_l:(r + 1) * B_l])
    for fr in fronts:
        fr.phase_push()
    outs = [fr.phase_reduce() for fr in fronts]
    assert torch.equal(torch.cat([o[0] for o in outs]).view(B, nf, dim), emb_ref.detach())
    zeros = torch.zeros(B_l, device="cuda")
    for r, fr in enumerate(fronts):
        fr.phase_gprep(gx[r * B_l:(r + 1) * B_l].reshape(B_l, -1).contiguous(), outs[r][0], None, zeros)
    egrads = [[torch.zeros_like(t) for t in shard_tabs[r]] for r in range(world)]
    for r, fr in enumerate(fronts):
        fr.phase_pull(egrads[r], None)
    torch.cuda.synchronize()
    for i, f in enumerate(names):
        full = SH.unshard_rows([egrads[r][i] for r in range(world)], fed.embedding_layers[f].num_embeddings)
        ref = fed.embedding_layers[f].weight.grad
        assert close(full, ref, RTOL, atol=RTOL * float(ref.abs().max())), f
