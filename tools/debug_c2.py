"""Diagnostic (not a test): where does the fused C2 step leave the oracle?"""
import sys, os
from collections import OrderedDict
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__; __graft_entry__.build()
from fuxictr_b200 import zoo, arena
from oracle import fuxictr_oracle as O
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from test_gpu_parity import criteo_shape

fm, specs, mat = criteo_shape()
torch.manual_seed(2019)
model = zoo.DeepFM(fm, gpu=-1, embedding_dim=16, hidden_units=[300, 300, 300])
with torch.no_grad():
    for m in model.modules():
        if isinstance(m, torch.nn.Embedding):
            m.weight[1:].normal_(0, 0.05)
state0 = OrderedDict((k, v.detach().clone()) for k, v in model.state_dict().items())
tr = O.OracleTrainer(state0, lambda s, X: torch.sigmoid(O.deepfm_logit(specs, s, X, 3)), specs, ["label"])
model.device = torch.device("cuda:0"); model.model_to_device(); model.compile("adam", "binary_crossentropy", 1e-3)
model.use_fused_optimizer()
batch = fm.batch_dict(mat.cuda())
opt = model._fused_optimizer
for step in range(2):
    # oracle: grads then step
    tr.optimizer.zero_grad()
    yp, y = tr.forward(fm.batch_dict(mat)); loss_ref = O.bce_mean(yp, y); loss_ref.backward()
    gref = {k: v.grad.clone() for k, v in tr.state.items() if v.grad is not None}
    torch.nn.utils.clip_grad_norm_(tr.params, 10.0); tr.optimizer.step()
    # ours: same, but look at G before the optimizer
    opt.zero_grad()
    from fuxictr_b200 import functional as F2
    loss, _ = F2.logit_bce(model.get_labels(batch), *model.forward_logits(batch))
    loss.backward()
    named = dict(model.named_parameters())
    worst = []
    for k, g in gref.items():
        mine = named[k].grad
        d = float((mine.cpu() - g).abs().max()); s = float(g.abs().max())
        worst.append((d / max(s, 1e-30), k, d, s))
    worst.sort(reverse=True)
    print("step", step, "loss", float(loss), float(loss_ref), "worst grad rel:", worst[:3])
    tot = float(sum((named[k].grad.double() ** 2).sum() for k in gref) ** 0.5)
    print("  grad norm ours %.6e ref %.6e sumsq-arena %.6e" % (tot, float(sum((g.double() ** 2).sum() for g in gref.values()) ** 0.5), float((model._arena.G.double() ** 2).sum() ** 0.5)))
    opt.step()
    torch.cuda.synchronize()
    w = []
    for k, v in model.state_dict().items():
        r = tr.state[k].detach()
        d = (v.cpu() - r).abs()
        w.append((float(d.max()) / max(float(r.abs().max()), 1e-30), k, float(d.max()), int((d > 1e-5).sum()), d.numel()))
    w.sort(reverse=True)
    print("  worst weights:", w[:4])
