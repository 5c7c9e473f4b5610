"""torchrun script (N GPUs, REAL ranks over NVLink peer memory + NCCL): a row-sharded model trained for 3
steps on its slice of the global batch must equal the CPU ORACLE (oracle/fuxictr_oracle.py, the
restatement of the reference's BaseModel.train_step) run on the whole global batch: per-step global
mean loss, this rank's table shards and the replicated dense weights, within 1e-5 relative.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_sharded_check.py \
        [--model DeepFM|DLRM] [--precision fp32|tf32x3] [--batch-local 64]

Checker use of oracle/ only (tests/test_gpu_multirank.py launches this script).
"""
import argparse
import os
import sys
from collections import OrderedDict

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--model", default="DeepFM", choices=["DeepFM", "DLRM"])
ap.add_argument("--precision", default="fp32", choices=["fp32", "tf32x3"])
ap.add_argument("--batch-local", type=int, default=64)
args = ap.parse_args()

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import __graft_entry__  # noqa: E402

if rank == 0:
    __graft_entry__.build()
dist.barrier()
from fuxictr_b200 import zoo, sharded as SH, functional as F2  # noqa: E402
from fuxictr_b200.schema import FeatureMap  # noqa: E402
from oracle import fuxictr_oracle as O  # noqa: E402

F2.set_matmul_precision(args.precision)
NF, D, B_l = (12, 8, args.batch_local) if args.model == "DeepFM" else (26, 16, args.batch_local)
specs = [("C%d" % i, {"type": "categorical", "source": "", "padding_idx": 0, "vocab_size": 50 + 37 * i})
         for i in range(NF)]
spec_map = OrderedDict(specs)
fm = FeatureMap.from_specs(specs, embedding_dim=D)
HID = [32, 16]


def make_model():
    torch.manual_seed(7)
    if args.model == "DeepFM":
        m = zoo.DeepFM(fm, gpu=local, embedding_dim=D, hidden_units=HID)
    else:
        m = zoo.DLRM(fm, gpu=local, embedding_dim=D, top_mlp_units=HID, interaction_op="dot")
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.Embedding):
                mod.weight[1:].normal_(0, 0.1)
    return m


gen = torch.Generator().manual_seed(11)
batches = []
for step in range(3):
    ids = torch.cat([torch.randint(0, s["vocab_size"], (B_l * world, 1), generator=gen) for _, s in specs], 1)
    lab = (torch.rand(B_l * world, 1, generator=gen) < 0.3)
    batches.append(torch.cat([ids.double(), lab.double()], 1))

# the checker: the reference's train_step restated on CPU, on the GLOBAL batch
model = make_model()
state0 = OrderedDict((k, v.detach().cpu().clone()) for k, v in model.state_dict().items())
if args.model == "DeepFM":
    pred = lambda s, X: torch.sigmoid(O.deepfm_logit(spec_map, s, X, len(HID)))      # noqa: E731
else:
    pred = lambda s, X: O.dlrm_pred(spec_map, s, X, len(HID))                        # noqa: E731
trainer = O.OracleTrainer(state0, pred, spec_map, ["label"])
ref_losses = [float(trainer.train_step(fm.batch_dict(b))) for b in batches]

model.enable_sharding(SH.SymmPeerGroup(), B_l, NF + 1, torch.float64, want_fm=(args.model == "DeepFM"))
model.use_fused_optimizer()
model.train()
losses = []
mine = [b[rank * B_l:(rank + 1) * B_l].contiguous().cuda() for b in batches]
for b in mine:
    losses.append(model.fused_train_step(fm.batch_dict(b)).detach().clone())
lt = torch.stack(losses)
dist.all_reduce(lt)
lt /= world          # global mean loss


def rel(a, b):
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


tol = 1e-5
ok = True
err = rel(lt.cpu(), torch.tensor(ref_losses))
ok &= err < tol
print("[r%d] %s loss err %.2e" % (rank, args.model, err), flush=True)
worst = 0.0
for k, v in model.state_dict().items():
    r = trainer.state[k].detach()
    if not r.dtype.is_floating_point:       # frozen index buffers (triu masks): identical or wrong
        ok &= bool(torch.equal(v.cpu(), r))
        continue
    if "embedding_layers" in k:
        r = SH.shard_rows(r, rank, world)
    e = rel(v.cpu(), r)
    worst = max(worst, e)
    if e >= tol:
        print("[r%d] MISMATCH %s %.2e" % (rank, k, e), flush=True)
        ok = False
print("[r%d] worst weight err after 3 steps %.2e -> %s" % (rank, worst, "OK" if ok else "FAIL"), flush=True)
flag = torch.tensor([0 if ok else 1], device="cuda")
dist.all_reduce(flag)
dist.barrier()
os._exit(0 if int(flag.item()) == 0 else 1)
