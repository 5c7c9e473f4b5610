"""torchrun script (N GPUs): a row-sharded DeepFM trained for 3 steps must equal the SAME model
trained unsharded on the concatenated global batch (each rank recomputes that reference locally):
logits, losses, the rank's table shards and the replicated dense weights, within 1e-5.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_sharded_check.py
"""
import os, sys
from collections import OrderedDict
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import __graft_entry__
if rank == 0:
    __graft_entry__.build()
dist.barrier()
from fuxictr_b200 import zoo, sharded as SH, functional as F2
from fuxictr_b200.schema import FeatureMap

NF, D, B_l = 12, 8, 64
specs = [("C%d" % i, {"type": "categorical", "source": "", "padding_idx": 0, "vocab_size": 50 + 3 * i}) for i in range(NF)]
fm = FeatureMap.from_specs(specs, embedding_dim=D)

def make_model():
    torch.manual_seed(7)
    m = zoo.DeepFM(fm, gpu=local, embedding_dim=D, hidden_units=[32, 16])
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
    batches.append(torch.cat([ids.double(), lab.double()], 1).cuda())

ref = make_model(); ref.use_fused_optimizer()
ref_losses = [float(ref.fused_train_step(fm.batch_dict(b)).detach()) for b in batches]

model = make_model()
model.enable_sharding(SH.SymmPeerGroup(), B_l, NF + 1, torch.float64)
model.use_fused_optimizer()
losses = []
for b in batches:
    mine = b[rank * B_l:(rank + 1) * B_l].contiguous()
    losses.append(model.fused_train_step(fm.batch_dict(mine)).detach().clone())
lt = torch.stack(losses); dist.all_reduce(lt); lt /= world          # global mean loss
def rel(a, b): return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)
ok = True
err = rel(lt.cpu(), torch.tensor(ref_losses)); ok &= err < 1e-5
print("[r%d] loss err %.2e" % (rank, err), flush=True)
ref_sd = ref.state_dict()
worst = 0.0
for k, v in model.state_dict().items():
    r = ref_sd[k]
    if "embedding_layers" in k:
        r = SH.shard_rows(r, rank, world)
    e = rel(v, r); worst = max(worst, e)
    if e >= 1e-5: print("[r%d] MISMATCH %s %.2e" % (rank, k, e), flush=True); ok = False
print("[r%d] worst weight err %.2e -> %s" % (rank, worst, "OK" if ok else "FAIL"), flush=True)
flag = torch.tensor([0 if ok else 1], device="cuda"); dist.all_reduce(flag)
dist.barrier()
os._exit(0 if int(flag.item()) == 0 else 1)
