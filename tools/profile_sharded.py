"""torchrun script (N GPUs): where does the row-sharded DeepFM C2 step spend its time?

Times, with CUDA events on the compute stream and all ranks aligned by a barrier before every
repetition, each phase of the sharded front in isolation (ids broadcast, push, reduce, gprep, pull,
the 4 signal-pad barriers), the NCCL all-reduce of the dense gradient tail, the Adam pass over this
rank's arena, and the whole captured step.  Rank 0 prints one JSON object (microseconds, max over ranks).

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/profile_sharded.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import __graft_entry__  # noqa: E402

if rank == 0:
    __graft_entry__.build()
dist.barrier()
import bench  # noqa: E402
from fuxictr_b200 import zoo, sharded as SH, functional as F2  # noqa: E402
from fuxictr_b200.schema import FeatureMap  # noqa: E402
from fuxictr_b200.pipeline import TrainPipeline  # noqa: E402

F2.set_matmul_precision("tf32x3")
B = 4096
fm = FeatureMap.from_specs(bench.make_specs(), embedding_dim=bench.DIM)
torch.manual_seed(2019)
model = zoo.DeepFM(fm, gpu=local, embedding_dim=bench.DIM, hidden_units=bench.HIDDEN)
front = model.enable_sharding(SH.SymmPeerGroup(), B, bench.NF + 1, torch.float64)
opt = model.use_fused_optimizer()
model.train()
mats = [m.cuda() for m in bench.make_batches(4, B, seed=1000 + rank)]
REPS = 50


def timed(fn, reps=REPS):
    """mean microseconds of fn() per repetition, max over ranks.  `reps` calls are captured into one CUDA
    graph (an eager ctypes call costs more host time than these kernels run), every rank replays it at
    the same moment (host barrier first), CUDA events bracket the replay."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    torch.cuda.synchronize()
    dist.barrier()
    g.replay()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


out = {"world": world, "batch_per_gpu": B}
g = front.group
out["barrier"] = timed(g.barrier)
out["ids_broadcast"] = timed(lambda: front.phase_ids(mats[0]))
front.phase_ids(mats[0]); g.barrier()
out["push"] = timed(front.phase_push)
front.phase_push(); g.barrier()
out["reduce"] = timed(front.phase_reduce)
emb, logit, sums = front.phase_reduce()
gx = torch.randn(B, front.F * front.dim, device="cuda")
gl = torch.randn(B, device="cuda")
out["gprep"] = timed(lambda: front.phase_gprep(gx, emb, sums, gl))
front.phase_gprep(gx, emb, sums, gl); g.barrier()
egr = [torch.zeros_like(t) for t in front.emb_tables]
lgr = [torch.zeros_like(t) for t in front.lr_tables]
out["pull"] = timed(lambda: front.phase_pull(egr, lgr))
arena = model._arena
tail = arena.G[arena.tail_offset:] if hasattr(arena, "tail_offset") else arena.G
out["dense_allreduce_elems"] = int(tail.numel())
out["dense_allreduce"] = timed(lambda: dist.all_reduce(tail))
out["optimizer_step"] = timed(opt.step)
pipe = TrainPipeline(model, B, bench.NF + 1, torch.float64, graph=False)
pipe.prime(mats[0])
pipe.capture(3)


def replay_steps(n=50):
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        pipe.step_device(mats[i % len(mats)])
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n * 1e3], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


replay_steps(10)
out["graph_step"] = replay_steps(50)
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
os._exit(0)
