"""world_size-2 `gloo` tests (CPU) of the multi-GPU host logic: row sharding round trip through
a real process group, the composite gradient norm of sharded + replicated parameters, and the
dense-slice averaging — the arithmetic FusedAdam's sharded mode performs with NCCL."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

sys.path.insert(0, ROOT)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fuxictr_b200 import sharded as SH
        torch.manual_seed(0)
        vocab, dim = 53, 4
        full = torch.randn(vocab, dim)                     # same on every rank (same seed)
        mine = SH.shard_rows(full, rank, world)
        assert mine.shape[0] == SH.local_rows(vocab, rank, world)
        # unshard through the process group: pad to the max shard size, all_gather, trim
        n_max = SH.local_rows(vocab, 0, world)
        pad = torch.zeros(n_max, dim)
        pad[:mine.shape[0]] = mine
        gathered = [torch.zeros(n_max, dim) for _ in range(world)]
        dist.all_gather(gathered, pad)
        shards = [g[:SH.local_rows(vocab, r, world)] for r, g in enumerate(gathered)]
        assert torch.equal(SH.unshard_rows(shards, vocab), full)
        # row ownership: every global row has exactly one owner and the right local index
        for row in range(vocab):
            owner, lrow = row % world, row // world
            if owner == rank:
                assert torch.equal(mine[lrow], full[row])
        # composite clip norm: shard parts summed over ranks + replicated dense part once
        g_full = torch.randn(vocab, dim)
        g_dense_local = torch.randn(7) * (rank + 1)        # per-rank dense grads (differ before all-reduce)
        g_shard = SH.shard_rows(g_full, rank, world)
        dense = g_dense_local.clone()
        dist.all_reduce(dense)
        dense /= world                                     # mean over the global batch
        sumsq = (g_shard ** 2).sum()
        dist.all_reduce(sumsq)
        sumsq = sumsq + (dense ** 2).sum()
        ref_dense = sum(torch.randn(7).new_tensor(g_dense_local / (rank + 1)) * (r + 1) for r in range(world)) / world
        ref = (g_full ** 2).sum() + (ref_dense ** 2).sum()
        assert abs(float(sumsq) - float(ref)) <= 1e-5 * float(ref)
        out.put((rank, "ok"))
    except Exception as exc:  # surface the failure in the parent
        out.put((rank, repr(exc)))
    finally:
        dist.destroy_process_group()


def test_sharding_logic_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
