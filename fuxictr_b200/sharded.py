"""Row-sharded embedding tables over the GPUs of one NVSwitch box.

Row r of every table (embedding AND LogisticRegression) lives on rank ``r % world`` at local row
``r // world``; the dense part (FM reduce, MLPs) is data parallel.  The lookup + exchange is one
kernel per direction over NVLink peer memory (csrc/shard.cu): owners PUSH the looked-up rows into
the requesting rank's buffer, and PULL the gradient rows back; no NCCL all-to-all, no
variable-size splits, CUDA-graph capturable.  The only NCCL traffic left is the all-reduce of the
dense-parameter slice of the gradient arena and one scalar for the global gradient norm.

A ``PeerGroup`` supplies peer-mapped buffers and the cross-rank barrier:
  * ``SymmPeerGroup``    torch symmetric memory (one process per GPU, torchrun)
  * ``VirtualPeerGroup`` N "virtual ranks" inside ONE process / ONE GPU — the kernels cannot tell a
    local pointer from a peer pointer, so the whole algorithm is testable on a single GPU
    (tests/test_gpu_sharded.py drives the phases of all virtual ranks in lock step).
"""
import ctypes

import torch

from . import _lib
from . import functional as F2
from ._lib import b2_field


# --------------------------------------------------------------------------------------------
# shard / unshard a (vocab, dim) table:  rank r keeps rows r, r+world, r+2*world, ...
# --------------------------------------------------------------------------------------------
def shard_rows(weight, rank, world):
    return weight[rank::world].contiguous()


def local_rows(vocab, rank, world):
    return (vocab - rank + world - 1) // world if vocab > rank else 0


def unshard_rows(shards, vocab):
    """Inverse of shard_rows given the list of all ranks' shards."""
    world = len(shards)
    full = shards[0].new_empty((vocab,) + tuple(shards[0].shape[1:]))
    for r, s in enumerate(shards):
        full[r::world] = s
    return full


# --------------------------------------------------------------------------------------------
# Peer groups
# --------------------------------------------------------------------------------------------
class PeerGroup(object):
    world = 1
    rank = 0

    def alloc(self, name, shape, dtype):
        """Returns (local tensor, [device pointer of that buffer on every rank], peer_view) where
        peer_view(r) is a tensor aliasing rank r's copy of the buffer."""
        raise NotImplementedError

    def barrier(self):
        raise NotImplementedError


class SymmPeerGroup(PeerGroup):
    """One process per GPU; buffers come from torch.distributed symmetric memory (NVLink P2P)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self._dist, self._symm = dist, symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self._handles = []

    def alloc(self, name, shape, dtype):
        t = self._symm.empty(tuple(shape), dtype=dtype, device=torch.device("cuda", torch.cuda.current_device()))
        hdl = self._symm.rendezvous(t, self.group.group_name)
        self._handles.append(hdl)
        t.zero_()
        shape = tuple(shape)
        return t, [int(p) for p in hdl.buffer_ptrs], (lambda r: hdl.get_buffer(r, shape, dtype))

    def barrier(self):
        self._handles[0].barrier()      # device-side signal-pad barrier on the current stream


class VirtualPeerGroup(PeerGroup):
    """`world` virtual ranks in one process: buffers are ordinary tensors shared through a dict;
    barriers are no-ops because the test harness runs each phase for all virtual ranks in order."""

    def __init__(self, rank, world, registry):
        self.rank, self.world, self._reg = rank, world, registry

    def alloc(self, name, shape, dtype):
        bufs = self._reg.setdefault(name, {})
        bufs[self.rank] = torch.zeros(tuple(shape), dtype=dtype, device="cuda")
        return bufs[self.rank], _LazyPtrs(bufs, self.world), (lambda r: bufs[r])

    def barrier(self):
        pass


class _LazyPtrs(object):
    """Pointer list that resolves once all virtual ranks have allocated."""

    def __init__(self, bufs, world):
        self._bufs, self._world = bufs, world

    def __iter__(self):
        return iter([self._bufs[r].data_ptr() for r in range(self._world)])

    def __getitem__(self, r):
        return self._bufs[r].data_ptr()


def _ptr_array(ptrs):
    ptrs = list(ptrs)
    return (ctypes.c_void_p * len(ptrs))(*ptrs)


# --------------------------------------------------------------------------------------------
# The sharded front: FeatureEmbedding (+ FM product_sum) (+ LogisticRegression) over row shards
# --------------------------------------------------------------------------------------------
class ShardedFront(object):
    """Holds the peer buffers and launches the phases.

    emb_tables / lr_tables: lists of this rank's SHARD parameters (one per field, fp32 CUDA);
    vocabs: global vocabulary sizes; columns: column of each field in the batch matrix."""

    def __init__(self, group, names, emb_tables, lr_tables, vocabs, columns, padding, dim, batch_local,
                 matrix_width, idx_dtype, bias=None, want_fm=True):
        self.group, self.names = group, list(names)
        self.emb_tables, self.lr_tables = list(emb_tables), (list(lr_tables) if lr_tables else None)
        self.vocabs, self.columns, self.padding = list(vocabs), list(columns), list(padding)
        self.dim, self.B, self.W = dim, batch_local, matrix_width
        self.F = len(self.names)
        self.bias, self.want_fm = bias, want_fm
        self.idx_dtype = idx_dtype
        self.idx_code = F2._IDX_CODE[idx_dtype]
        if max(self.vocabs) >= 2 ** 31:
            raise NotImplementedError("the id exchange narrows row numbers to int32 (vocabulary >= 2^31)")
        g = group
        # ids_all[p] = batch matrix of rank p: every rank BROADCASTS its ids into slot `rank` of all
        # peers (one launch of P2P stores), so the push kernel walks local memory only.
        # the slots hold int32 row numbers: the exchange narrows the collator's float64 on the fly
        self.ids_all, self._ids_all_ptrs, self._ids_peer = g.alloc("ids_all", (g.world, batch_local, matrix_width),
                                                                   torch.int32)
        esz = self.ids_all.element_size()
        self.src_code = self.idx_code          # dtype of the batch matrix handed to phase_ids
        self.idx_code = F2._IDX_CODE[torch.int32]
        self._slot_bytes = batch_local * matrix_width * esz
        self.ids_ptrs = [self.ids_all.data_ptr() + p * self._slot_bytes for p in range(g.world)]
        self._ids_src = torch.empty((batch_local, matrix_width), dtype=idx_dtype, device="cuda")
        # rows this rank served in the forward (filled by the push, walked by the pull): int32[4] entries
        self.owned_cap = g.world * batch_local * self.F
        self.owned = torch.empty((self.owned_cap, 4), dtype=torch.int32, device="cuda")
        self.owned_count = torch.zeros(1, dtype=torch.int32, device="cuda")
        self.emb, self.emb_ptrs, _ = g.alloc("emb", (batch_local, self.F * dim), torch.float32)
        self.lrw, self.lrw_ptrs, _ = g.alloc("lrw", (batch_local, self.F), torch.float32)
        self.gemb, self.gemb_ptrs, _ = g.alloc("gemb", (batch_local, self.F * dim), torch.float32)
        self.glogit, self.glogit_ptrs, _ = g.alloc("glogit", (batch_local,), torch.float32)
        self.status = torch.zeros(1, dtype=torch.int32, device="cuda")
        # mean over the GLOBAL batch (rank_model.py:130): either the pull scales every gradient row by
        # 1/world (default), or the caller seeds backward() with 1/world and sets pull_scale = 1
        self.pull_scale = 1.0 / g.world
        self.on_dense_grads_ready = None     # set by RankModel.use_fused_optimizer (overlapped dense all-reduce)

    # -- descriptors --------------------------------------------------------------------------
    def _descs(self, tables, dim):
        descs = (b2_field * self.F)()
        for d, t, v, c, pad in zip(descs, tables, self.vocabs, self.columns, self.padding):
            d.table = t.data_ptr() if t is not None else 0
            d.vocab, d.idx_stride, d.dim, d.seq_len, d.pool = v, c, dim, 1, 0
            d.padding_idx = -1 if pad is None else int(pad)
            d.idx, d.out, d.out_stride = 0, 0, 0
        return descs

    # -- forward phases -------------------------------------------------------------------------
    def phase_ids(self, batch_matrix):
        g = self.group
        src = batch_matrix
        if (not src.is_contiguous()) or src.data_ptr() % 16 != 0:
            self._ids_src.copy_(batch_matrix)
            src = self._ids_src
        dst = [int(base) + g.rank * self._slot_bytes for base in self._ids_all_ptrs]     # my slot on every rank
        _lib.call("b2_peer_bcast_ids", F2._ptr(src), self.src_code, self.B * self.W, _ptr_array(dst), g.world,
                  F2._stream())

    def phase_push(self):
        g = self.group
        lr = self._descs(self.lr_tables, 1) if self.lr_tables else None
        _lib.call("b2_shard_push", self._descs(self.emb_tables, self.dim), lr, self.F, self.B, g.world, g.rank,
                  _ptr_array(self.ids_ptrs), self.idx_code, self.W, _ptr_array(self.emb_ptrs),
                  _ptr_array(self.lrw_ptrs) if lr is not None else None, F2._ptr(self.status),
                  F2._ptr(self.owned), F2._ptr(self.owned_count), self.owned_cap, F2._stream())

    def phase_reduce(self):
        """Local: logit (B,1) and field sums from the landed rows.  Returns (emb copy, logit, sums)."""
        # The landed rows are consumed in place: the next overwrite of this peer buffer is the NEXT
        # step's push, which is ordered after this step's backward (and its closing barrier).
        emb = self.emb
        if not self.lr_tables and not self.want_fm:     # embeddings only (DLRM): nothing to reduce
            return emb, torch.zeros((self.B, 1), dtype=torch.float32, device="cuda"), None
        logit = torch.empty((self.B, 1), dtype=torch.float32, device="cuda")
        sums = torch.empty((self.B, self.dim), dtype=torch.float32, device="cuda") if self.want_fm else None
        _lib.call("b2_front_reduce", F2._ptr(emb), F2._ptr(self.lrw) if self.lr_tables else None,
                  F2._ptr(self.bias), self.B, self.F, self.dim, 1 if self.want_fm else 0, F2._ptr(logit),
                  F2._ptr(sums), F2._stream())
        return emb, logit, sums

    # -- backward phases ------------------------------------------------------------------------
    def phase_gprep(self, gx, emb, sums, glogit, gbias=None):
        _lib.call("b2_front_gprep", F2._ptr(gx), F2._ptr(emb), F2._ptr(sums), F2._ptr(glogit), self.B, self.F,
                  self.dim, 1 if self.want_fm else 0, F2._ptr(self.gemb),
                  F2._ptr(self.glogit) if glogit is not None else None, F2._ptr(gbias),
                  1 if (gbias is not None and F2._is_zeroed(gbias)) else 0, F2._stream())

    def phase_pull(self, emb_grads, lr_grads):
        g = self.group
        lr = self._descs(lr_grads, 1) if lr_grads else None
        _lib.call("b2_shard_pull", self._descs(emb_grads, self.dim), lr, self.F, self.B, g.world, g.rank,
                  _ptr_array(self.gemb_ptrs), _ptr_array(self.glogit_ptrs) if lr is not None else None,
                  self.pull_scale, F2._ptr(self.owned), F2._ptr(self.owned_count), self.owned_cap, F2._stream())


class _ShardedFrontFn(torch.autograd.Function):
    """(emb (B,F,D), logit (B,1)) with sharded tables; 2 barriers forward, 1 backward.

    No closing barrier: a peer buffer of this step is next written only behind the NEXT step's first
    barrier (ids visible), which no rank passes before every rank has finished this step's backward
    on its stream — and the pull walks the local owned-row list, not the peers' id matrices, so the
    early overwrite of `ids_all` by the next step's broadcast cannot race with it."""

    @staticmethod
    def forward(ctx, front, batch_matrix, bias, *tables):
        g = front.group
        front.phase_ids(batch_matrix)
        g.barrier()                      # every rank's ids are visible
        front.phase_push()
        g.barrier()                      # every owner's rows have landed here
        emb, logit, sums = front.phase_reduce()
        ctx.front, ctx.tables, ctx.bias = front, tables, bias
        ctx.save_for_backward(emb, sums)
        return emb.view(front.B, front.F, front.dim), logit

    @staticmethod
    def backward(ctx, gemb, glogit):
        front, tables, bias = ctx.front, ctx.tables, ctx.bias
        emb, sums = ctx.saved_tensors
        g = front.group
        gx = None if gemb is None else F2._f32c(gemb).view(front.B, -1)
        needs_logit = bool(front.lr_tables) or front.want_fm
        gl = None
        if needs_logit:
            gl = (torch.zeros(front.B, device="cuda") if glogit is None else F2._f32c(glogit).view(-1))
        gbias = None
        if bias is not None and bias.requires_grad:
            gbias = F2._grad_buffer(bias, zero=False)
        front.phase_gprep(gx, emb, sums, gl, gbias)      # also: LR bias gradient = sum_b glogit[b]
        if front.on_dense_grads_ready is not None:      # every dense gradient now exists: start their all-reduce
            front.on_dense_grads_ready()
        g.barrier()                      # every rank's gradient rows are ready to be pulled
        n = front.F
        egrads = [(F2._grad_buffer(t, zero=True) if t.requires_grad else None) for t in tables[:n]]
        lgrads = [(F2._grad_buffer(t, zero=True) if t.requires_grad else None) for t in tables[n:]]
        front.phase_pull(egrads, lgrads if front.lr_tables else None)
        return (None, None, gbias) + tuple(egrads) + tuple(lgrads)


def sharded_front(front, batch_matrix):
    tables = tuple(front.emb_tables) + tuple(front.lr_tables or ())
    return _ShardedFrontFn.apply(front, batch_matrix, front.bias, *tables)
