"""Multi-GPU host logic (one process per GPU, torch.distributed; NCCL over NVLink on the GPU box).

Two ways the path spreads over the GPUs of one box:

* sample-per-GPU (BASELINE configs C4/C5): samples are independent — `shard_samples` gives every rank its
  samples, no collective is involved (bench.py --gpus N measures this).
* read-sharded deep sample (C3): `call_sample_sharded`.  The reference's pileup is a MAX over globally
  summed, thresholded, saturated k-mer counts (src/call.rs:1172-1173, 1342-1343), so per-rank pileups
  cannot simply be all-reduced: counts are merged first (all-reduce SUM of the dense reference-k-mer
  counts + all-to-all of the novel (k-mer, count) pairs to their owner rank), every k-mer is then
  thresholded and mapped by exactly one rank, and only then are depth (MAX), support (SUM) and the
  tallies (SUM) combined.

The collectives run on an "engine" — `GpuShardEngine` wraps the bk_shard_* C ABI and hands out zero-copy
torch views of the library's device buffers; tests drive the same orchestration over gloo with a CPU
engine.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L


def shard_samples(n_samples, rank, world):
    """Round-robin sample-per-GPU assignment (no collective on the data path)."""
    return list(range(rank, n_samples, world))


def split_reads(bases, off, rank, world):
    """This rank's contiguous share of one file's reads: (bases view, offsets rebased to 0)."""
    n = len(off) - 1
    r0, r1 = n * rank // world, n * (rank + 1) // world
    b0, b1 = int(off[r0]), int(off[r1])
    return bases[b0:b1], (off[r0:r1 + 1].astype(np.int64) - b0).astype(np.uint32)


class _DevArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _view(ptr, n, typestr, device):
    if n == 0 or not ptr:
        return torch.empty(0, dtype={"<i4": torch.int32, "<i8": torch.int64}[typestr], device=device)
    return torch.as_tensor(_DevArray(ptr, n, typestr), device=device)


class GpuShardEngine:
    """bk_shard_* on one GPU.  All tensors are views of library-owned device memory (int32 / int64 views of
    u32 / u64 data: sums wrap identically, depths are <= 10^6 so signed MAX is the unsigned MAX)."""

    def __init__(self, ctx, rank, world):
        self.ctx, self.rank, self.world = ctx, rank, world
        self.lib = ctx._lib
        self.device = torch.device("cuda", ctx.device)
        ctx._check(self.lib.bk_shard_config(ctx.h, rank, world))
        self._keep = []

    def begin(self, file_slot):
        d_ref, n_ref, d_k, d_c = C.c_void_p(), C.c_uint64(), C.c_void_p(), C.c_void_p()
        part = (C.c_uint64 * (self.world + 1))()
        self.ctx._check(self.lib.bk_shard_begin(self.ctx.h, file_slot, C.byref(d_ref), C.byref(n_ref), C.byref(d_k), C.byref(d_c), part))
        part = [int(x) for x in part]
        return (_view(d_ref.value, n_ref.value, "<i4", self.device), _view(d_k.value, part[-1], "<i8", self.device),
                _view(d_c.value, part[-1], "<i4", self.device), part)

    def import_novel(self, file_slot, kmers, counts):
        kmers, counts = kmers.contiguous(), counts.contiguous()
        torch.cuda.synchronize(self.device)
        self.ctx._check(self.lib.bk_shard_import_novel(self.ctx.h, file_slot, kmers.data_ptr() if kmers.numel() else None,
                                                       counts.data_ptr() if counts.numel() else None, kmers.numel()))

    def map_stats(self):
        t0, t1, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        partial = (L.KmcStats * 2)()
        self.ctx._check(self.lib.bk_shard_map_stats(self.ctx.h, C.byref(t0), C.byref(t1), C.byref(n), partial))
        tallies = [_view(t0.value, n.value, "<i4", self.device)]
        if t1.value:
            tallies.append(_view(t1.value, n.value, "<i4", self.device))
        kmc = torch.tensor([[p.total_reads, p.total_kmers, p.unique_kmers, p.unique_counted] for p in partial],
                           dtype=torch.int64, device=self.device)
        return tallies, kmc

    def select_pileup(self, kmc_global):
        g = (L.KmcStats * 2)()
        for f in range(2):
            g[f].total_reads, g[f].total_kmers, g[f].unique_kmers, g[f].unique_counted = [int(x) for x in kmc_global[f].tolist()]
        d_pile, n = C.c_void_p(), C.c_uint64()
        torch.cuda.synchronize(self.device)
        self.ctx._check(self.lib.bk_shard_select_pileup(self.ctx.h, g, C.byref(d_pile), C.byref(n)))
        return _view(d_pile.value, 4 * n.value, "<i4", self.device).view(4, n.value)

    def score(self):
        from .api import Sample
        torch.cuda.synchronize(self.device)
        res = L.SampleResult()
        self.ctx._check(self.lib.bk_shard_score(self.ctx.h, C.byref(res)))
        return Sample(self.ctx, res)


def _staged(t, group):
    """gloo has no CUDA all-to-all: with a gloo group, collectives on device tensors go through the host
    (used by the single-GPU functional test; NCCL groups operate on the device buffers directly)."""
    return t.is_cuda and dist.get_backend(group) == "gloo"


def all_reduce(t, op, group=None):
    if t.numel() == 0:
        return
    if _staged(t, group):
        c = t.cpu()
        dist.all_reduce(c, op=op, group=group)
        t.copy_(c)
    else:
        dist.all_reduce(t, op=op, group=group)


def all_to_all(out, inp, out_splits, in_splits, group=None):
    if _staged(inp, group):
        co = torch.empty(out.shape, dtype=out.dtype)
        dist.all_to_all_single(co, inp.cpu().contiguous(), output_split_sizes=out_splits, input_split_sizes=in_splits, group=group)
        out.copy_(co)
    else:
        dist.all_to_all_single(out, inp.contiguous(), output_split_sizes=out_splits, input_split_sizes=in_splits, group=group)


def exchange_novel(kmers, counts, part_off, group=None):
    """All-to-all of the (k-mer, count) pairs grouped by owner rank; returns the pairs this rank owns
    (duplicates across ranks are summed when the owner re-inserts them)."""
    world = dist.get_world_size(group)
    send = torch.tensor([part_off[r + 1] - part_off[r] for r in range(world)], dtype=torch.int64, device=kmers.device)
    recv = torch.zeros(world, dtype=torch.int64, device=kmers.device)
    all_to_all(recv, send, [1] * world, [1] * world, group)
    send_l, recv_l = [int(x) for x in send.tolist()], [int(x) for x in recv.tolist()]
    out_k = torch.empty(sum(recv_l), dtype=kmers.dtype, device=kmers.device)
    out_c = torch.empty(sum(recv_l), dtype=counts.dtype, device=counts.device)
    all_to_all(out_k, kmers, recv_l, send_l, group)
    all_to_all(out_c, counts, recv_l, send_l, group)
    return out_k, out_c


def finish_sharded(engine, n_files, group=None):
    """The collective part of one read-sharded sample, after every rank pushed its reads."""
    for f in range(n_files):
        ref_counts, nk, nc, part = engine.begin(f)
        all_reduce(ref_counts, dist.ReduceOp.SUM, group)                     # dense counts of reference k-mers
        mk, mc = exchange_novel(nk, nc, part, group)                          # novel k-mers go to their owner
        engine.import_novel(f, mk, mc)
    tallies, kmc = engine.map_stats()
    for t in tallies:
        all_reduce(t, dist.ReduceOp.SUM, group)
    all_reduce(kmc, dist.ReduceOp.SUM, group)
    pile = engine.select_pileup(kmc)
    all_reduce(pile[0:2], dist.ReduceOp.MAX, group)                           # depth = max over k-mers (Q3)
    all_reduce(pile[2:4], dist.ReduceOp.SUM, group)                           # support = number of hits (Q4)
    return engine.score()


def call_sample_sharded(ctx, files, args=None, group=None):
    """files: this rank's share of the reads, [(bases, offsets)] or [(r1...), (r2...)].  Returns the Sample
    (identical on every rank)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    engine = GpuShardEngine(ctx, rank, world)
    ctx.begin(args)
    for slot, (bases, off) in enumerate(files):
        ctx.push(slot, bases, off)
    return finish_sharded(engine, len(files), group)
