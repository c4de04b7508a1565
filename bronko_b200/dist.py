"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing; NCCL over NVLink on the GPU box).

Two ways the path spreads over the GPUs of one box:

* sample-per-GPU (BASELINE configs C4/C5): samples are independent — `shard_samples` gives every rank its
  samples, no collective is involved (bench.py --gpus N measures this).
* read-sharded deep sample (C3).  The reference's pileup is a MAX over globally summed, thresholded,
  saturated k-mer counts (src/call.rs:1172-1173, 1341-1345; R1 / R2 separately, 302-317), so per-rank pileups
  cannot simply be all-reduced: counts are merged first (all-reduce SUM of the dense reference-k-mer counts
  + all-to-all of the novel (k-mer, partial count) pairs to their owner rank), every k-mer is then
  thresholded and mapped by exactly one rank, and only then are depth (MAX), support (SUM) and the
  tallies (SUM) combined.

  On the GPU all of that — kernels AND collectives, enqueued on the context's stream — lives behind the C ABI
  (bk_shard_init + bk_sample_finish; bronko_b200/csrc/bk_shard.inc).  The host only has to give every rank
  the NCCL unique id: `init_sharded` does it through torch.distributed.  `ShardedLocal` runs the same sharded
  kernels with all ranks in one process on one device (the in-process transport; single-GPU tests).

  `finish_sharded` is the protocol itself, engine-agnostic, in Python: the CPU tests drive it over gloo with an
  engine built on the CPU checker (tests/test_dist_cpu.py) to pin down what the C implementation must do — and that
  the tempting shortcut (threshold per shard, all-reduce pileups) is wrong.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L

OWNER_UNITS_LOG2 = 6      # bk_bins.cuh: owner ranks split the hash space in 64 units


def shard_samples(n_samples, rank, world):
    """Round-robin sample-per-GPU assignment (no collective on the data path)."""
    return list(range(rank, n_samples, world))


def split_reads(bases, off, rank, world):
    """This rank's contiguous share of one file's reads: (bases view, offsets rebased to 0)."""
    n = len(off) - 1
    r0, r1 = n * rank // world, n * (rank + 1) // world
    b0, b1 = int(off[r0]), int(off[r1])
    return bases[b0:b1], (off[r0:r1 + 1].astype(np.int64) - b0).astype(np.uint32)


def owner_of(kmers, world):
    """Owner rank of novel k-mers (numpy uint64) — bk_bins.cuh: the top 6 bits of bin_hash() are a unit, rank r owns
    units [ceil(64 r / n), ceil(64 (r + 1) / n))."""
    k = np.asarray(kmers, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = (k ^ (k >> np.uint64(31))) * np.uint64(0x9E3779B97F4A7C15)
    unit = (h >> np.uint64(64 - OWNER_UNITS_LOG2)).astype(np.int64)
    lo = np.array([(r * (1 << OWNER_UNITS_LOG2) + world - 1) // world for r in range(world + 1)], dtype=np.int64)
    return np.searchsorted(lo, unit, side="right") - 1


# ---- GPU: the sharded finish lives in the library ----------------------------------------------------------------

def init_sharded(ctx, group=None):
    """Make `ctx` one rank of a read-sharded group spanning the torch.distributed group (one rank per process and
    GPU).  Rank 0 draws the NCCL unique id, the group broadcasts it, every rank calls bk_shard_init.  Afterwards
    ctx.begin / push* / finish process this rank's SHARE of one sample; finish() must be called by every rank and
    returns the same Sample everywhere.  Undo with leave_sharded(ctx)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lib = ctx._lib
    idbuf = (C.c_uint8 * 128)()
    if rank == 0:
        rc = lib.bk_shard_unique_id(idbuf)
        if rc != 0:
            raise L.BkError(rc, lib.bk_last_error(None).decode())
    holder = [bytes(idbuf)]
    dist.broadcast_object_list(holder, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    idbuf = (C.c_uint8 * 128).from_buffer_copy(holder[0])
    ctx._check(lib.bk_shard_init(ctx.h, rank, world, idbuf))
    return rank, world


def leave_sharded(ctx):
    ctx._check(ctx._lib.bk_shard_init(ctx.h, 0, 1, None))


def call_sample_sharded(ctx, files, args=None):
    """files: this rank's share of the reads, [(bases, offsets)] or [(r1...), (r2...)], on a context prepared by
    init_sharded.  Returns the Sample (identical on every rank)."""
    ctx.begin(args)
    for slot, (bases, off) in enumerate(files):
        ctx.push(slot, bases, off)
    return ctx.finish()


class ShardedLocal:
    """n ranks of one read-sharded sample as n contexts of THIS process on one device (bk_shard_local): the sharded
    kernels and the protocol are the library's, the collectives are kernels / copies instead of NCCL."""

    def __init__(self, owner_ctx, n):
        from .api import Bronko
        self.ctxs = [owner_ctx]
        for _ in range(n - 1):
            c = Bronko(owner_ctx.device)
            c.share_index(owner_ctx)
            self.ctxs.append(c)
        self._own = self.ctxs[1:]
        self.lib = owner_ctx._lib
        self._arr = (C.c_void_p * n)(*[c.h for c in self.ctxs])
        owner_ctx._check(self.lib.bk_shard_local(self._arr, n))

    def call_sample(self, files_per_rank, args=None):
        """files_per_rank[r] = rank r's share, [(bases, offsets)] (+ R2).  Returns rank 0's Sample."""
        from .api import Sample
        for c, files in zip(self.ctxs, files_per_rank):
            c.begin(args)
            for slot, (bases, off) in enumerate(files):
                c.push(slot, bases, off)
        res = L.SampleResult()
        self.ctxs[0]._check(self.lib.bk_shard_finish_local(self._arr, len(self.ctxs), C.byref(res)))
        return Sample(self.ctxs[0], res)

    def sample_of(self, r):
        """The Sample rank r ended with (every rank must hold the same result)."""
        from .api import Sample
        c = self.ctxs[r]
        res = L.SampleResult()
        c._check(self.lib.bk_sample_result_get(c.h, C.byref(res)))
        return Sample(c, res)

    def close(self):
        one = (C.c_void_p * 1)()
        for c in self.ctxs:
            one[0] = c.h
            self.lib.bk_shard_local(one, 1)
        for c in self._own:
            c.close()
        self._own = []


# ---- the protocol, engine-agnostic (CPU tests) ---------------------------------------------------------------------

def all_reduce(t, op, group=None):
    if t.numel():
        dist.all_reduce(t, op=op, group=group)


def exchange_novel(kmers, counts, part_off, group=None):
    """All-to-all of the (k-mer, count) pairs grouped by owner rank; returns the pairs this rank owns
    (duplicates across ranks are summed when the owner merges them)."""
    world = dist.get_world_size(group)
    send = torch.tensor([part_off[r + 1] - part_off[r] for r in range(world)], dtype=torch.int64)
    recv = torch.zeros(world, dtype=torch.int64)
    dist.all_to_all_single(recv, send, output_split_sizes=[1] * world, input_split_sizes=[1] * world, group=group)
    send_l, recv_l = [int(x) for x in send.tolist()], [int(x) for x in recv.tolist()]
    out_k = torch.empty(sum(recv_l), dtype=kmers.dtype)
    out_c = torch.empty(sum(recv_l), dtype=counts.dtype)
    dist.all_to_all_single(out_k, kmers.contiguous(), output_split_sizes=recv_l, input_split_sizes=send_l, group=group)
    dist.all_to_all_single(out_c, counts.contiguous(), output_split_sizes=recv_l, input_split_sizes=send_l, group=group)
    return out_k, out_c


def finish_sharded(engine, n_files, group=None):
    """The collective part of one read-sharded sample, after every rank counted its reads (what
    bronko_b200/csrc/bk_shard.inc: shard_finish does on the GPU)."""
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
