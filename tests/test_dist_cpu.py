"""world_size-2 gloo test of the read-sharded protocol (bronko_b200/dist.py: finish_sharded) with a CPU
engine built on the oracle: sharding the reads of one sample over ranks and merging counts BEFORE the
threshold must reproduce the unsharded oracle bit for bit — and the tempting shortcut (threshold per
shard, all-reduce pileups) must not."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K = 21


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def ref_kmer_ids(oindex):
    """Sorted distinct reference k-mers of both orientations — the dense 'reference id' space of a shard engine."""
    out = set()
    comp = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")
    code = {65: 0, 67: 1, 71: 2, 84: 3, 97: 0, 99: 1, 103: 2, 116: 3}
    for _, seqs in oindex.genomes():
        for _, _, raw in seqs:
            for s in (raw, raw.translate(comp)[::-1]):
                v = 0
                mask = (1 << (2 * K)) - 1
                for i, c in enumerate(s):
                    v = ((v << 2) | code.get(c, 0)) & mask
                    if i >= K - 1:
                        out.add(v)
    return np.array(sorted(out), dtype=np.uint64)


class OracleShardEngine:
    """Per-rank semantics of the bk_shard_* stages, on numpy + the oracle (CPU tensors)."""

    def __init__(self, O, oindex, params, files, rank, world):
        self.O, self.ix, self.p, self.rank, self.world = O, oindex, params, rank, world
        self.ids = ref_kmer_ids(oindex)
        self.files = files
        self.merged = {}
        self.total_reads = [len(off) - 1 for _, off in files]

    @staticmethod
    def owner(kmers, world):
        from bronko_b200.dist import owner_of            # the library's owner function (bk_bins.cuh: hash units)
        return owner_of(kmers, world)

    def begin(self, f):
        b, off = self.files[f]
        km, ct = self.O.Counts.count(K, b, off.astype(np.uint64), 1, 2 ** 62, 1).get()      # every k-mer, no cap
        pos = np.searchsorted(self.ids, km)
        pos[pos >= len(self.ids)] = 0
        is_ref = self.ids[pos] == km
        dense = np.zeros(len(self.ids), dtype=np.int32)
        dense[pos[is_ref]] = ct[is_ref].astype(np.int32)
        nk, nc = km[~is_ref], ct[~is_ref]
        own = self.owner(nk, self.world).astype(np.int64)
        order = np.argsort(own, kind="stable")
        part = [0] + np.cumsum(np.bincount(own, minlength=self.world)).tolist()
        self._dense = torch.from_numpy(dense)
        return self._dense, torch.from_numpy(nk[order].astype(np.int64)), torch.from_numpy(nc[order].astype(np.int32)), part

    def import_novel(self, f, kmers, counts):
        d = {}
        for k_, c_ in zip(kmers.numpy().astype(np.uint64).tolist(), counts.numpy().tolist()):
            d[k_] = d.get(k_, 0) + c_
        dense = self._dense.numpy().astype(np.uint64)
        owned = (np.arange(len(self.ids)) % self.world) == self.rank
        km = np.concatenate([self.ids[owned & (dense > 0)], np.array(sorted(d), dtype=np.uint64)])
        ct = np.concatenate([dense[owned & (dense > 0)], np.array([d[x] for x in sorted(d)], dtype=np.uint64)])
        self.merged[f] = (km, ct)

    def map_stats(self):
        ci, cs = int(self.p.min_kmers), 1000000
        self.counts, kmc = [], []
        for f in range(len(self.files)):
            km, ct = self.merged[f]
            keep = ct >= ci
            self.counts.append(self.O.Counts.from_list(km[keep], np.minimum(ct[keep], cs)))
            kmc.append([self.total_reads[f], int(ct.sum()), len(km), int(keep.sum())])
        while len(kmc) < 2:
            kmc.append([0, 0, 0, 0])
        self.sample = self.O.Sample(self.ix, self.p, self.counts, map_only=True)
        self.tallies = [torch.from_numpy(self.sample.stats(f).astype(np.int32).reshape(-1).copy()) for f in range(len(self.files))]
        return self.tallies, torch.tensor(kmc, dtype=torch.int64)

    def select_pileup(self, kmc):
        self.kmc = kmc
        st = [t.numpy().reshape(-1, 4) for t in self.tallies]
        lens = [sum(n for _, n, _ in seqs) for _, seqs in self.ix.genomes()]
        best, score = -1, 0.0
        for g in range(len(lens)):                      # pick_best_genome(_paired), src/call.rs:422-502
            if not any(s[g][3] for s in st):
                continue
            sc = float(sum(int(s[g][0]) for s in st)) / float(lens[g]) / 2.0
            if sc > score:
                best, score = g, sc
        self.best = best
        rows = max(lens)
        pile = np.zeros((4, rows * 4), dtype=np.int32)
        if best >= 0:
            p = self.sample.pileup(best)
            pile[:, :p.shape[1] * 4] = p.reshape(4, -1).astype(np.int32)
        self.pile = torch.from_numpy(pile)
        return self.pile

    def score(self):
        if self.best < 0:
            return None
        rows = self.O.lib().orc_sample_genome_rows(self.sample.h, self.best)
        arr = self.pile.numpy().astype(np.uint64)[:, :rows * 4].reshape(4, rows, 4)
        st = [t.numpy().astype(np.uint64) for t in self.tallies]
        self.sample.call_with(self.best, arr, st, [int(self.kmc[f][3]) for f in range(len(self.files))])
        return self.sample


def _worker(rank, world, port, depth, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from bronko_b200 import sim
    from bronko_b200.dist import finish_sharded, split_reads
    oi = O.Index.build(K, [sim.genome_path(sim.HPV16)])
    r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.HPV16), depth, 77)
    params = O.Params.defaults(k=K)
    files = [split_reads(r1, o1, rank, world), split_reads(r2, o2, rank, world)]
    eng = OracleShardEngine(O, oi, params, files, rank, world)
    s = finish_sharded(eng, 2)
    # the unsharded truth
    full = [O.Counts.count(K, b, o.astype(np.uint64), 3, 1000000, 1) for b, o in ((r1, o1), (r2, o2))]
    ref = O.Sample(oi, params, full)
    ok = (s.best == ref.best and s.variants().tobytes() == ref.variants().tobytes() and (s.pileup() == ref.pileup()).all()
          and s.summary() == ref.summary() and s.unmapped() == ref.unmapped()
          and all((s.stats(f)[:, :3] == ref.stats(f)[:, :3]).all() for f in range(2))
          and [int(x) for x in eng.kmc[0].tolist()] == list(full[0].stats()))
    # the naive protocol: threshold per shard, MAX the pileups — must differ (counts of 2+2 are lost, depths too low)
    naive = O.Sample(oi, params, [O.Counts.count(K, b, o.astype(np.uint64), 3, 1000000, 1) for b, o in files])
    t = torch.from_numpy(naive.pileup()[:2].astype(np.int64))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    naive_differs = bool((t.numpy() != ref.pileup()[:2].astype(np.int64)).any())
    q.put((rank, bool(ok), naive_differs, len(ref.variants())))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_read_sharded_protocol_equals_unsharded(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 400, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res), res
    assert all(nd for _, _, nd, _ in res), "per-shard thresholding should NOT reproduce the reference"
    assert res[0][3] > 0


def test_sample_per_gpu_assignment_and_split():
    from bronko_b200.dist import shard_samples, split_reads
    assert sorted(sum((shard_samples(100, r, 8) for r in range(8)), [])) == list(range(100))
    off = np.array([0, 3, 3, 10, 14], dtype=np.uint32)
    bases = np.arange(14, dtype=np.uint8)
    got = [split_reads(bases, off, r, 3) for r in range(3)]
    assert sum(len(o) - 1 for _, o in got) == 4
    assert np.concatenate([b for b, _ in got]).tolist() == bases.tolist()
    assert all(o[0] == 0 and o[-1] == len(b) for b, o in got)
