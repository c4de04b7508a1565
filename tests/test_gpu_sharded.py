"""Read-sharded deep sample on the GPU (BASELINE config C3; bk_shard_* in include/bronko_b200.h): ranks scan disjoint
shares of the reads, the library merges the counts BEFORE the cut-offs (src/call.rs:1172-1173, 1341-1345) and every rank
must end with exactly the unsharded result.

* in-process transport (bk_shard_local): all ranks are contexts on ONE device — runs on the single-GPU test box and
  exercises every kernel of the sharded path against the oracle;
* NCCL transport (bk_shard_init): one process per GPU; needs >= 2 GPUs, skipped otherwise (bench.py's sharded leg
  runs it under torchrun on the multi-GPU box and fails the bench if the result is not bit-equal)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _assert_downstream_equal(s, counts, osample):
    """A sharded run keeps no per-rank copy of the full k-mer list: compare everything downstream of it."""
    for f, oc in enumerate(counts):
        assert s.kmc_stats(f) == oc.stats(), (s.kmc_stats(f), oc.stats())
        os_, gs = osample.stats(f), s.mapping_data(f)
        assert (gs["perfect"] == os_[:, 0]).all() and (gs["variant"] == os_[:, 1]).all() and (gs["unique_perfect"] == os_[:, 2]).all()
        assert (gs["present"].astype(bool) == os_[:, 3].astype(bool)).all()
    assert s.best_genome == osample.best
    assert (s.pileup() == osample.pileup()).all()
    assert np.array_equal(s.noise_max(), osample.noise_max())
    gv, ov = s.variants, osample.variants()
    assert len(gv) == len(ov)
    for fld in ("seq", "pos", "ref_base", "alt_base", "fwd_ref", "rev_ref", "fwd_alt", "rev_alt", "depth", "af"):
        assert (gv[fld] == ov[fld]).all(), fld
    assert np.allclose(gv["sor"], ov["sor"], rtol=0, atol=1e-9)
    assert s.num_unmapped_kmers == osample.unmapped()
    major, minor, breadth, depth = osample.summary()
    assert (s.num_major_variants, s.num_minor_variants) == (major, minor)
    assert s.breadth_coverage == breadth and s.depth_coverage == depth


@pytest.mark.parametrize("world,two_pass", [(2, False), (3, False), (4, True)])
def test_local_group_matches_oracle(world, two_pass, sars_paths, oracle, monkeypatch):
    import bronko_b200
    from bronko_b200 import sim
    from bronko_b200.dist import ShardedLocal, split_reads
    from util import oracle_sample
    if two_pass:
        monkeypatch.setenv("BK_NO_FUSED_MAP", "1")
    ctx = bronko_b200.Bronko(0)
    try:
        ctx.build_index(21, sars_paths)
        oi = oracle.Index.build(21, sars_paths)
        r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[2]), 700, sim.SEED0 + 61)
        args = bronko_b200.CallArgs()
        counts, osample = oracle_sample(oi, [(r1, o1), (r2, o2)], args)
        sl = ShardedLocal(ctx, world)
        try:
            for rep in range(2):                                        # a second sample on the same group
                s0 = sl.call_sample([[split_reads(r1, o1, r, world), split_reads(r2, o2, r, world)] for r in range(world)], args)
                _assert_downstream_equal(s0, counts, osample)
                for r in range(1, world):                               # every rank holds the same result
                    sr = sl.sample_of(r)
                    assert sr.variants.tobytes() == s0.variants.tobytes()
                    assert (sr.pileup() == s0.pileup()).all()
                    assert [sr.kmc_stats(f) for f in range(2)] == [s0.kmc_stats(f) for f in range(2)]
            t = ctx.stage_times()
            assert t["coll_calls"] >= 8
        finally:
            sl.close()
        # the contexts are whole-sample contexts again
        g = ctx.call_sample([(r1, o1), (r2, o2)], args)
        _assert_downstream_equal(g, counts, osample)
    finally:
        ctx.close()


def test_local_group_uneven_and_empty_shares(sars_paths, oracle):
    """Shares of very different sizes (the bin counts differ per rank: owners are hash ranges, not bins), a rank without
    reads, single-end."""
    import bronko_b200
    from bronko_b200 import sim
    from bronko_b200.dist import ShardedLocal
    from util import oracle_sample
    ctx = bronko_b200.Bronko(0)
    try:
        ctx.build_index(21, sars_paths)
        oi = oracle.Index.build(21, sars_paths)
        r1, o1, _, _, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[0]), 500, sim.SEED0 + 62)
        n = len(o1) - 1
        cuts = [0, n // 50, n // 50, n // 3, n]                         # 2 %, nothing, 31 %, 67 %
        shares = []
        for r in range(4):
            lo, hi = cuts[r], cuts[r + 1]
            b0, b1 = int(o1[lo]), int(o1[hi])
            shares.append([(r1[b0:b1], (o1[lo:hi + 1].astype(np.int64) - b0).astype(np.uint32))])
        args = bronko_b200.CallArgs()
        counts, osample = oracle_sample(oi, [(r1, o1)], args)
        sl = ShardedLocal(ctx, 4)
        try:
            _assert_downstream_equal(sl.call_sample(shares, args), counts, osample)
        finally:
            sl.close()
    finally:
        ctx.close()


def _nccl_worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
        import torch.distributed as dist
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        import bronko_b200
        from bronko_b200 import sim
        from bronko_b200.dist import call_sample_sharded, init_sharded, leave_sharded, split_reads
        from oracle import oracle as O
        from util import oracle_sample
        paths = [sim.genome_path(n) for n in sim.SARS4]
        ctx = bronko_b200.Bronko(rank)
        ctx.build_index(21, paths)
        init_sharded(ctx)
        r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[2]), 700, sim.SEED0 + 61)
        args = bronko_b200.CallArgs()
        s = call_sample_sharded(ctx, [split_reads(r1, o1, rank, world), split_reads(r2, o2, rank, world)], args)
        counts, osample = oracle_sample(O.Index.build(21, paths), [(r1, o1), (r2, o2)], args)
        _assert_downstream_equal(s, counts, osample)
        leave_sharded(ctx)
        ctx.close()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception:      # noqa: BLE001
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()[-1500:]))


@pytest.mark.parametrize("world", [2, 4])
def test_nccl_group_matches_oracle(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (NCCL refuses two ranks on one device); the in-process tests cover the kernels" % world)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    assert all(msg == "ok" for _, msg in res), res
