"""Read-sharded deep sample on the GPU (bk_shard_* + bronko_b200/dist.py): ranks scan disjoint shares of
the reads, counts are merged over NCCL (or, with a single GPU, over gloo with both ranks on device 0),
and every rank must end with exactly the unsharded result."""
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


def _worker(rank, world, port, n_gpus, q):
    try:
        sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
        import torch.distributed as dist
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dev = rank % n_gpus
        torch.cuda.set_device(dev)
        if n_gpus >= world:
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
        else:
            dist.init_process_group("gloo", rank=rank, world_size=world)
        import bronko_b200
        from bronko_b200 import sim
        from bronko_b200.dist import call_sample_sharded, split_reads
        from oracle import oracle as O
        from util import assert_sample_equal, oracle_sample
        paths = [sim.genome_path(n) for n in sim.SARS4]
        ctx = bronko_b200.Bronko(dev)
        ctx.build_index(21, paths)
        r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[2]), 700, sim.SEED0 + 61)
        args = bronko_b200.CallArgs()
        s = call_sample_sharded(ctx, [split_reads(r1, o1, rank, world), split_reads(r2, o2, rank, world)], args)
        counts, osample = oracle_sample(O.Index.build(21, paths), [(r1, o1), (r2, o2)], args)
        # the sharded run keeps no per-rank copy of the full k-mer list: compare everything downstream of it
        for f, oc in enumerate(counts):
            assert s.kmc_stats(f) == oc.stats(), (s.kmc_stats(f), oc.stats())
            os_, gs = osample.stats(f), s.mapping_data(f)
            assert (gs["perfect"] == os_[:, 0]).all() and (gs["variant"] == os_[:, 1]).all() and (gs["unique_perfect"] == os_[:, 2]).all()
        assert s.best_genome == osample.best
        assert (s.pileup() == osample.pileup()).all()
        assert np.array_equal(s.noise_max(), osample.noise_max())
        gv, ov = s.variants, osample.variants()
        assert len(gv) == len(ov)
        for fld in ("seq", "pos", "ref_base", "alt_base", "fwd_ref", "rev_ref", "fwd_alt", "rev_alt", "depth", "af"):
            assert (gv[fld] == ov[fld]).all(), fld
        assert np.allclose(gv["sor"], ov["sor"], rtol=0, atol=1e-9)
        assert s.num_unmapped_kmers == osample.unmapped()
        ctx.close()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:      # noqa: BLE001
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()[-1500:]))


@pytest.mark.parametrize("world", [2, 4])
def test_read_sharded_matches_oracle(world):
    n_gpus = torch.cuda.device_count()
    if n_gpus < 1:
        pytest.skip("no GPU")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_gpus, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    assert all(msg == "ok" for _, msg in res), res
