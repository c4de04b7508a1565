"""Parity at the FULL sizes of the BASELINE configs, through the C ABI.

The oracle finishes config C2 (2.99e8 bases) in seconds on a multi-core host, so the full-size case is compared
bit for bit like the small ones; on top of that the size-independent properties the domain offers are checked on
the CUDA path alone: the KMC totals in closed form, a checksum of the counts, the planted truth, invariance under
the order of the reads and of the two files, and additivity of the counts over a split of the reads (the property
the read-sharded mode relies on)."""
import os

import numpy as np
import pytest

from bronko_b200 import sim

pytestmark = pytest.mark.gpu

K = 21
THREADS = max(4, min(32, os.cpu_count() or 4))


@pytest.fixture(scope="module")
def ctx_sars(sars_paths, oracle):
    import bronko_b200
    c = bronko_b200.Bronko(0)
    c.build_index(K, sars_paths)
    oi = oracle.Index.build(K, sars_paths)
    yield c, oi
    c.close()


@pytest.fixture(scope="module")
def c2_reads():
    """BASELINE config C2: SARS-CoV-2 (wuhan_ref) 10,000x, 150 bp PE, the bench workload of rank 0."""
    r1, o1, r2, o2, truth = sim.simulate_pairs(sim.load_genome(sim.SARS4[0]), 10000, sim.SEED0)
    return [(r1, o1), (r2, o2)], truth


def snapshot(g, n_files):
    """Everything observable of one finished sample, as comparable values."""
    return {"kmers": [tuple(a.copy() for a in g.kmers(f)) for f in range(n_files)],
            "kmc": [g.kmc_stats(f) for f in range(n_files)],
            "stats": [g.mapping_data(f).copy() for f in range(n_files)],
            "best": g.best_genome, "pileup": g.pileup(), "noise": g.noise_max(), "variants": g.variants.copy(),
            "summary": (g.num_major_variants, g.num_minor_variants, g.breadth_coverage, g.depth_coverage,
                        g.num_unmapped_kmers)}


def assert_snapshots_equal(a, b, files_swapped=False):
    nf = len(a["kmers"])
    for f in range(nf):
        fb = nf - 1 - f if files_swapped else f
        assert a["kmc"][f] == b["kmc"][fb]
        assert np.array_equal(a["kmers"][f][0], b["kmers"][fb][0]) and np.array_equal(a["kmers"][f][1], b["kmers"][fb][1])
        assert a["stats"][f].tobytes() == b["stats"][fb].tobytes()
    assert a["best"] == b["best"]
    assert np.array_equal(a["pileup"], b["pileup"])
    assert np.array_equal(a["noise"], b["noise"])
    assert a["variants"].tobytes() == b["variants"].tobytes()
    assert a["summary"] == b["summary"]


def test_c2_full_size_bit_exact(ctx_sars, c2_reads):
    """Config C2 at its full depth against the oracle: k-mer sets, counts, KMC numbers, tallies, selection, the four
    pileup arrays, Noise.max and every variant record; then the closed-form totals and the planted truth."""
    import bronko_b200
    from util import assert_sample_equal, oracle_sample
    c, oi = ctx_sars
    files, truth = c2_reads
    args = bronko_b200.CallArgs()
    g = c.call_sample(files, args)
    counts, osample = oracle_sample(oi, files, args, threads=THREADS)
    assert_sample_equal(g, counts, osample)
    for f, (b, off) in enumerate(files):
        n_reads = len(off) - 1
        total_reads, total_kmers, unique, unique_counted = g.kmc_stats(f)
        assert total_reads == n_reads == 996767                      # 996,767 pairs
        assert total_kmers == n_reads * (150 - K + 1)                # no N, every read 150 bp
        km, ct = g.kmers(f)
        assert unique_counted == len(km) <= unique
        assert (np.diff(km.astype(np.int64)) > 0).all()              # sorted, distinct (k = 21: values < 2^42)
        assert ct.min() >= args.min_kmers and ct.max() <= 1000000
    # planted truth: 10 SNVs + 20 iSNVs, all found, nothing else called, AF near the planted frequency
    v = g.variants
    assert g.best_genome == 0 and len(v) == 30
    got = {(int(p), int(a)): float(af) for p, a, af in zip(v["pos"], v["alt_base"], v["af"])}
    for p, a, af in zip(truth["pos"], truth["alt"], truth["af"]):
        assert (int(p) + 1, int(a)) in got                           # VCF positions are 1-based
        assert abs(got[(int(p) + 1, int(a))] - af) < 0.02
    assert (g.num_major_variants, g.num_minor_variants) == (10, 20)


def test_c2_full_size_order_and_file_invariance(ctx_sars, c2_reads):
    """Integer pileups do not depend on the order in which reads arrive (atomics commute) nor on which file is R1:
    depth is a max and support a sum over both files (src/call.rs:316-317, 1337-1345)."""
    import bronko_b200
    c, _ = ctx_sars
    files, _ = c2_reads
    base = snapshot(c.call_sample(files), 2)
    rng = np.random.default_rng(7)
    shuffled = []
    for b, off in files:
        perm = rng.permutation(len(off) - 1)
        shuffled.append((b.reshape(-1, 150)[perm].reshape(-1).copy(), off))
    assert_snapshots_equal(base, snapshot(c.call_sample(shuffled), 2))
    assert_snapshots_equal(base, snapshot(c.call_sample([files[1], files[0]]), 2), files_swapped=True)
    assert_snapshots_equal(base, snapshot(c.call_sample(files), 2))          # and the context is reusable: idempotent


def test_c2_full_size_counts_are_additive(ctx_sars, c2_reads):
    """Linearity of the counting stage at full size: with every k-mer kept (min_kmers = 1) the counts of the whole
    file are the sum of the counts of its two halves, and they add up to the number of k-mer occurrences."""
    import bronko_b200
    c, _ = ctx_sars
    b, off = c2_reads[0][0]
    n = len(off) - 1
    args = bronko_b200.CallArgs(min_kmers=1)
    whole = c.call_sample([(b, off)], args)
    wk, wc = (a.copy() for a in whole.kmers(0))
    assert int(wc.astype(np.uint64).sum()) == n * (150 - K + 1) == whole.kmc_stats(0)[1]
    assert whole.kmc_stats(0)[2] == whole.kmc_stats(0)[3] == len(wk)
    h = n // 2
    parts = []
    for lo, hi in ((0, h), (h, n)):
        s = c.call_sample([(b[lo * 150:hi * 150], (off[lo:hi + 1] - off[lo]).astype(np.uint32))], args)
        parts.append(tuple(a.copy() for a in s.kmers(0)))
    keys = np.concatenate([parts[0][0], parts[1][0]])
    vals = np.concatenate([parts[0][1], parts[1][1]]).astype(np.uint64)
    uk, inv = np.unique(keys, return_inverse=True)
    summed = np.zeros(len(uk), dtype=np.uint64)
    np.add.at(summed, inv, vals)
    assert np.array_equal(uk, wk) and np.array_equal(summed, wc.astype(np.uint64))


def test_c4_full_depth_other_strain(ctx_sars):
    """Config C4 shape at its full depth: a sample of strain 2 against the 4-strain db, selection on device."""
    import bronko_b200
    from util import assert_sample_equal, oracle_sample
    c, oi = ctx_sars
    r1, o1, r2, o2, _ = sim.config_reads("C4", sample=2)
    files = [(r1, o1), (r2, o2)]
    g = c.call_sample(files)
    counts, osample = oracle_sample(oi, files, bronko_b200.CallArgs(), threads=THREADS)
    assert_sample_equal(g, counts, osample)
    assert g.best_genome == 2


@pytest.mark.parametrize("n_strains,source", [(48, 17), (200, 117)])
def test_c5_many_strain_db(tmp_path, oracle, n_strains, source):
    """Config C5: synthetic strains (wuhan_ref with i.i.d. 1 % substitutions, seed as in SURVEY.md 8d) and a 2,000x
    sample of one of them — at 48 strains (6.2 M keys / 30 M entries) and at the config's own 200 strains (18.7 M keys /
    125 M entries, ~150 entries per bucket: the large-table map kernel with its warp-flattened entry walk).  The oracle
    needs about a minute to build the 200-strain index."""
    import bronko_b200
    from util import assert_sample_equal, oracle_sample
    g0 = sim._CODE[sim.load_genome(sim.SARS4[0])]
    paths, strains = [], []
    for s in range(n_strains):
        cod = sim.mutate_genome(g0, 0.01, sim.SEED0 + 10 ** 6 + s)
        p = str(tmp_path / ("strain%03d.fa" % s))
        sim.write_fasta(p, "strain%03d" % s, cod)
        paths.append(p)
        strains.append(cod)
    c = bronko_b200.Bronko(0)
    try:
        c.build_index(K, paths)
        oi = oracle.Index.build(K, paths)
        info = c.index_info()
        assert info["n_genomes"] == n_strains
        r1, o1, r2, o2, _ = sim.simulate_pairs(sim._ASCII[strains[source]], 2000, sim.SEED0 + source)
        files = [(r1, o1), (r2, o2)]
        g = c.call_sample(files)
        counts, osample = oracle_sample(oi, files, bronko_b200.CallArgs(), threads=THREADS)
        assert_sample_equal(g, counts, osample)
        assert g.best_genome == source
    finally:
        c.close()
