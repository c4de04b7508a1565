"""Shared helpers of the parity tests: run one sample through the oracle and through the CUDA path
and compare every observable the reference produces."""
import numpy as np

from oracle import oracle as O


def oracle_params(args):
    return O.Params.defaults(k=args.kmer, min_kmers=args.min_kmers, use_full_kmer=int(args.use_full_kmer),
                             n_fixed=args.n_fixed, min_af=args.min_af, no_end_filter=int(args.no_end_filter),
                             no_strand_filter=int(args.no_strand_filter),
                             no_strand_balance_filter=int(args.no_strand_balance_filter),
                             strand_balance_ratio=args.strand_balance_ratio, n_per_strand=args.n_per_strand,
                             strand_odds_max=args.strand_odds_max, min_depth=args.min_depth,
                             min_variant_depth=args.min_variant_depth, variant_multiplier=args.variant_multiplier)


def oracle_sample(oindex, files, args, threads=4):
    counts = [O.Counts.count(args.kmer, b, np.asarray(off, dtype=np.uint64), args.min_kmers, 1000000, threads)
              for b, off in files]
    return counts, O.Sample(oindex, oracle_params(args), counts)


SOR_TOL = 1e-9   # CUDA log() vs glibc log(): <= 2 ulp on terms of magnitude <= ~30
AF_TOL = 1e-6    # BASELINE.json north_star: "floating-point allele frequencies compared within 1e-6"


def assert_sample_equal(gpu, counts, osample, check_pileup=True):
    """gpu: bronko_b200.Sample; counts/osample: oracle Counts list and Sample."""
    for f, oc in enumerate(counts):
        ok, ov = oc.get()
        gk, gv = gpu.kmers(f)
        assert gpu.kmc_stats(f) == oc.stats(), "KMC stats of file %d" % f
        assert len(gk) == len(ok) and (gk == ok).all(), "k-mer set of file %d" % f
        assert (gv.astype(np.uint64) == ov).all(), "k-mer counts of file %d" % f
        os_ = osample.stats(f)
        gs = gpu.mapping_data(f)
        assert (gs["perfect"] == os_[:, 0]).all() and (gs["variant"] == os_[:, 1]).all(), "perfect/variant tallies"
        assert (gs["unique_perfect"] == os_[:, 2]).all(), "unique-perfect tallies"
        assert (gs["present"].astype(bool) == os_[:, 3].astype(bool)).all()
    assert gpu.best_genome == osample.best
    if osample.best < 0:
        return
    if check_pileup:
        gp, op = gpu.pileup(), osample.pileup()
        for a, name in enumerate(("fwd depth", "rev depth", "fwd support", "rev support")):
            assert (gp[a] == op[a]).all(), name
    assert np.array_equal(gpu.noise_max(), osample.noise_max()), "Noise.max"
    gv, ovr = gpu.variants, osample.variants()
    assert len(gv) == len(ovr), "number of variant records"
    for fld in ("seq", "pos", "ref_base", "alt_base", "fwd_ref", "rev_ref", "fwd_alt", "rev_alt", "depth"):
        assert (gv[fld] == ovr[fld]).all(), fld
    assert np.allclose(gv["af"], ovr["af"], rtol=0, atol=AF_TOL)
    assert (gv["af"] == ovr["af"]).all(), "AF is one IEEE division: expected bit-equal"
    assert np.allclose(gv["sor"], ovr["sor"], rtol=0, atol=SOR_TOL)
    major, minor, breadth, depth = osample.summary()
    assert (gpu.num_major_variants, gpu.num_minor_variants) == (major, minor)
    assert gpu.breadth_coverage == breadth
    assert gpu.depth_coverage == depth or (np.isnan(depth) and np.isnan(gpu.depth_coverage))
    assert gpu.num_unmapped_kmers == osample.unmapped()


def reads_from_strings(seqs):
    bases = np.frombuffer("".join(seqs).encode(), dtype=np.uint8).copy()
    off = np.zeros(len(seqs) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(s) for s in seqs])
    return bases, off
