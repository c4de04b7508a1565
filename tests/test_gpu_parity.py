"""Parity of the CUDA path (through the C ABI) against the oracle on the same seeded inputs.
Bit-exact for k-mer counts, stats, pileups, selection, Noise.max and the integer fields of every
variant; AF within 1e-6 (observed: bit-equal); SOR within 1e-9 (CUDA vs glibc log)."""
import numpy as np
import pytest

from bronko_b200 import sim

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import bronko_b200
    c = bronko_b200.Bronko(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def sars(ctx_sars):
    return ctx_sars


@pytest.fixture(scope="module")
def ctx_sars(sars_paths, oracle):
    import bronko_b200
    c = bronko_b200.Bronko(0)
    c.build_index(21, sars_paths)
    oi = oracle.Index.build(21, sars_paths)
    yield c, oi
    c.close()


@pytest.fixture(scope="module")
def ctx_hpv(hpv_bkdb_path, oracle):
    import bronko_b200
    c = bronko_b200.Bronko(0)
    c.load_index(hpv_bkdb_path)
    oi = oracle.Index.load(hpv_bkdb_path)
    yield c, oi
    c.close()


def run_both(c, oi, files, args=None, **kw):
    import bronko_b200
    from util import assert_sample_equal, oracle_sample
    args = args or bronko_b200.CallArgs()
    g = c.call_sample(files, args)
    counts, osample = oracle_sample(oi, files, args)
    assert_sample_equal(g, counts, osample, **kw)
    return g, osample


def test_index_builder_matches_oracle(ctx_sars):
    c, oi = ctx_sars
    k1, o1, e1 = oi.export()
    k2, o2, e2 = c.index_export()
    assert (k1 == k2).all() and (o1 == o2).all() and e1.tobytes() == e2.tobytes()


def test_hpv16_paired_bundled_db(ctx_hpv):
    """BASELINE config C1 (reads simulated from HPV16.fa; the bundled rep1_R*.fastq.gz are missing)."""
    c, oi = ctx_hpv
    r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.HPV16), 2000, sim.SEED0)
    g, o = run_both(c, oi, [(r1, o1), (r2, o2)])
    assert g.best_genome == 0 and len(g.variants) > 0


def test_hpv16_single_end(ctx_hpv):
    c, oi = ctx_hpv
    r1, o1, _, _, _ = sim.simulate_pairs(sim.load_genome(sim.HPV16), 800, sim.SEED0 + 1)
    run_both(c, oi, [(r1, o1)])


@pytest.mark.parametrize("strain", [0, 1, 2, 3])
def test_sars_four_strain_selection(ctx_sars, strain):
    """BASELINE config C4 shape: sample from strain s vs the 4-strain db, per-sample selection."""
    c, oi = ctx_sars
    r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[strain]), 600, sim.SEED0 + strain)
    g, o = run_both(c, oi, [(r1, o1), (r2, o2)])
    assert g.best_genome == strain


def test_sars_deeper_sample(ctx_sars):
    c, oi = ctx_sars
    r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[0]), 3000, sim.SEED0 + 7)
    g, o = run_both(c, oi, [(r1, o1), (r2, o2)])
    assert g.num_minor_variants > 0 and g.num_major_variants > 0


@pytest.mark.parametrize("depth", [400, 1500, 2000, 2500, 5000])
def test_sars_depth_ladder_noise_regimes(ctx_sars, depth):
    """The regimes of the noise chains (bk_noise.cuh: nz_chain_block): 90 active iterations of 29,953 at 400x (all walked in
    real FP64), a few thousand at 1,500-2,500x (serial runs and rounds mixed, the look-ahead choosing), two thirds at 5,000x
    (mostly rounds).  Noise.max and everything behind it bit for bit."""
    c, oi = ctx_sars
    r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[0]), depth, sim.SEED0 + depth)
    run_both(c, oi, [(r1, o1), (r2, o2)])


def test_stage_timing_off_changes_nothing(ctx_sars):
    c, oi = ctx_sars
    r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[1]), 300, sim.SEED0 + 31)
    c.set_stage_timing(False)
    try:
        g, o = run_both(c, oi, [(r1, o1), (r2, o2)])
        t = c.stage_times()
        assert t["scan_ms"] == 0 and t["score_ms"] == 0 and t["total_ms"] > 0 and t["launches"] > 0
    finally:
        c.set_stage_timing(True)
    c.call_sample([(r1, o1), (r2, o2)])
    assert c.stage_times()["scan_ms"] > 0


def test_vcf_and_pileup_text_identical(ctx_sars, tmp_path):
    c, oi = ctx_sars
    r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[2]), 500, sim.SEED0 + 11)
    g, o = run_both(c, oi, [(r1, o1), (r2, o2)])
    g.write_vcf("reads/sample_R1.fastq.gz", str(tmp_path / "s.vcf"))
    g.write_pileup(str(tmp_path / "s.tsv"))
    import re
    strip = lambda t: re.sub(r"SOR=[-0-9.a-zA-Z]+", "SOR=x", t)
    mine, ref = (tmp_path / "s.vcf").read_text(), o.vcf_text("reads/sample_R1.fastq.gz")
    assert strip(mine) == strip(ref)
    assert mine == ref        # SOR printed with 3 decimals: identical in practice
    assert (tmp_path / "s.tsv").read_text() == o.pileup_text()


@pytest.mark.parametrize("kw", [dict(use_full_kmer=True), dict(n_fixed=0), dict(n_fixed=5), dict(n_fixed=9),
                                dict(min_kmers=1), dict(min_kmers=10), dict(no_end_filter=True),
                                dict(no_strand_filter=True), dict(no_strand_balance_filter=True, strand_balance_ratio=0.4),
                                dict(min_af=0.2, min_depth=10, min_variant_depth=1), dict(n_per_strand=0),
                                dict(variant_multiplier=1.0, strand_odds_max=2.0)])
def test_flags(ctx_sars, kw):
    import bronko_b200
    c, oi = ctx_sars
    r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[1]), 300, sim.SEED0 + 21)
    run_both(c, oi, [(r1, o1), (r2, o2)], bronko_b200.CallArgs(**kw))


def test_n_fixed_too_large_queries_no_bucket(ctx_sars, oracle):
    """2*n_fixed+1 >= k → empty bucket slice (src/call.rs:1294-1296) → no genome → the reference exits 1."""
    import bronko_b200
    from util import oracle_sample
    c, oi = ctx_sars
    r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[1]), 100, sim.SEED0 + 22)
    args = bronko_b200.CallArgs(n_fixed=10)
    _, osample = oracle_sample(oi, [(r1, o1), (r2, o2)], args)
    assert osample.best == -1
    with pytest.raises(bronko_b200.BkError) as e:
        c.call_sample([(r1, o1), (r2, o2)], args)
    assert e.value.code == -4


@pytest.mark.parametrize("depth,kw", [(12, dict(min_kmers=1, min_depth=1, min_variant_depth=1)), (40, dict(min_kmers=2, min_depth=1)),
                                      (150, dict())])
def test_sparse_coverage_noise_chain(ctx_hpv, depth, kw):
    """Thin, gappy coverage: the windowed sums keep returning to (near) zero and cross binades all the time —
    the worst case for the exact parallel replication of the reference's sequential FP64 sums."""
    import bronko_b200
    c, oi = ctx_hpv
    r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.HPV16), depth, sim.SEED0 + 71 + depth)
    run_both(c, oi, [(r1, o1), (r2, o2)], bronko_b200.CallArgs(**kw))


def test_many_genome_map_kernel_on_small_db(sars_paths, oracle, monkeypatch):
    """The warp-per-k-mer map kernel (used for > 4 genomes) must agree with the thread-per-k-mer one."""
    import bronko_b200
    monkeypatch.setenv("BK_FORCE_WARP_MAP", "1")
    c = bronko_b200.Bronko(0)
    try:
        c.build_index(21, sars_paths)
        oi = oracle.Index.build(21, sars_paths)
        r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[3]), 500, sim.SEED0 + 41)
        run_both(c, oi, [(r1, o1), (r2, o2)])
    finally:
        c.close()


@pytest.mark.parametrize("warp_map", [False, True])
def test_id_keyed_map_tables(sars_paths, oracle, monkeypatch, warp_map):
    """BK_NO_REKEY keeps the bucket table keyed by the reference's bucket ids (the path used for k = 31 and for
    indexes whose keys do not verify): both map kernels must agree with the re-keyed default."""
    import bronko_b200
    monkeypatch.setenv("BK_NO_REKEY", "1")
    if warp_map:
        monkeypatch.setenv("BK_FORCE_WARP_MAP", "1")
    c = bronko_b200.Bronko(0)
    try:
        c.build_index(21, sars_paths)
        oi = oracle.Index.build(21, sars_paths)
        r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[2]), 400, sim.SEED0 + 43)
        run_both(c, oi, [(r1, o1), (r2, o2)])
    finally:
        c.close()


def _segmented_db(tmp_path, n_genomes=7, seed=5):
    """Synthetic db of n_genomes related 'segmented' genomes (1-3 contigs of 1.2-3 kb each, 2 % apart), with
    lower-case stretches and a few non-ACGT reference bases (encoded as A: Q11).  Returns (paths, contigs per genome)."""
    rng = np.random.default_rng(seed)
    base = [rng.integers(0, 4, size=int(n)).astype(np.uint8) for n in (3000, 1800, 1200)]
    paths, genomes = [], []
    for g in range(n_genomes):
        contigs = [sim.mutate_genome(b, 0.02, 1000 + 17 * g + i) for i, b in enumerate(base[:1 + g % 3])]
        p = tmp_path / ("strain%d.fa" % g)
        with open(p, "w") as f:
            for i, cod in enumerate(contigs):
                seq = "".join("ACGT"[x] for x in cod)
                if g % 2:
                    seq = seq[:200] + seq[200:260].lower() + seq[260:]          # soft-masked stretch
                if g == 3 and i == 0:
                    seq = seq[:500] + "N" + seq[501:900] + "R" + seq[901:]      # non-ACGT reference bases
                f.write(">seg%d strain%d\n" % (i + 1, g))
                for j in range(0, len(seq), 70):
                    f.write(seq[j:j + 70] + "\n")
        paths.append(str(p))
        genomes.append(contigs)
    return paths, genomes


@pytest.mark.parametrize("source", [2, 3, 5])
def test_many_genomes_multi_contig_db(tmp_path, oracle, source):
    """More than four genomes (warp-per-k-mer map, two passes) with several contigs each: per-contig noise,
    contig lookup in the variant kernel, selection among close strains, REF of non-ACGT bases."""
    import bronko_b200
    paths, genomes = _segmented_db(tmp_path)
    c = bronko_b200.Bronko(0)
    try:
        c.build_index(21, paths)
        oi = oracle.Index.build(21, paths)
        parts = []
        for i, cod in enumerate(genomes[source]):
            ascii_ = np.frombuffer("".join("ACGT"[x] for x in cod).encode(), dtype=np.uint8)
            parts.append(sim.simulate_pairs(ascii_, 400, sim.SEED0 + 70 + 3 * source + i, n_snv=3, n_isnv=5))
        def cat(idx_b, idx_o):
            bases = np.concatenate([p[idx_b] for p in parts])
            offs, shift = [np.zeros(1, np.uint32)], 0
            for p in parts:
                offs.append(p[idx_o][1:] + np.uint32(shift))
                shift += int(p[idx_o][-1])
            return bases, np.concatenate(offs)
        files = [cat(0, 1), cat(2, 3)]
        g, osample = run_both(c, oi, files, bronko_b200.CallArgs(min_depth=50))
        assert g.best_genome == source
    finally:
        c.close()


def test_shared_index_between_contexts(ctx_sars):
    """bk_index_share: a second context on the same GPU reads the first one's tables; results identical, and the
    index outlives the context that loaded it."""
    import bronko_b200
    c0, oi = ctx_sars
    c1 = bronko_b200.Bronko(0)
    try:
        c1.share_index(c0)
        r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[3]), 300, sim.SEED0 + 45)
        run_both(c1, oi, [(r1, o1), (r2, o2)])
        c2 = bronko_b200.Bronko(0)
        c2.build_index(21, [sim.genome_path(sim.HPV16)])
        c3 = bronko_b200.Bronko(0)
        c3.share_index(c2)
        c2.close()                                   # c3 keeps the HPV16 index alive
        import oracle.oracle as O
        h1, ho1, h2, ho2, _ = sim.simulate_pairs(sim.load_genome(sim.HPV16), 200, sim.SEED0 + 46)
        run_both(c3, O.Index.build(21, [sim.genome_path(sim.HPV16)]), [(h1, ho1), (h2, ho2)])
        c3.close()
    finally:
        c1.close()


def test_chunked_pushes_with_buffer_reuse(ctx_sars):
    """A file pushed in chunks from ONE pair of pinned host buffers that is overwritten the moment bk_reads_push
    returns (the contract: "both buffers have been copied when the call returns"), with chunks of very different
    sizes so that the novel k-mer list of the file has to learn its real fill and grow between pushes."""
    import ctypes as C
    import bronko_b200
    from bronko_b200 import _lib as L
    c, oi = ctx_sars
    r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[1]), 900, sim.SEED0 + 47)
    lib = L.lib()
    cuts = [0, 40, 45, 3000, 3100, 20000, len(o1) - 1]             # reads per chunk: 40, 5, 2955, 100, 16900, rest
    cap_b = max(int(o1[cuts[i + 1]]) - int(o1[cuts[i]]) for i in range(len(cuts) - 1)) + 64
    cap_o = max(cuts[i + 1] - cuts[i] for i in range(len(cuts) - 1)) + 1
    hb_p, ho_p = lib.bk_host_alloc(cap_b), lib.bk_host_alloc(cap_o * 4)
    assert hb_p and ho_p
    hb = np.ctypeslib.as_array(C.cast(hb_p, C.POINTER(C.c_uint8)), shape=(cap_b,))
    ho = np.ctypeslib.as_array(C.cast(ho_p, C.POINTER(C.c_uint32)), shape=(cap_o,))
    try:
        c.begin(bronko_b200.CallArgs())
        for slot, (b, o) in enumerate(((r1, o1), (r2, o2))):
            for i in range(len(cuts) - 1):
                lo, hi = cuts[i], cuts[i + 1]
                b0, b1 = int(o[lo]), int(o[hi])
                hb[:b1 - b0] = b[b0:b1]
                hb[b1 - b0:b1 - b0 + 64] = ord("*")
                ho[:hi - lo + 1] = (o[lo:hi + 1].astype(np.int64) - b0).astype(np.uint32)
                c.push_ptr(slot, hb_p, ho_p, hi - lo)
                hb[:] = ord("N")                                   # the caller reuses its buffers at once
                ho[:] = 0
        g = c.finish()
        from util import assert_sample_equal, oracle_sample
        counts, osample = oracle_sample(oi, [(r1, o1), (r2, o2)], bronko_b200.CallArgs())
        assert_sample_equal(g, counts, osample)
    finally:
        lib.bk_host_free(hb_p); lib.bk_host_free(ho_p)


def test_deep_duplication_single_round_bins(ctx_hpv):
    """The same few thousand novel k-mers repeated until every bin holds far more occurrences than a worst-case round
    (what a 10^6x sample looks like): bins are counted in ONE optimistic round; and a list with as many DISTINCT
    novel k-mers, where the optimistic round must give up and fall back to worst-case rounds."""
    from util import reads_from_strings
    c, oi = ctx_hpv
    rng = np.random.default_rng(19)
    g = "".join(chr(x) for x in sim.load_genome(sim.HPV16))
    own = [g[i:i + 150] for i in range(0, 7000, 7)] * 4                     # (something to select a genome with)
    few = ["".join("ACGT"[i] for i in rng.integers(0, 4, size=150)) for _ in range(40)]
    run_both(c, oi, [reads_from_strings(few * 1500 + own)])                # 7.8 M occurrences of 5,200 k-mers
    many = ["".join("ACGT"[i] for i in rng.integers(0, 4, size=150)) for _ in range(30000)]
    run_both(c, oi, [reads_from_strings(many * 3 + own)])                  # 3.9 M distinct k-mers, three times each


def test_heavy_hitter_and_fixed_capacity(ctx_hpv):
    """One foreign read repeated 3,000 times (a novel k-mer with thousands of copies lands in one bin: equal keys are
    merged per warp, large bins are walked in rounds) and a caller-fixed list capacity (table_log2)."""
    import bronko_b200
    from util import reads_from_strings
    c, oi = ctx_hpv
    rng = np.random.default_rng(9)
    g = "".join(chr(x) for x in sim.load_genome(sim.HPV16))
    foreign = "".join("ACGT"[i] for i in rng.integers(0, 4, size=150))
    seqs = [foreign] * 3000 + [g[i:i + 150] for i in range(0, 7000, 7)] * 4
    run_both(c, oi, [reads_from_strings(seqs)])
    run_both(c, oi, [reads_from_strings(seqs)], bronko_b200.CallArgs(table_log2=20))
    with pytest.raises(bronko_b200.BkError) as e:           # 2^12 entries cannot hold 390,000 novel occurrences
        c.call_sample([reads_from_strings(seqs)], bronko_b200.CallArgs(table_log2=12))
    assert e.value.code == -5


@pytest.mark.parametrize("two_pass", [False, True])
def test_per_bucket_table_map(sars_paths, oracle, monkeypatch, two_pass):
    """BK_NO_GROUP_MAP: 16 independent probes of the per-bucket table instead of two group probes."""
    import bronko_b200
    monkeypatch.setenv("BK_NO_GROUP_MAP", "1")
    if two_pass:
        monkeypatch.setenv("BK_NO_FUSED_MAP", "1")
    c = bronko_b200.Bronko(0)
    try:
        c.build_index(21, sars_paths)
        oi = oracle.Index.build(21, sars_paths)
        r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[3]), 400, sim.SEED0 + 48)
        run_both(c, oi, [(r1, o1), (r2, o2)])
    finally:
        c.close()


def test_two_pass_map_on_small_db(sars_paths, oracle, monkeypatch):
    """BK_NO_FUSED_MAP: tallies, selection, then a second pass for the selected genome's pileup (what databases of
    more than four genomes and the read-sharded mode use) instead of the one-pass map of small databases."""
    import bronko_b200
    monkeypatch.setenv("BK_NO_FUSED_MAP", "1")
    c = bronko_b200.Bronko(0)
    try:
        c.build_index(21, sars_paths)
        oi = oracle.Index.build(21, sars_paths)
        r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[1]), 400, sim.SEED0 + 44)
        run_both(c, oi, [(r1, o1), (r2, o2)])
    finally:
        c.close()


@pytest.mark.parametrize("k", [15, 19, 29, 31])
def test_other_k(oracle, k):
    """k = 15 .. 31 (odd): k <= 29 probes the re-keyed table, k = 31 the id-keyed one (bucket ids wrap there, Q20)."""
    import bronko_b200
    paths = [sim.genome_path(sim.HPV16)]
    c = bronko_b200.Bronko(0)
    try:
        c.build_index(k, paths)
        oi = oracle.Index.build(k, paths)
        r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.HPV16), 300, sim.SEED0 + 50 + k)
        run_both(c, oi, [(r1, o1), (r2, o2)], bronko_b200.CallArgs(kmer=k))
    finally:
        c.close()


def test_edge_reads(ctx_hpv):
    """Ragged input: empty reads, reads shorter than k, N / lower-case / junk bytes, a read longer than a
    tile, reads hanging over both genome ends, foreign reads, an indel."""
    from util import reads_from_strings
    c, oi = ctx_hpv
    g = sim.load_genome(sim.HPV16).tobytes().decode()
    rc = g[::-1].translate(str.maketrans("ACGT", "TGCA"))
    rng = np.random.default_rng(5)
    seqs = []
    for rep in range(4):          # >= min_kmers copies so things survive the -ci3 threshold
        seqs += ["", "A", g[100:120], g[100:121], g[200:350], g[200:350].lower(),
                 g[400:470] + "N" + g[471:560], g[600:640] + "n*" + g[642:700], "N" * 40,
                 "GATTACA" * 30, g[1000:1100] + g[1103:1250], g[1500:1560] + "ACGT" + g[1560:1700],
                 "TTTTTTTTTT" + g[0:140], g[-140:] + "GGGGGGGGGG", rc[0:150], "CCCCC" + rc[-100:] + "AAAAA",
                 g[3000:3000 + 9000 % 4100], g[2000:2300], g[2000:2021], g[2001:2022]]
        seqs.append("".join("ACGT"[i] for i in rng.integers(0, 4, size=300)))
    files = [reads_from_strings(seqs)]
    import bronko_b200
    run_both(c, oi, files, bronko_b200.CallArgs(min_depth=1))


def test_long_reads_beyond_tile(ctx_hpv):
    from util import reads_from_strings
    c, oi = ctx_hpv
    g = sim.load_genome(sim.HPV16).tobytes().decode()
    seqs = [g, g[10:7000], g[500:], g[:3000] + "N" + g[3001:]] * 3
    run_both(c, oi, [reads_from_strings(seqs)])


def test_no_reads_matching_gives_no_genome(ctx_hpv):
    import bronko_b200
    from util import reads_from_strings
    c, oi = ctx_hpv
    rng = np.random.default_rng(9)
    seqs = ["".join("ACGT"[i] for i in rng.integers(0, 4, size=150)) for _ in range(50)] * 3
    with pytest.raises(bronko_b200.BkError) as e:
        c.call_sample([reads_from_strings(seqs)])
    assert e.value.code == -4 and "Unable to pick a best genome" in str(e.value)


def test_counter_saturation_and_thresholds(ctx_hpv):
    """-cs saturation (lowered so a small input reaches it) and R1/R2 thresholded separately (Q1, Q2)."""
    from util import reads_from_strings
    import bronko_b200
    c, oi = ctx_hpv
    g = sim.load_genome(sim.HPV16).tobytes().decode()
    r1 = reads_from_strings([g[100:250]] * 2 + [g[300:450]] * 5)
    r2 = reads_from_strings([g[100:250]] * 2 + [g[500:650]] * 4)
    gs, o = run_both(c, oi, [r1, r2], bronko_b200.CallArgs(min_depth=1))
    km, _ = gs.kmers(0)
    # the k-mers of g[100:250] occur twice in each file: dropped from both (never merged across files)
    from oracle import oracle as O
    first = O.canonical_kmer(g[100:121])[0]
    assert first not in set(km.tolist())


def test_device_resident_push_matches_host_push(ctx_sars):
    import torch
    c, oi = ctx_sars
    r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[0]), 400, sim.SEED0 + 31)
    host = c.call_sample([(r1, o1), (r2, o2)])
    hv, hp = host.variants.copy(), host.pileup()
    dev = []
    for b, o in ((r1, o1), (r2, o2)):
        pad = np.concatenate([b, np.full(64, ord("*"), dtype=np.uint8)])
        dev.append((torch.from_numpy(pad).cuda(), torch.from_numpy(o.astype(np.int64)).to(torch.int32).cuda(), len(o) - 1, len(b)))
    torch.cuda.synchronize()
    c.begin()
    for slot, (tb, to, n, nb) in enumerate(dev):
        c.push_device(slot, tb.data_ptr(), to.data_ptr(), n, nb, 150)
    s = c.finish()
    assert s.variants.tobytes() == hv.tobytes() and (s.pileup() == hp).all()


def test_decoded_fastq_push_matches_host_push(ctx_sars, tmp_path):
    """bk_fastq_decode + bk_reads_push_decoded (and bk_reads_push_fastq on top of them) against pushing the same
    reads from numpy buffers: gz and plain files, one of them with CRLF line ends."""
    import bronko_b200
    from test_gpu_fullsize import assert_snapshots_equal, snapshot
    c, _ = ctx_sars
    r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[3]), 300, sim.SEED0 + 91)
    f1, f2 = str(tmp_path / "d_R1.fastq.gz"), str(tmp_path / "d_R2.fastq")
    sim.write_fastq(f1, r1, o1, "d", 1)
    sim.write_fastq(f2, r2, o2, "d", 2)
    with open(f2, "rb") as f:
        crlf = f.read().replace(b"\n", b"\r\n")
    with open(f2, "wb") as f:
        f.write(crlf)
    base = snapshot(c.call_sample([(r1, o1), (r2, o2)]), 2)
    d1, d2 = bronko_b200.DecodedReads(f1), bronko_b200.DecodedReads(f2)
    c.begin()
    c.push_decoded(0, d1)
    c.push_decoded(1, d2)
    d1.close(); d2.close()
    assert_snapshots_equal(base, snapshot(c.finish(), 2))
    c.begin()
    c.push_fastq(0, f1)
    c.push_fastq(1, f2)
    assert_snapshots_equal(base, snapshot(c.finish(), 2))


def test_packed_push_equals_ascii_push(ctx_sars):
    """bk_reads_push_packed + the ASCII push of the reads that cannot be packed = the ASCII push of everything: a sample
    with N / junk / lower-case reads mixed in, R1 packed in two chunks, R2 as ASCII."""
    import bronko_b200
    from bronko_b200 import _lib as L
    from util import assert_sample_equal, oracle_sample
    c, oi = ctx_sars
    r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[2]), 500, sim.SEED0 + 49)
    r1 = r1.copy()
    rng = np.random.default_rng(3)
    for r in rng.integers(0, len(o1) - 1, size=300):                 # N's, junk and lower case into ~300 reads
        p = int(o1[r]) + int(rng.integers(0, 150))
        r1[p] = [ord("N"), ord("*"), r1[p] | 0x20][int(rng.integers(0, 3))]
    args = bronko_b200.CallArgs()
    counts, osample = oracle_sample(oi, [(r1, o1), (r2, o2)], args)
    half = (len(o1) - 1) // 2
    c.begin(args)
    for lo, hi in ((0, half), (half, len(o1) - 1)):
        b0, b1 = int(o1[lo]), int(o1[hi])
        packed, poff, rest, roff = bronko_b200.pack_reads(r1[b0:b1], (o1[lo:hi + 1].astype(np.int64) - b0).astype(np.uint32))
        assert len(roff) - 1 > 0 and len(poff) - 1 > 0
        c.push_packed_ptr(0, L.ptr(packed), L.ptr(poff), len(poff) - 1)
        c.push(0, rest, roff)
    c.push(1, r2, o2)
    assert_sample_equal(c.finish(), counts, osample)
