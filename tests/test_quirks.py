"""Oracle tests for the bug-compatible quirks of SURVEY.md Appendix C: each fails if the quirk is 'fixed'."""
import numpy as np
import pytest

from bronko_b200 import sim
from util import reads_from_strings

K = 21


@pytest.fixture(scope="module")
def hpv(oracle, hpv_fasta):
    return oracle.Index.build(K, [hpv_fasta]), sim.load_genome(sim.HPV16).tobytes().decode()


def rc(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def enc(s):
    v = 0
    for c in s:
        v = (v << 2) | "ACGT".index(c)
    return v


def run(oracle, ix, kmers_counts, **kw):
    km = np.array([enc(s) for s, _ in kmers_counts], dtype=np.uint64)
    ct = np.array([c for _, c in kmers_counts], dtype=np.uint64)
    p = oracle.Params.defaults(k=K, **kw)
    return oracle.Sample(ix, p, [oracle.Counts.from_list(km, ct)])


def test_q2_noncanonical_counting_threshold_and_saturation(oracle):
    s = "ACGTTGCAAGGCTTAACCGTAGGCAT"
    reads = [s] * 5 + [rc(s)] * 2
    b, off = reads_from_strings(reads)
    c = oracle.Counts.count(K, b, off.astype(np.uint64), ci=3, cs=4, threads=2)
    km, ct = c.get()
    fwd = {enc(s[i:i + K]) for i in range(len(s) - K + 1)}
    rev = {enc(rc(s)[i:i + K]) for i in range(len(s) - K + 1)}
    assert set(km.tolist()) == fwd              # rc k-mers (count 2) are separate keys and fall below ci
    assert not (rev & set(km.tolist()))
    assert set(ct.tolist()) == {4}              # 5 occurrences saturate at cs=4
    tr, tk, uk, uc = c.stats()
    assert (tr, tk, uk, uc) == (7, 7 * 6, 12, 6)


def test_kmc_splits_reads_at_non_acgt_and_accepts_lowercase(oracle):
    s = "ACGTTGCAAGGCTTAACCGTAGGCATTTGACC"
    b, off = reads_from_strings([s[:25] + "N" + s[26:], s.lower(), s[:20], ""])
    c = oracle.Counts.count(K, b, off.astype(np.uint64), ci=1, cs=1000000, threads=1)
    km, ct = c.get()
    want = {}
    for i in range(len(s) - K + 1):
        want[enc(s[i:i + K])] = 1                       # lower-case read counts like upper case
    for i in range(25 - K + 1):
        want[enc(s[i:i + K])] += 1                      # left piece of the N-split read (25 bases)
    assert dict(zip(km.tolist(), ct.tolist())) == want  # right piece (6 bases) and the 20-base read: nothing
    assert c.stats()[0] == 4                            # reads shorter than k still count as reads


def test_q3_depth_is_max_not_sum_and_q4_support_counts_hits(oracle, hpv):
    ix, g = hpv
    a, b_ = g[1000:1021], g[1005:1026]
    s = run(oracle, ix, [(a, 7), (b_, 4)])
    assert s.best == 0
    p = s.pileup()
    # position 1010 is covered by both k-mers: depth = max(7, 4), support = 2 hits on one strand
    ref = "ACGT".index(g[1010])
    assert int(p[0][1010][ref] + p[1][1010][ref]) == 7
    assert int(p[2][1010][ref] + p[3][1010][ref]) == 2


def test_q5_asymmetric_bucket_slice(oracle, hpv):
    """buckets[n_fixed .. k-n_fixed-1): canonical indices 2..17 for k=21 — 16 buckets, 2 dropped left, 3 right."""
    ix, g = hpv
    kmer = g[2000:2021]
    s = run(oracle, ix, [(kmer, 5)])
    p = s.pileup()
    cov = np.nonzero((p[0] + p[1]).sum(axis=1))[0]
    canon_is_rc = enc(kmer) >= enc(rc(kmer))
    lo = 2000 + (3 if canon_is_rc else 2)       # canonical index i sits at forward offset i or k-1-i
    assert cov.tolist() == list(range(lo, lo + 16))
    s_full = run(oracle, ix, [(kmer, 5)], use_full_kmer=1)
    assert np.count_nonzero((s_full.pileup()[0] + s_full.pileup()[1]).sum(axis=1)) == 21
    s_none = run(oracle, ix, [(kmer, 5)], n_fixed=10)
    assert s_none.best == -1                    # 2*n_fixed+1 >= k → no bucket → no genome (reference exits 1)


def test_q6_rc_canonical_reference_kmers_record_alt_only_via_noncanonical_entries(oracle, hpv):
    """A 1-mismatch k-mer against a reference k-mer whose canonical form is the reverse complement deposits
    a REF base at the mirrored position; the ALT is only recorded for canonical=false entries."""
    ix, g = hpv
    found = {}
    for start in range(500, 3000):
        kmer = g[start:start + K]
        is_rc = enc(kmer) >= enc(rc(kmer))
        if is_rc not in found and len(set(kmer)) == 4:
            found[is_rc] = start
        if len(found) == 2:
            break
    for is_rc, start in found.items():
        kmer = g[start:start + K]
        pos = 6                                  # not the middle base: there the mirrored position coincides
        alt = "ACGT"[("ACGT".index(kmer[pos]) + 1) % 4]
        mut = kmer[:pos] + alt + kmer[pos + 1:]
        assert (enc(mut) >= enc(rc(mut))) == is_rc, "pick another site"
        s = run(oracle, ix, [(mut, 9)])
        assert s.best == -1                     # one variant k-mer: perfect = 0 → score 0 → nothing selected
        p = s.pileup(0)
        tot = (p[0] + p[1])
        rows = np.nonzero(tot.sum(axis=1))[0]
        assert len(rows) == 1
        row = int(rows[0])
        base = int(np.argmax(tot[row]))
        if not is_rc:
            assert row == start + pos and "ACGT"[base] == alt           # ALT at the true position
        else:
            assert row == start + (K - 1 - pos)                          # mirrored position ...
            assert "ACGT"[base] == g[row]                                 # ... and it is the REF base there


def test_q7_strand_assignment(oracle, hpv):
    ix, g = hpv
    kmer = g[3000:3021]
    fwd = run(oracle, ix, [(kmer, 5)]).pileup()
    rev = run(oracle, ix, [(rc(kmer), 5)]).pileup()
    assert fwd[0].sum() > 0 and fwd[1].sum() == 0      # read k-mer in genome orientation → forward arrays
    assert rev[1].sum() > 0 and rev[0].sum() == 0
    assert (fwd[0] == rev[1]).all()


def test_q8_q9_perfect_needs_exact_hit_count(oracle, sars_paths):
    ix = oracle.Index.build(K, sars_paths)
    wuhan = sim.load_genome(sim.SARS4[0]).tobytes().decode()
    polya = "A" * K
    assert polya in wuhan[-40:]
    s = run(oracle, ix, [(polya, 5), (wuhan[100:121], 5)])
    st = s.stats(0)
    # the poly-A k-mer hits > 16 entries of genome 0 → tallied as variant, not perfect
    assert st[0][0] == 1 and st[0][1] == 1
    # a k-mer shared by all four strains is perfect in each and unique in none
    assert (st[:, 0] >= 1).all() and st[:, 2].sum() == 0


def test_q10_selection_strict_greater_from_zero(oracle, hpv):
    ix, g = hpv
    far = "".join("ACGT"[(i * 7 + i // 3) % 4] for i in range(K))
    s = run(oracle, ix, [(far, 50)])
    assert s.best == -1 and len(s.variants()) == 0


def test_q12_noise_filter_details(oracle):
    L = 400
    rng = np.random.default_rng(3)
    fwd = np.zeros((L, 4), dtype=np.uint64)
    rev = np.zeros((L, 4), dtype=np.uint64)
    fwd[:, 0] = 1000
    fwd[:, 1] = rng.integers(0, 4, size=L)
    rev[:, 2] = rng.integers(0, 3, size=L)
    fwd[200, 3] = 300                                   # an outlier minor allele
    mx, mean, sd = oracle.baseline_noise(fwd, rev)
    frac200 = 300 / float(fwd[200].sum() + rev[200].sum())
    window = range(200 - 50, 200 + 50)                  # output i-50: window [p-49, p+50]
    assert all(mx[p] < frac200 for p in window)         # the outlier is rejected wherever it is in the window
    assert mx[200 + 60] < frac200 and mx[200 - 60] < frac200
    assert (mx >= 0).all() and mx.max() <= 4 / 1000.0
    with pytest.raises(ValueError):                     # len < 100: the reference indexes out of bounds
        oracle.baseline_noise(fwd[:50], rev[:50])


def test_q13_end_filter_and_breadth_denominator(oracle, hpv):
    ix, g = hpv
    reads = [g[0:150]] * 5 + [g[-150:]] * 5
    b, off = reads_from_strings(reads)
    c = oracle.Counts.count(K, b, off.astype(np.uint64), 3, 1000000, 1)
    s = oracle.Sample(ix, oracle.Params.defaults(k=K), [c])
    major, minor, breadth, depth = s.summary()
    # positions [k, len-k) only: 150-2-21 covered on the left (bucket slice drops ends), denominator = full length
    p = s.pileup()
    covered_all = np.count_nonzero((p[0] + p[1]).sum(axis=1))
    inside = np.count_nonzero((p[0] + p[1]).sum(axis=1)[K:len(g) - K])
    assert inside < covered_all
    assert breadth == inside / len(g)


def test_q14_q16_variant_gates(oracle, hpv):
    ix, g = hpv
    r1, o1, r2, o2, truth = sim.simulate_pairs(sim.load_genome(sim.HPV16), 1500, 21)
    cs = [oracle.Counts.count(K, b, o.astype(np.uint64), 3, 1000000, 2) for b, o in ((r1, o1), (r2, o2))]
    planted_minor = {int(p) + 1 for p, af in zip(truth["pos"], truth["af"]) if 0.15 < af < 0.5}
    # Q16: a --min-depth above the total depth drops every minor variant, majors stay (no depth gate on them)
    v = oracle.Sample(ix, oracle.Params.defaults(k=K, min_depth=10 ** 6), cs).variants()
    assert len(v) > 0 and (v["af"] >= 0.5).all()
    # at the default gate the planted iSNVs at AF 0.2 / 0.4 are there
    v1 = oracle.Sample(ix, oracle.Params.defaults(k=K, min_depth=1), cs).variants()
    assert planted_minor & set(v1["pos"].tolist()) and (v1["af"] < 0.5).any()
    assert (v1["sor"] <= 6.0).all()                      # Q14: everything kept passed sor <= --strand_odds
    # min_variant_depth gate
    v2 = oracle.Sample(ix, oracle.Params.defaults(k=K, min_depth=1, min_variant_depth=10 ** 6), cs).variants()
    assert (v2["af"] >= 0.5).all()
    # --no-strand-filter: SOR is reported as strand_odds_max + 1 and nothing is strand-filtered
    v3 = oracle.Sample(ix, oracle.Params.defaults(k=K, min_depth=1, no_strand_filter=1), cs).variants()
    assert (v3["sor"] == 7.0).all() and len(v3) >= len(v1)
    # --no-strand-balance-filter with a ratio nothing can reach: SOR = -1, both strand gates skipped
    v4 = oracle.Sample(ix, oracle.Params.defaults(k=K, min_depth=1, no_strand_balance_filter=1, strand_balance_ratio=0.9), cs).variants()
    assert (v4["sor"] == -1.0).all()


def test_q18_vcf_text_format(oracle, hpv):
    ix, g = hpv
    r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.HPV16), 400, 11)
    cs = [oracle.Counts.count(K, b, o.astype(np.uint64), 3, 1000000, 2) for b, o in ((r1, o1), (r2, o2))]
    s = oracle.Sample(ix, oracle.Params.defaults(k=K), cs)
    txt = s.vcf_text("some/dir/rep1_R1.fastq.gz").splitlines()
    assert txt[0] == "##fileformat=VCFv4.5" and txt[1] == "##source=bronko-v0.1.0"
    assert txt[2] == "##reference=file://some/dir/rep1_R1.fastq.gz"
    assert txt[3] == "##contig=<ID=HPV16REF,length=7906>"
    assert txt[7] == '##INFO=<ID=SOR,Number=4,Type=Float,Description="SOR">'
    assert txt[8] == "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO"
    import re
    pat = re.compile(r"^HPV16REF\t\d+\t\.\t[ACGT]\t[ACGT]\t\.\tPASS\tDP=\d+;AF=\d\.\d{3};DP4=\d+,\d+,\d+,\d+;SOR=-?\d+\.\d{3}$")
    assert len(txt) > 9 and all(pat.match(line) for line in txt[9:])
    assert s.pileup_text().splitlines()[0] == "reference\tindex\tref\tA\tC\tG\tT\ta\tc\tg\tt"


def test_q22_clean_sample_id(oracle):
    cases = {"a/b/rep1_R1.fastq.gz": "rep1_R1", "x.fq": "x", "x.fq.gz": "x", "s.fastq.fastq": "s",
             "weird.fna.gz": "weird.", "reads.txt": "reads", "noext": "noext", "a.b.fasta": "a.b", "q.fnq": "q"}
    for path, want in cases.items():
        assert oracle.clean_sample_id(path) == want, path


def test_q23_unmapped_uses_best_genome_tallies(oracle, hpv):
    ix, g = hpv
    r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.HPV16), 300, 12)
    cs = [oracle.Counts.count(K, b, o.astype(np.uint64), 3, 1000000, 2) for b, o in ((r1, o1), (r2, o2))]
    s = oracle.Sample(ix, oracle.Params.defaults(k=K), cs)
    uc = sum(c.stats()[3] for c in cs)
    st = [s.stats(f)[s.best] for f in range(2)]
    assert s.unmapped() == uc - sum(int(x[0] + x[1]) for x in st)
