"""CPU tests of the product's host logic and of the counting-stage algorithm the CUDA kernels are built
from (stepped on the CPU by tests/emul), each against the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from bronko_b200 import sim
from emul_lib import Emul, lib as emul_lib, ptr
from util import reads_from_strings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sars_emul(sars_paths):
    return Emul.from_fasta(21, sars_paths)


@pytest.fixture(scope="module")
def hpv_emul(hpv_bkdb_path):
    return Emul.from_bkdb(hpv_bkdb_path)


def test_builder_matches_oracle_and_bundled_db(oracle, sars_paths, sars_emul, hpv_fasta, hpv_bkdb_bytes):
    a, b = oracle.Index.build(21, sars_paths).export(), sars_emul.export()
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and a[2].tobytes() == b[2].tobytes()
    assert len(a[0]) == 703025 and len(a[2]) == 2501142
    h = Emul.from_fasta(21, [hpv_fasta]).export()
    g = oracle.Index.decode(hpv_bkdb_bytes).export()
    assert (h[0] == g[0]).all() and (h[1] == g[1]).all() and h[2].tobytes() == g[2].tobytes()


@pytest.mark.parametrize("k", [15, 19, 31])
def test_builder_other_k(oracle, hpv_fasta, k):
    a, b = oracle.Index.build(k, [hpv_fasta]).export(), Emul.from_fasta(k, [hpv_fasta]).export()
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and a[2].tobytes() == b[2].tobytes()


def test_bkdb_reader_and_writer(oracle, hpv_emul, hpv_bkdb_bytes, tmp_path):
    g = oracle.Index.decode(hpv_bkdb_bytes).export()
    h = hpv_emul.export()
    assert (h[0] == g[0]).all() and (h[1] == g[1]).all() and h[2].tobytes() == g[2].tobytes()
    p = str(tmp_path / "w.bkdb")
    hpv_emul.save(p)
    back = oracle.Index.load(p)                         # the oracle's reader decodes what the product wrote
    assert back.consumed == back.file_size
    r = back.export()
    assert (r[0] == g[0]).all() and r[2].tobytes() == g[2].tobytes()
    assert back.genomes() == oracle.Index.decode(hpv_bkdb_bytes).genomes()


def test_assign_buckets_closed_form_matches_loop_form(oracle):
    rng = np.random.default_rng(1)
    for k in (15, 21, 27, 31):                          # k = 31 exercises the u64 wrap-around (Q20)
        for _ in range(200):
            kmer = int(rng.integers(0, 2 ** 62, dtype=np.uint64)) & ((1 << (2 * k)) - 1)
            out = np.zeros(k, dtype=np.uint64)
            emul_lib().emul_assign_buckets(kmer, k, ptr(out))
            assert [int(x) for x in out] == oracle.assign_buckets(kmer, k)
            assert emul_lib().emul_revcomp(kmer, k) == oracle.lib().orc_reverse_complement(kmer, k)


def test_tau_table_matches_oracle(oracle):
    t = np.zeros(301)
    emul_lib().emul_tau_table(ptr(t))
    for n in range(3, 301):
        assert t[n] == oracle.thompson_tau(n)
    assert (t[:3] == 0).all()


def test_product_tau_table_against_scipy():
    """The product's own Student-t / Thompson-tau table (bk_io.cpp: tau_table, uploaded to the device as c_tau) against an
    INDEPENDENT implementation — scipy's t.ppf — to 1e-12 relative.  (The comparison with the oracle above cannot
    catch an error the two restatements share; this one can.)"""
    import math
    stats = pytest.importorskip("scipy.stats")
    t = np.zeros(301)
    emul_lib().emul_tau_table(ptr(t))                     # tests/emul links the product's bk_io.cpp: the same function the library calls
    for n in range(3, 301):
        q = stats.t.ppf(1 - 0.001 / n, n - 2)
        want = q * (n - 1) / (math.sqrt(n) * math.sqrt(n - 2 + q * q))
        assert abs(t[n] - want) < 1e-12 * want, n


def test_clean_sample_id_matches_oracle(oracle):
    buf = C.create_string_buffer(512)
    for p in ["a/b/rep1_R1.fastq.gz", "x.fq", "x.fq.gz", "s.fastq.fastq", "weird.fna.gz", "reads.txt", "noext",
              "a.b.fasta", "q.fnq", "/abs/dir.d/file.fa", ".hidden", "double.fq.fq"]:
        emul_lib().emul_clean_sample_id(p.encode(), buf, 512)
        assert buf.value.decode() == oracle.clean_sample_id(p), p


@pytest.mark.parametrize("dense", [False, True])
@pytest.mark.parametrize("strain,depth", [(0, 200), (1, 200), (3, 120)])
def test_counting_logic_matches_oracle(oracle, sars_emul, strain, depth, dense):
    """dense: k-mers with exactly one mismatch against their read's diagonal are counted on mismatch lines (two atomics
    per sequencing error) instead of being listed one by one — same counts, and most of the leftover volume is gone."""
    r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[strain]), depth, 100 + strain)
    for b, o in ((r1, o1), (r2, o2)):
        want = oracle.Counts.count(21, b, o.astype(np.uint64), 3, 1000000, 2)
        km, ct, st, dbg = sars_emul.count(b, o, dense=dense)
        wk, wc = want.get()
        assert st == want.stats()
        assert len(km) == len(wk) and (km == wk).all() and (ct.astype(np.uint64) == wc).all()
        assert dbg[1] < 0.2 * st[1]                     # most k-mers ride on runs, not the leftover path
        if dense:
            assert dbg[1] < 0.02 * st[1] and dbg[2] > 5 * dbg[1]       # ... and most of the rest on mismatch lines


def test_counting_edge_cases(oracle, hpv_emul):
    g = sim.load_genome(sim.HPV16).tobytes().decode()
    rcg = g[::-1].translate(str.maketrans("ACGT", "TGCA"))
    rng = np.random.default_rng(5)
    seqs = ["", "A", g[100:120], g[100:121], g[200:350], g[200:350].lower(), g[400:470] + "N" + g[471:560],
            g[600:640] + "n*" + g[642:700], "N" * 40, "GATTACA" * 30, g[1000:1100] + g[1103:1250],
            g[1500:1560] + "ACGT" + g[1560:1700], "TTTTTTTTTT" + g[0:140], g[-140:] + "GGGGGGGGGG", rcg[0:150],
            "CCCCC" + rcg[-100:] + "AAAAA", g[3000:3800], g, g[2000:2021], g[2001:2022],
            "".join("ACGT"[i] for i in rng.integers(0, 4, size=300))]
    # substitutions everywhere: next to the read ends, next to each other, runs of them, on both strands, as lower case / N
    def sub(s, at, to=None):
        s = list(s)
        for p in at:
            s[p] = to if to else "ACGT"[("ACGT".index(s[p].upper()) + 1) % 4]
        return "".join(s)
    r = g[4000:4150]
    seqs += [sub(r, [0]), sub(r, [149]), sub(r, [20]), sub(r, [21]), sub(r, [129]), sub(r, [128]), sub(r, [70]), sub(r, [70, 71]),
             sub(r, [70, 90]), sub(r, [70, 91]), sub(r, [70, 92]), sub(r, [10, 75, 140]), sub(r, [5, 6, 7, 8]), sub(r, [70], "n"),
             sub(r, [70], "t" if r[70] != "T" else "a"), sub(r, [30, 60, 61, 100, 121]), sub(r, range(40, 49)),
             sub(rcg[500:650], [0, 75, 149]), sub(g[0:150], [3, 30]), sub(g[-150:], [120, 147]), sub(g[5000:5040], [19]),
             sub(g[5000:5041], [20]), sub(g[5000:5042], [0, 41])]
    for ci in (1, 2):
        for dense in (False, True):
            b, off = reads_from_strings(seqs * 2)
            want = oracle.Counts.count(21, b, off.astype(np.uint64), ci, 1000000, 1)
            km, ct, st, _ = hpv_emul.count(b, off, ci=ci, dense=dense)
            wk, wc = want.get()
            assert st == want.stats()
            assert (km == wk).all() and (ct.astype(np.uint64) == wc).all()


def test_counting_with_full_leftover_queue(oracle, hpv_emul):
    """Queue overflow falls back to in-place counting: still exact."""
    r1, o1, _, _, _ = sim.simulate_pairs(sim.load_genome(sim.HPV16), 60, 7)
    want = oracle.Counts.count(21, r1, o1.astype(np.uint64), 3, 1000000, 1)
    for dense in (False, True):
        km, ct, st, _ = hpv_emul.count(r1, o1, desc_cap=8, dense=dense)
        wk, wc = want.get()
        assert st == want.stats() and (km == wk).all() and (ct.astype(np.uint64) == wc).all()


def test_c_abi_exports_every_declared_symbol():
    from bronko_b200 import _lib
    header = open(_lib.HEADER).read()
    declared = set(re.findall(r"\b(bk_[a-z_0-9]+)\s*\(", header))
    declared -= {"bk_ctx"}
    assert declared == set(_lib.SIGNATURES), "header and ctypes signatures disagree"
    L = _lib.lib()                                      # loads without a GPU
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.bk_version()


def test_no_cpu_fallback_without_device():
    """On a box without a B200 the context must fail loudly (no silent CPU path)."""
    import bronko_b200
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(bronko_b200.BkError) as e:
        bronko_b200.Bronko(0)
    assert e.value.code in (-6, -2)


def test_product_does_not_reference_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "bronko_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f in ("__init__.py",) and "oracle" not in txt, os.path.join(dirpath, f)


def test_simulator_is_seeded_and_shaped():
    a = sim.simulate_pairs(sim.load_genome(sim.HPV16), 50, 3)
    b = sim.simulate_pairs(sim.load_genome(sim.HPV16), 50, 3)
    assert (a[0] == b[0]).all() and (a[2] == b[2]).all()
    n = round(50 * 7906 / 300)
    assert len(a[1]) == n + 1 and len(a[0]) == n * 150 and set(np.unique(a[0]).tolist()) <= set(b"ACGT")
    assert len(a[4]["pos"]) == 30 and sorted(set(a[4]["af"].tolist())) == [0.03, 0.05, 0.1, 0.2, 0.4, 1.0]


def _canonical(kmers, k):
    """canonical form (src/lcb.rs:87-95) of an array of k-mer values"""
    out = np.empty_like(kmers)
    for i, v in enumerate(kmers.tolist()):
        r, x = 0, v
        for _ in range(k):
            r = (r << 2) | (3 - (x & 3))
            x >>= 2
        out[i] = v if v < r else r
    return out


@pytest.mark.parametrize("which,k", [("sars", 21), ("hpv", 21), ("hpv15", 15), ("hpv29", 29)])
def test_grouped_map_tables_are_an_exact_reindexing(sars_emul, hpv_emul, hpv_fasta, which, k):
    """group_slots / group_centers / group_buckets (bk_host.h) against k direct probes of the bucket table, with the
    lookup of k_map_grp stepped on the CPU: reference k-mers, all kinds of neighbours (one substitution anywhere — the
    case that must hit exactly one bucket —, two substitutions, substitutions at both ends) and random k-mers; the
    default bucket range, --use-full-kmer and an n_fixed that leaves one side empty."""
    e = {"sars": sars_emul, "hpv": hpv_emul}.get(which) or Emul.from_fasta(k, [hpv_fasta])
    genome = sim._CODE[sim.load_genome(sim.SARS4[1] if which == "sars" else sim.HPV16)].astype(np.uint64)
    rng = np.random.default_rng(11)
    starts = rng.integers(0, len(genome) - k, size=1500)
    ref = np.array([int("".join(str(int(b)) for b in genome[s:s + k]), 4) for s in starts], dtype=np.uint64)
    queries = [ref]
    for n_sub in (1, 1, 2):
        q = ref.copy()
        for _ in range(n_sub):
            pos = rng.integers(0, k, size=len(q)).astype(np.uint64)
            q ^= rng.integers(1, 4, size=len(q)).astype(np.uint64) << (np.uint64(2) * pos)
        queries.append(q)
    ends = ref ^ np.uint64(1) ^ (np.uint64(2) << np.uint64(2 * (k - 1)))
    queries += [ends, rng.integers(0, 4 ** k, size=500, dtype=np.uint64)]
    q = _canonical(np.concatenate(queries), k)
    for b0, b1 in ((2, k - 3), (0, k), (k // 2, k - k // 2 - 1 + (k // 2 + 1 > k - k // 2 - 1)), (0, 1)):
        if b0 < b1:
            assert e.group_check(q, b0, b1) == 0


def test_multi_genome_db_roundtrip_through_bkdb(oracle, sars_emul, sars_paths, tmp_path):
    """4-strain db (703,025 keys / 2,501,142 entries — large enough for the threaded sort and the per-file builder):
    product writer → product reader and → oracle reader, all three equal to the oracle's own build."""
    want = oracle.Index.build(21, sars_paths).export()
    p = str(tmp_path / "sars4.bkdb")
    sars_emul.save(p)
    back = Emul.from_bkdb(p).export()
    assert (back[0] == want[0]).all() and (back[1] == want[1]).all() and back[2].tobytes() == want[2].tobytes()
    r = oracle.Index.load(p).export()
    assert (r[0] == want[0]).all() and (r[1] == want[1]).all() and r[2].tobytes() == want[2].tobytes()
    # file order matters (entries of a key are in file order, then position): a permuted file list is another index
    perm = Emul.from_fasta(21, sars_paths[::-1]).export()
    assert (perm[0] == want[0]).all() and perm[2].tobytes() != want[2].tobytes()
    want_perm = oracle.Index.build(21, sars_paths[::-1]).export()
    assert (perm[1] == want_perm[1]).all() and perm[2].tobytes() == want_perm[2].tobytes()


def test_reads_pack_splits_clean_and_dirty_reads():
    """bk_reads_pack (host helper of the decode stage): reads of ACGT / acgt only become one continuous 2-bit stream
    (16 bases per u32, base i at bits 2 * (i % 16)), every other read is passed on as ASCII, order kept on both sides."""
    import bronko_b200
    from util import reads_from_strings
    rng = np.random.default_rng(11)
    seqs = ["".join("ACGT"[i] for i in rng.integers(0, 4, size=int(n))) for n in rng.integers(0, 200, size=60)]
    seqs[3] = seqs[3][:10] + "N" + seqs[3][10:]
    seqs[7] = seqs[7].lower()
    seqs[11] = "ACGT*ACGT"
    seqs[12] = ""
    seqs[40] = seqs[40][:5] + "n" + seqs[40][5:]
    b, off = reads_from_strings(seqs)
    packed, poff, rest, roff = bronko_b200.pack_reads(b, off)
    dirty = [s_ for s_ in seqs if set(s_) - set("ACGTacgt")]
    clean = [s_ for s_ in seqs if not (set(s_) - set("ACGTacgt"))]
    assert len(poff) - 1 == len(clean) and len(roff) - 1 == len(dirty)
    assert [rest[roff[i]:roff[i + 1]].tobytes().decode() for i in range(len(dirty))] == dirty
    total = int(poff[-1])
    codes = np.array([(int(packed[i // 16]) >> (2 * (i % 16))) & 3 for i in range(total)], dtype=np.uint8)
    got = ["".join("ACGT"[c] for c in codes[poff[i]:poff[i + 1]]) for i in range(len(clean))]
    assert got == [s_.upper() for s_ in clean]
