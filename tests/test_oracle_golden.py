"""The oracle pinned against every golden artefact the reference holds for this path:
the two assign_buckets vectors (src/lcb.rs:146-154) and test_data/hpv.bkdb (+ HPV16.fa)."""
import math
import os

import numpy as np
import pytest

REF_BKDB = "/root/reference/test_data/hpv.bkdb"


def test_assign_buckets_astring(oracle):          # src/lcb.rs:146-149
    assert oracle.assign_buckets(0, 4) == [1, 2, 3, 4]


def test_assign_buckets_kstring(oracle):          # src/lcb.rs:151-154
    want = [238258108556, 47877379752, 215381104296, 227729135272, 235782198952, 237342480040, 238258108557,
            238236915369, 238248449705, 238254544553, 238258108558, 238257944234, 238258089642, 238258095018,
            238258106282, 238258108559, 238258108483, 238258108525, 238258108547]
    assert oracle.assign_buckets(41547505179, 19) == want


@pytest.mark.parametrize("k", [3, 4, 5, 6])
def test_assign_buckets_is_a_bijection(oracle, k):
    """(index j, k-mer with base j wild-carded) -> [1, k*4^(k-1)] is a bijection and
    bucket j does not depend on base j."""
    seen = {}
    for kmer in range(4 ** k):
        b = oracle.assign_buckets(kmer, k)
        for j, bid in enumerate(b):
            sh = 2 * (k - 1 - j)
            masked = (j, kmer & ~(3 << sh))
            assert seen.setdefault(bid, masked) == masked
    assert len(seen) == k * 4 ** (k - 1)
    assert min(seen) == 1 and max(seen) == k * 4 ** (k - 1)


def test_hpv_bkdb_parses_to_eof(oracle, hpv_bkdb_bytes):
    ix = oracle.Index.decode(hpv_bkdb_bytes)
    assert ix.consumed == len(hpv_bkdb_bytes) == 2810072
    assert (ix.k, ix.meta_k) == (21, 21)
    assert (ix.n_keys, ix.n_entries) == (165603, 165606)
    (gname, seqs), = ix.genomes()
    assert gname == "HPV16" and [(n, ln) for n, ln, _ in seqs] == [("HPV16REF", 7906)]


def test_hpv_bkdb_fixture_is_the_reference_file(hpv_bkdb_bytes):
    if not os.path.exists(REF_BKDB):
        pytest.skip("reference tree not mounted (GPU box)")
    assert open(REF_BKDB, "rb").read() == hpv_bkdb_bytes


def test_rebuild_from_fasta_equals_bundled_bkdb(oracle, hpv_bkdb_bytes, hpv_fasta):
    """Known-answer test of lcb.rs + build.rs + the bincode layout: same keys, entries and per-key order."""
    a = oracle.Index.decode(hpv_bkdb_bytes).export()
    b = oracle.Index.build(21, [hpv_fasta]).export()
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and a[2].tobytes() == b[2].tobytes()


def test_oracle_bkdb_writer_roundtrip(oracle, hpv_bkdb_bytes, tmp_path):
    ix = oracle.Index.decode(hpv_bkdb_bytes)
    p = str(tmp_path / "x.bkdb")
    ix.save(p)
    ix2 = oracle.Index.load(p)
    assert ix2.consumed == ix2.file_size
    a, b = ix.export(), ix2.export()
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and a[2].tobytes() == b[2].tobytes()
    assert ix.genomes() == ix2.genomes()


def test_thompson_tau_anchors(oracle):
    """statrs StudentsT::inverse_cdf restated (parity unpinned): anchors from SURVEY.md §8c (scipy)."""
    for n, want in ((3, 1.15469990524389), (10, 2.60593403789741), (100, 4.0839659911532), (300, 4.4320798372553)):
        assert abs(oracle.thompson_tau(n) - want) < 1e-11 * want


def test_thompson_tau_against_scipy(oracle):
    stats = pytest.importorskip("scipy.stats")
    for n in range(3, 301):
        t = stats.t.ppf(1 - 0.001 / n, n - 2)
        want = t * (n - 1) / (math.sqrt(n) * math.sqrt(n - 2 + t * t))
        assert abs(oracle.thompson_tau(n) - want) < 1e-12 * want
