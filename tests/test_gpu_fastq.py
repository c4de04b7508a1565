"""The decode stage on the device (bk_reads_push_fastq / _fastq_mem: inflate — BGZF on the GPU's decompression engine —
and FASTQ parsing by kernels) against the host reader of the same library (bk_fastq_decode, whose KMC reader contract is
pinned in tests/test_fastq_decode.py) and against the oracle: same reads, same counts, for every container (text, gzip,
multi-member gzip, BGZF) and every edge of the text (CRLF, no final newline, truncated records, empty lines, lines
longer than a tile, records across segment boundaries)."""
import gzip
import os
import zlib

import numpy as np
import pytest

from bronko_b200 import sim

pytestmark = pytest.mark.gpu


def bgzf(data, block=65280):
    out = []
    for i in range(0, max(len(data), 1), block):
        chunk = data[i:i + block]
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = co.compress(chunk) + co.flush()
        out.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + (len(body) + 25).to_bytes(2, "little") + body +
                   (zlib.crc32(chunk) & 0xFFFFFFFF).to_bytes(4, "little") + len(chunk).to_bytes(4, "little"))
    out.append(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    return b"".join(out)


def containers(text):
    half = len(text) // 2
    return {"text": text, "gzip": gzip.compress(text, 1), "multi_member": gzip.compress(text[:half], 1) + gzip.compress(text[half:], 1),
            "bgzf": bgzf(text), "bgzf_small_blocks": bgzf(text, 997)}


def fastq_of(seqs, eol=b"\n", last_eol=True):
    recs = []
    for i, s in enumerate(seqs):
        recs.append(b"@r%d some comment" % i + eol + s + eol + b"+" + eol + b"I" * len(s) + eol)
    t = b"".join(recs)
    return t if last_eol else t[:-len(eol)]


@pytest.fixture(scope="module")
def ctx(sars_paths):
    import bronko_b200
    c = bronko_b200.Bronko(0)
    c.build_index(21, sars_paths)
    yield c
    c.close()


def host_reads(tmp_path, data, name):
    import bronko_b200
    p = tmp_path / name
    p.write_bytes(data)
    d = bronko_b200.DecodedReads(str(p))
    chunks = [(b.copy(), o.copy()) for b, o in d.chunks()]
    d.close()
    return str(p), chunks


def _extract(c, sample):
    """(k-mers, counts, KMC numbers) of file slot 0 of the sample the context just finished (also when no genome could
    be selected: the counts are there all the same)."""
    import ctypes as C
    from bronko_b200 import _lib as L
    if sample is not None:
        k, v = sample.kmers(0)
        return k, v, sample.kmc_stats(0)
    n = C.c_uint64(0)
    c._check(c._lib.bk_kmer_counts_get(c.h, 0, None, None, C.byref(n)))
    km, ct = np.zeros(n.value, dtype=np.uint64), np.zeros(n.value, dtype=np.uint32)
    if n.value:
        c._check(c._lib.bk_kmer_counts_get(c.h, 0, L.ptr(km), L.ptr(ct), C.byref(n)))
    res = L.SampleResult()
    c._check(c._lib.bk_sample_result_get(c.h, C.byref(res)))
    k0 = res.kmc[0]
    return km, ct, (k0.total_reads, k0.total_kmers, k0.unique_kmers, k0.unique_counted)


def _finish(c):
    import bronko_b200
    try:
        return _extract(c, c.finish())
    except bronko_b200.BkError as e:
        if e.code != -4:                                # -4: no genome selected
            raise
        return _extract(c, None)


def run_device(c, data, args=None):
    import bronko_b200
    c.begin(args or bronko_b200.CallArgs(min_kmers=1))
    c.push_fastq_mem(0, data)
    return _finish(c)


def run_host(c, chunks, args=None):
    import bronko_b200
    c.begin(args or bronko_b200.CallArgs(min_kmers=1))
    if not chunks:
        c.push(0, np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.uint32))
    for b, o in chunks:
        c.push(0, b, o)
    return _finish(c)


EDGE_TEXTS = {
    "plain": fastq_of([b"ACGTACGTACGTAGCTAGCTAGCATCGATCGAT" * 3, b"TTGACCAGTACCAGTTGACAGTTTGACCCAGTAG" * 2]),
    "crlf": fastq_of([b"ACGTACGTACGTAGCTAGCTAGCATCGATCGAT" * 3, b"TTGACCAGTACCAGTTGACAGTTTGACCCAGTAG" * 2], eol=b"\r\n"),
    "no_final_newline": fastq_of([b"ACGTACGTACGTAGCTAGCTAGCATCGATCGAT" * 3, b"TTGACCAGTACCAGTTGACAGTTTGACCCAGTAG" * 2], last_eol=False),
    "ends_in_sequence_line": b"@a\n" + b"ACGTTGCATGCATGCATGGGCATGCAAACGT" * 2 + b"\n+\nIIII\n@b\n" + b"GGGTTTCCCAAAGGGTTTCCCAAAG" * 2,
    "ends_after_header": b"@a\n" + b"ACGTTGCATGCATGCATGGGCATGCAAACGT" * 2 + b"\n+\nIIII\n@b\n",
    "empty_sequence_lines": b"@a\n\n+\n\n@b\n" + b"ACGTTGCATGCATGCATGGGCATGCAAACGT" * 2 + b"\n+\nII\n@c\n\n+\n\n",
    "only_newlines": b"\n\n\n\n\n\n\n",
    "one_byte": b"@",
    "long_line": fastq_of([b"ACGGTCATTG" * 1500, b"TTGACCAGTACCAGTTGACAGTTTGACCCAGTAG" * 2]),          # 15 kb read: crosses tiles
    "junk_and_lower_case": fastq_of([b"acgtacgtNNNNacgtagctagcatcgatcgatacgtagctagctagc", b"ACGT*ACGT-ACGTAGCTAGCTAGCTAGCATCGACTAGCTAGCTACGACT"]),
}


@pytest.mark.parametrize("name", sorted(EDGE_TEXTS))
def test_edge_texts_every_container(ctx, tmp_path, name):
    text = EDGE_TEXTS[name]
    for cname, data in containers(text).items():
        _, chunks = host_reads(tmp_path, data, "%s_%s.fq%s" % (name, cname, "" if cname == "text" else ".gz"))
        hk, hv, hs = run_host(ctx, chunks)
        dk, dv, ds = run_device(ctx, data)
        assert ds == hs, (name, cname, ds, hs)
        assert np.array_equal(dk, hk) and np.array_equal(dv, hv), (name, cname)
        info = ctx.decode_info(0)
        assert info["mode"] == {"text": 1, "gzip": 2, "multi_member": 2, "bgzf": 3, "bgzf_small_blocks": 3}[cname], (cname, info)


def test_empty_file(ctx):
    import bronko_b200
    ctx.begin(bronko_b200.CallArgs())
    ctx.push_fastq_mem(0, b"")
    with pytest.raises(bronko_b200.BkError) as e:
        ctx.finish()
    assert e.value.code == -4


@pytest.mark.parametrize("segment", [None, 1 << 16, 150_000])
def test_simulated_sample_all_containers_match_oracle(sars_paths, oracle, tmp_path, monkeypatch, segment):
    """A simulated paired sample as FASTQ files: device decode (every container, one or many segments with records
    carried across the boundaries) → the same result as the oracle on the reads themselves."""
    import bronko_b200
    from util import assert_sample_equal, oracle_sample
    if segment:
        monkeypatch.setenv("BK_FQ_SEGMENT", str(segment))
    c = bronko_b200.Bronko(0)
    try:
        c.build_index(21, sars_paths)
        oi = oracle.Index.build(21, sars_paths)
        r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[1]), 300, sim.SEED0 + 71)
        args = bronko_b200.CallArgs()
        counts, osample = oracle_sample(oi, [(r1, o1), (r2, o2)], args)
        texts = []
        for m, (b, o) in enumerate(((r1, o1), (r2, o2))):
            raw = b.tobytes()
            texts.append(b"".join(b"@s_%d/%d\n%s\n+\n%s\n" % (i, m + 1, raw[o[i]:o[i + 1]], b"I" * int(o[i + 1] - o[i])) for i in range(len(o) - 1)))
        for cname in ("text", "gzip", "bgzf"):
            files = [containers(t)[cname] for t in texts]
            paths = []
            for m, data in enumerate(files):
                p = tmp_path / ("s_%s_R%d.fastq%s" % (cname, m + 1, "" if cname == "text" else ".gz"))
                p.write_bytes(data)
                paths.append(str(p))
            c.begin(args)
            c.push_fastq(0, paths[0])
            c.push_fastq_mem(1, files[1])
            g = c.finish()
            assert_sample_equal(g, counts, osample)
            info = c.decode_info(0)
            assert info["n_reads"] == len(o1) - 1 and info["n_bases"] == len(r1)
            if segment:
                assert info["segments"] > 1
            t = c.stage_times()
            assert t["decode_ms"] > 0
    finally:
        c.close()


def test_corrupt_bgzf_is_reported(tmp_path, sars_paths):
    """A BGZF block whose deflate payload is damaged: the decompression engine faults on it, which costs the process its
    CUDA context — so this runs in a process of its own (the CLI, like a user would): the run must end with an error
    message and exit code 1, not hang and not produce a VCF.  (BK_NO_DECOMP_ENGINE=1 sends BGZF through zlib, which
    reports the damage without losing the context.)"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bronko = os.path.join(root, "bronko_b200", "csrc", "bronko")
    subprocess.check_call(["make", "-C", os.path.join(root, "bronko_b200", "csrc"), "-s", "bronko"])
    text = fastq_of([b"ACGTACGTACGTAGCTAGCTAGCATCGATCGAT" * 3] * 50)
    data = bytearray(bgzf(text, 997))
    data[40] ^= 0x55                                     # inside the first member's deflate payload
    bad = tmp_path / "bad.fastq.gz"
    bad.write_bytes(bytes(data))
    for env_extra in ({}, {"BK_NO_DECOMP_ENGINE": "1"}):
        out = tmp_path / ("out%d" % len(env_extra))
        r = subprocess.run([bronko, "call", "-g"] + sars_paths + ["-r", str(bad), "-o", str(out)], capture_output=True, text=True,
                           timeout=120, env=dict(os.environ, **env_extra))
        assert r.returncode == 1, (r.returncode, r.stderr[-400:])
        assert "ERROR" in r.stderr
        assert not (out / "bad.vcf").exists()
