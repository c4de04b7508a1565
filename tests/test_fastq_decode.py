"""The host decode stage of the C ABI (bk_fastq_decode, include/bronko_b200.h) — the KMC reader contract of SURVEY.md
Appendix B — through the product library itself: it needs no context and no GPU."""
import gzip
import threading

import numpy as np
import pytest

import bronko_b200


def parse_py(data: bytes):
    """the contract, restated: 4-line records, line 1 of each is the sequence, '\\r' before the line end dropped,
    a last line without a newline counts"""
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    return [ln[:-1] if ln.endswith(b"\r") else ln for i, ln in enumerate(lines) if i % 4 == 1]


def decode(path):
    d = bronko_b200.DecodedReads(str(path))
    try:
        reads = []
        for bases, off in d.chunks():
            raw = bases.tobytes()
            reads += [raw[off[i]:off[i + 1]] for i in range(len(off) - 1)]
        return reads
    finally:
        d.close()


def fastq_bytes(seqs, eol=b"\n", last_eol=True):
    out = b"".join(b"@r%d some text" % i + eol + s + eol + b"+" + eol + b"I" * len(s) + eol for i, s in enumerate(seqs))
    return out if last_eol else out[:-len(eol)]


@pytest.mark.parametrize("gz", [False, True])
@pytest.mark.parametrize("eol,last_eol", [(b"\n", True), (b"\r\n", True), (b"\n", False)])
def test_decode_matches_contract(tmp_path, gz, eol, last_eol):
    rng = np.random.default_rng(3)
    seqs = [bytes(rng.choice(np.frombuffer(b"ACGTNacgt", dtype=np.uint8), size=int(n)).tobytes())
            for n in rng.integers(0, 400, size=3000)]
    data = fastq_bytes(seqs, eol, last_eol)
    p = tmp_path / ("r.fastq.gz" if gz else "r.fastq")
    p.write_bytes(gzip.compress(data) if gz else data)
    got = decode(p)
    assert got == parse_py(data) == seqs


def test_lines_longer_than_a_block_and_truncated_records(tmp_path):
    big = b"ACGT" * (3 << 20)                                   # 12 MB: spans the 8 MB inflate blocks
    data = fastq_bytes([b"AC", big, b"GGT"]) + b"@tail\nACGTA"   # a record cut off after its sequence line, no newline
    p = tmp_path / "long.fq"
    p.write_bytes(data)
    got = decode(p)
    assert [len(x) for x in got] == [2, len(big), 3, 5] and got == parse_py(data)
    p2 = tmp_path / "empty.fq.gz"
    p2.write_bytes(gzip.compress(b""))
    assert decode(p2) == []
    with pytest.raises(bronko_b200.BkError):
        decode(tmp_path / "missing.fastq")


def test_decode_runs_on_several_threads(tmp_path):
    """thread-safe and GIL-free: the files of the next samples are decoded while the GPU works"""
    files = []
    for t in range(6):
        seqs = [b"ACGT"[(i + t) % 4:(i + t) % 4 + 1] * (50 + (i * 7 + t) % 100) for i in range(20000)]
        p = tmp_path / ("t%d.fastq.gz" % t)
        p.write_bytes(gzip.compress(fastq_bytes(seqs), 1))
        files.append((p, seqs))
    out = [None] * len(files)
    th = [threading.Thread(target=lambda i=i: out.__setitem__(i, decode(files[i][0]))) for i in range(len(files))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for (p, seqs), got in zip(files, out):
        assert got == seqs


@pytest.mark.parametrize("gz", [False, True])
@pytest.mark.parametrize("shift", [-2, -1, 0, 1, 2])
def test_line_ends_at_block_boundaries(tmp_path, gz, shift):
    """the 8 MB inflate blocks may end anywhere: inside a sequence line, between its '\\r' and '\\n', right after the
    '\\n', inside the header of the next record"""
    block = 8 << 20
    first = b"@pad " + b"x" * 100 + b"\r\nACGTACGT\r\n+\r\nIIIIIIII\r\n"
    seq = b"GATTACA" * 11
    # choose the second header so that the '\n' ending the second sequence line is byte (block + shift - 1)
    head_len = block + shift - 1 - len(first) - len(seq) - 1 - 2        # "\r\n" after the header, '\r' before the '\n'
    second = b"@" + b"h" * (head_len - 1) + b"\r\n" + seq + b"\r\n+\r\n" + b"I" * len(seq) + b"\r\n"
    data = first + second + b"@last\r\nTTTT\r\n+\r\nIIII\r\n"
    assert data[block + shift - 1:block + shift] == b"\n" and data[block + shift - 2:block + shift - 1] == b"\r"
    p = tmp_path / ("b.fastq.gz" if gz else "b.fastq")
    p.write_bytes(gzip.compress(data, 1) if gz else data)
    assert decode(p) == [b"ACGTACGT", seq, b"TTTT"] == parse_py(data)
