"""The `bronko` host CLI (bronko_b200/csrc/bronko_main.cpp): `build` on the CPU, `call` end to end on the GPU
from FASTQ(.gz) files to VCF / pileup TSV / overview TSV / MFA, against the oracle's text."""
import os
import subprocess

import numpy as np
import pytest

from bronko_b200 import sim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BRONKO = os.path.join(ROOT, "bronko_b200", "csrc", "bronko")


def run(args, **kw):
    return subprocess.run([BRONKO] + args, capture_output=True, text=True, **kw)


@pytest.fixture(scope="module", autouse=True)
def built():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "bronko_b200", "csrc"), "-s", "bronko"])


def test_build_matches_reference_fixture(oracle, hpv_fasta, hpv_bkdb_bytes, tmp_path):
    """reference tests/build_tests.rs runs these three builds and only checks the exit code; here the HPV16
    k=21 output is also compared with the bundled hpv.bkdb (same map, same size)."""
    out = str(tmp_path / "bronko")
    r = run(["build", "-g"] + [sim.genome_path(n) for n in sim.SARS4] + ["-t", "2", "-o", out])
    assert r.returncode == 0, r.stderr
    r = run(["build", "-g", hpv_fasta, "-k", "19", "-t", "2", "-o", out])
    assert r.returncode == 0, r.stderr
    r = run(["build", "-g", hpv_fasta, "-t", "2", "-o", out])
    assert r.returncode == 0, r.stderr
    mine = oracle.Index.load(out + ".bkdb")
    ref = oracle.Index.decode(hpv_bkdb_bytes)
    assert mine.file_size == len(hpv_bkdb_bytes)
    a, b = mine.export(), ref.export()
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and a[2].tobytes() == b[2].tobytes()
    assert mine.genomes() == ref.genomes()


def test_argument_checks_exit_1(hpv_fasta, tmp_path):
    assert run(["build", "-g", hpv_fasta, "-k", "20"]).returncode == 1          # even k
    assert run(["build", "-g", hpv_fasta, "-k", "33"]).returncode == 1
    assert run(["build", "-g", "genome.txt"]).returncode == 1                    # suffix check
    assert run(["call", "-d", "x.bkdb", "-g", hpv_fasta, "-r", "a.fq"]).returncode == 1     # db and genomes
    assert run(["call", "-r", "a.fq"]).returncode == 1                           # neither
    assert run(["call", "-d", "x.bkdb", "-r", "reads.txt"]).returncode == 1      # -r suffix check
    assert run(["call", "-d", "x.bkdb", "-1", "a.fq", "b.fq", "-2", "c.fq"]).returncode == 1   # pair count
    assert run(["call", "-d", "x.bkdb", "-r", "a.fq", "--noise-multiplier", "0.5"]).returncode == 1
    assert run(["call", "-d", "x.bkdb", "-r", "a.fq", "--min-af", "1.5"]).returncode == 1
    assert run(["build"]).returncode == 2 and run([]).returncode == 2            # arg_required_else_help
    # what needletail rejects (src/build.rs:156-159 logs "<error> | Failed to parse fasta file: <path>" and exits 1)
    empty, junk, short = tmp_path / "empty.fa", tmp_path / "junk.fa", tmp_path / "short.fa"
    empty.write_text(""); junk.write_text("ACGT\n>x\nACGT\n"); short.write_text(">s\nACGTACGT\n")
    for bad in (empty, junk, short):
        r = run(["build", "-g", str(bad), "-o", str(tmp_path / "bad")])
        assert r.returncode == 1 and "Failed to parse fasta file: " + str(bad) in r.stderr, r.stderr


@pytest.mark.gpu
def test_call_end_to_end(oracle, sars_paths, tmp_path):
    from util import oracle_sample
    import bronko_b200
    out = tmp_path / "out"
    db = str(tmp_path / "sars4")
    assert run(["build", "-g"] + sars_paths + ["-o", db]).returncode == 0
    oi = oracle.Index.load(db + ".bkdb")
    pairs, singles, expect = [], [], {}
    for s in range(4):
        r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[1 if s < 3 else 2]), 500, sim.SEED0 + 50 + s)
        f1, f2 = str(tmp_path / ("s%d_R1.fastq.gz" % s)), str(tmp_path / ("s%d_R2.fastq" % s))
        sim.write_fastq(f1, r1, o1, "s%d" % s, 1)
        sim.write_fastq(f2, r2, o2, "s%d" % s, 2)
        if s == 0:
            singles.append(f1)
            expect[f1] = oracle_sample(oi, [(r1, o1)], bronko_b200.CallArgs())[1]
        else:
            pairs.append((f1, f2))
            expect[f1] = oracle_sample(oi, [(r1, o1), (r2, o2)], bronko_b200.CallArgs())[1]
    r = run(["call", "-d", db + ".bkdb", "-r"] + singles + ["-1"] + [p[0] for p in pairs] + ["-2"] + [p[1] for p in pairs] +
            ["-o", str(out), "--pileup", "--alignment", "--keep-kmer-info", "-t", "2"])
    assert r.returncode == 0, r.stderr
    rows = (out / "bronko_overview.tsv").read_text().splitlines()
    assert rows[0] == ("filename\tselected_genome\tnum_major_variants\tnum_minor_variants\tbreadth_coverage\t"
                       "depth_coverage\tnum_perfect_kmers\tnum_variant_kmers\tnum_unmapped_kmers")
    order = singles + [p[0] for p in pairs]
    assert [x.split("\t")[0] for x in rows[1:]] == order            # SE first, then PE, in argument order
    for f1, row in zip(order, rows[1:]):
        o = expect[f1]
        major, minor, breadth, depth = o.summary()
        best_name = oi.genomes()[o.best][0]
        st = [o.stats(f)[o.best] for f in range(o.n_files)]
        want = "\t".join([f1, best_name, str(major), str(minor), "%.4f" % breadth, "%.4f" % depth,
                          str(sum(int(x[0]) for x in st)), str(sum(int(x[1]) for x in st)), str(o.unmapped())])
        assert row == want
        stem = os.path.basename(f1).replace(".fastq.gz", "")
        assert (out / (stem + ".vcf")).read_text() == o.vcf_text(f1)
        assert (out / (stem + ".tsv")).read_text() == o.pileup_text()
        assert (out / (stem + "_counts.txt")).exists()
    # three samples (>= 3) chose genome OM223929.1 with breadth >= 0.9 → one alignment
    mfa = (out / "OM223929.1.mfa").read_text().splitlines()
    assert mfa[0] == ">OM223929.1" and len(mfa) == 2 * (1 + 3)
    assert [m[1:] for m in mfa[2::2]] == ["s0_R1", "s1_R1", "s2_R1"]
    assert len({len(x) for x in mfa[1::2]}) == 1 and set("".join(mfa[1::2])) <= set("ACGT")
    assert not (out / "ON765678.1.mfa").exists()                     # only one sample picked that genome
    # k mismatch between -k and the db → exit 1 (src/call.rs:193-197)
    assert run(["call", "-d", db + ".bkdb", "-r", singles[0], "-k", "19", "-o", str(out)]).returncode == 1


@pytest.mark.gpu
def test_call_with_genomes_and_bgzf_reads(oracle, sars_paths, tmp_path):
    """`bronko call -g <fasta>...` builds the index on the fly (src/call.rs:170-178) and must give what `call -d` gives
    on the db `bronko build` wrote; R1 comes as BGZF (inflated by the GPU's decompression engine), R2 as plain text."""
    import zlib
    from util import oracle_sample
    import bronko_b200
    r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[3]), 400, sim.SEED0 + 58)
    f1_plain, f2 = str(tmp_path / "g_R1.fastq"), str(tmp_path / "g_R2.fastq")
    sim.write_fastq(f1_plain, r1, o1, "g", 1)
    sim.write_fastq(f2, r2, o2, "g", 2)
    text = open(f1_plain, "rb").read()
    f1 = str(tmp_path / "g_R1.fastq.gz")
    with open(f1, "wb") as f:
        for i in range(0, len(text), 65280):
            chunk = text[i:i + 65280]
            co = zlib.compressobj(6, zlib.DEFLATED, -15)
            body = co.compress(chunk) + co.flush()
            f.write(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + (len(body) + 25).to_bytes(2, "little") + body +
                    (zlib.crc32(chunk) & 0xFFFFFFFF).to_bytes(4, "little") + len(chunk).to_bytes(4, "little"))
        f.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    out_g, out_d = tmp_path / "out_g", tmp_path / "out_d"
    r = run(["call", "-g"] + sars_paths + ["-1", f1, "-2", f2, "-o", str(out_g), "--pileup", "-t", "2"])
    assert r.returncode == 0, r.stderr
    db = str(tmp_path / "sars4")
    assert run(["build", "-g"] + sars_paths + ["-o", db]).returncode == 0
    r = run(["call", "-d", db + ".bkdb", "-1", f1, "-2", f2, "-o", str(out_d), "--pileup", "-t", "2"])
    assert r.returncode == 0, r.stderr
    for name in ("g_R1.vcf", "g_R1.tsv", "bronko_overview.tsv"):
        assert (out_g / name).read_bytes() == (out_d / name).read_bytes(), name
    osample = oracle_sample(oracle.Index.build(21, sars_paths), [(r1, o1), (r2, o2)], bronko_b200.CallArgs())[1]
    assert (out_g / "g_R1.vcf").read_text() == osample.vcf_text(f1)
    assert (out_g / "g_R1.tsv").read_text() == osample.pileup_text()
