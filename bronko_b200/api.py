"""Host-side mirror of the reference's seam for the k-mer→pileup path (treangenlab/bronko
src/call.rs:151-402), on top of the C ABI (include/bronko_b200.h).

    reference (Rust)                                   here
    ------------------------------------------------   ------------------------------------------
    BronkoIndex decode / build_indexes (call.rs:170)   Bronko.load_index / Bronko.build_index
    CallArgs (cli.rs:61-166)                           CallArgs
    get_kmers (call.rs:630)                            Sample.kmers(file) / Sample.kmc_stats(file)
    map_kmers (call.rs:1257)                           Sample.mapping_data(file)
    pick_best_genome(_paired) (call.rs:422/452)        Sample.best_genome
    call_variants (call.rs:969)                        Sample.variants, .num_major, .num_minor, ...
    print_output / print_pileup (call.rs:735/648)      Sample.write_vcf / Sample.write_pileup

Errors keep the reference's messages; where the reference calls std::process::exit(1) this raises
BkError (the CLI turns it back into exit code 1).
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L
from ._lib import BkError, Params


@dataclass
class CallArgs:
    """The CallArgs fields the path consumes, defaults from the reference's src/consts.rs."""
    kmer: int = 21
    min_kmers: int = 3
    use_full_kmer: bool = False
    n_fixed: int = 2
    min_af: float = 0.03
    no_end_filter: bool = False
    no_strand_filter: bool = False
    no_strand_balance_filter: bool = False
    strand_balance_ratio: float = 0.1
    n_per_strand: int = 2
    strand_odds_max: float = 6.0
    min_depth: int = 300
    min_variant_depth: int = 3
    variant_multiplier: float = 1.5
    table_log2: int = 0

    def to_params(self):
        p = Params()
        L.lib().bk_params_default(C.byref(p))
        p.k, p.min_kmers, p.use_full_kmer, p.n_fixed = self.kmer, self.min_kmers, int(self.use_full_kmer), self.n_fixed
        p.min_af, p.no_end_filter, p.no_strand_filter = self.min_af, int(self.no_end_filter), int(self.no_strand_filter)
        p.no_strand_balance_filter, p.strand_balance_ratio = int(self.no_strand_balance_filter), self.strand_balance_ratio
        p.n_per_strand, p.strand_odds_max, p.min_depth = self.n_per_strand, self.strand_odds_max, self.min_depth
        p.min_variant_depth, p.variant_multiplier, p.table_log2 = self.min_variant_depth, self.variant_multiplier, self.table_log2
        return p


def pack_reads(bases, read_off):
    """bk_reads_pack: (packed u32 words, packed offsets u32, rest bases u8, rest offsets u32) — the packable reads of a
    chunk as 2-bit words and the others as ASCII (for Bronko.push_packed_ptr / push)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    read_off = np.ascontiguousarray(read_off, dtype=np.uint32)
    n = len(read_off) - 1
    nb = int(read_off[-1]) if n else 0
    packed = np.zeros(nb // 16 + 2, dtype=np.uint32)
    poff = np.zeros(n + 1, dtype=np.uint32)
    rest = np.full(nb + 64, ord("*"), dtype=np.uint8)
    roff = np.zeros(n + 1, dtype=np.uint32)
    npk, nrs = C.c_uint64(), C.c_uint64()
    rc = L.lib().bk_reads_pack(L.ptr(bases), L.ptr(read_off), n, L.ptr(packed), L.ptr(poff), C.byref(npk), L.ptr(rest), L.ptr(roff), C.byref(nrs))
    if rc != 0:
        raise BkError(rc, "bk_reads_pack")
    return packed, poff[:npk.value + 1].copy(), rest, roff[:nrs.value + 1].copy()


def clean_sample_id(path):
    buf = C.create_string_buffer(4096)
    L.lib().bk_clean_sample_id(path.encode(), buf, 4096)
    return buf.value.decode()


class DecodedReads:
    """One FASTQ(.gz) file decoded on the host (bk_fastq_decode: no context, no GPU, thread-safe — ctypes drops the
    GIL, so files of the next samples can be decoded on other threads while the GPU works)."""

    def __init__(self, path):
        self._lib = L.lib()
        h, err = C.c_void_p(), C.create_string_buffer(512)
        rc = self._lib.bk_fastq_decode(path.encode(), C.byref(h), err, 512)
        if rc != 0:
            raise BkError(rc, err.value.decode())
        self.h = h

    def chunks(self):
        """[(bases uint8 view, offsets uint32 view)] — valid until close()."""
        out = []
        for i in range(self._lib.bk_reads_n_chunks(self.h)):
            b, o, nr, nb = C.c_void_p(), C.c_void_p(), C.c_uint64(), C.c_uint64()
            self._lib.bk_reads_chunk(self.h, i, C.byref(b), C.byref(o), C.byref(nr), C.byref(nb))
            bases = np.ctypeslib.as_array(C.cast(b, C.POINTER(C.c_uint8)), shape=(max(nb.value, 1),))[:nb.value]
            off = np.ctypeslib.as_array(C.cast(o, C.POINTER(C.c_uint32)), shape=(nr.value + 1,))
            out.append((bases, off))
        return out

    def close(self):
        if getattr(self, "h", None):
            self._lib.bk_reads_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Bronko:
    """One GPU context (bk_ctx).  Samples are processed sequentially per context."""

    def __init__(self, device=0):
        self._lib = L.lib()
        h = C.c_void_p()
        rc = self._lib.bk_create(C.byref(h), device)
        if rc != 0:
            raise BkError(rc, self._lib.bk_last_error(None).decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self._lib.bk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise BkError(rc, self._lib.bk_last_error(self.h).decode())

    @property
    def stream(self):
        return self._lib.bk_stream(self.h)

    # ---- index --------------------------------------------------------------------------------
    def load_index(self, bkdb_path):
        self._check(self._lib.bk_index_load_file(self.h, bkdb_path.encode()))

    def build_index(self, k, fasta_paths):
        arr = (C.c_char_p * len(fasta_paths))(*[p.encode() for p in fasta_paths])
        self._check(self._lib.bk_index_build(self.h, k, len(fasta_paths), arr))

    def save_index(self, path):
        self._check(self._lib.bk_index_save(self.h, path.encode()))

    def share_index(self, owner):
        """Read the index of another context on the same GPU instead of loading a copy (one context per sample in
        flight, one set of tables)."""
        self._check(self._lib.bk_index_share(self.h, owner.h))

    def load_index_arrays(self, k, keys, entry_off, entries, genome_seq_off, seq_len, seq_base_off, ref_bases):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        entry_off = np.ascontiguousarray(entry_off, dtype=np.uint64)
        entries = np.ascontiguousarray(entries, dtype=L.BUCKETINFO_DTYPE)
        genome_seq_off = np.ascontiguousarray(genome_seq_off, dtype=np.uint32)
        seq_len = np.ascontiguousarray(seq_len, dtype=np.uint64)
        seq_base_off = np.ascontiguousarray(seq_base_off, dtype=np.uint64)
        ref_bases = np.ascontiguousarray(ref_bases, dtype=np.uint8)
        self._check(self._lib.bk_index_load(self.h, k, len(keys), L.ptr(keys), L.ptr(entry_off), L.ptr(entries),
                                            len(genome_seq_off) - 1, L.ptr(genome_seq_off), L.ptr(seq_len),
                                            L.ptr(seq_base_off), L.ptr(ref_bases)))

    def index_info(self):
        k, nk, ne, ng = C.c_uint32(), C.c_uint64(), C.c_uint64(), C.c_uint32()
        self._check(self._lib.bk_index_info(self.h, C.byref(k), C.byref(nk), C.byref(ne), C.byref(ng)))
        return {"k": k.value, "n_keys": nk.value, "n_entries": ne.value, "n_genomes": ng.value}

    def index_export(self):
        info = self.index_info()
        keys = np.zeros(info["n_keys"], dtype=np.uint64)
        off = np.zeros(info["n_keys"] + 1, dtype=np.uint64)
        ent = np.zeros(info["n_entries"], dtype=L.BUCKETINFO_DTYPE)
        self._check(self._lib.bk_index_export(self.h, L.ptr(keys), L.ptr(off), L.ptr(ent)))
        return keys, off, ent

    def genomes(self):
        out = []
        for g in range(self.index_info()["n_genomes"]):
            seqs = []
            for s in range(self._lib.bk_genome_n_seqs(self.h, g)):
                n = self._lib.bk_seq_len(self.h, g, s)
                seqs.append((self._lib.bk_seq_name(self.h, g, s).decode(), n,
                             C.string_at(self._lib.bk_seq_bases(self.h, g, s), n)))
            out.append((self._lib.bk_genome_name(self.h, g).decode(), seqs))
        return out

    # ---- one sample ---------------------------------------------------------------------------
    def begin(self, args: CallArgs = None):
        self._args = args or CallArgs()
        p = self._args.to_params()
        self._check(self._lib.bk_sample_begin(self.h, C.byref(p)))

    def push(self, file_slot, bases, read_off):
        """Host buffers (numpy uint8 bases, uint32 offsets with n_reads+1 entries)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        read_off = np.ascontiguousarray(read_off, dtype=np.uint32)
        self._check(self._lib.bk_reads_push(self.h, file_slot, L.ptr(bases), L.ptr(read_off), len(read_off) - 1))

    def push_ptr(self, file_slot, bases_ptr, off_ptr, n_reads):
        self._check(self._lib.bk_reads_push(self.h, file_slot, bases_ptr, off_ptr, n_reads))

    def push_device(self, file_slot, d_bases_ptr, d_off_ptr, n_reads, n_bases, max_read_len):
        self._check(self._lib.bk_reads_push_device(self.h, file_slot, d_bases_ptr, d_off_ptr, n_reads, n_bases, max_read_len))

    def push_fastq(self, file_slot, path):
        self._check(self._lib.bk_reads_push_fastq(self.h, file_slot, path.encode()))

    def push_packed_ptr(self, file_slot, packed_ptr, off_ptr, n_reads):
        self._check(self._lib.bk_reads_push_packed(self.h, file_slot, packed_ptr, off_ptr, n_reads))

    def push_fastq_mem(self, file_slot, data):
        """The bytes of a FASTQ(.gz) file (bytes / numpy uint8): inflated (BGZF: by the GPU's decompression engine) and
        parsed on the device."""
        a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, dtype=np.uint8)
        self._check(self._lib.bk_reads_push_fastq_mem(self.h, file_slot, L.ptr(a) if len(a) else None, len(a)))

    def decode_info(self, file_slot=0):
        d = L.DecodeInfo()
        self._check(self._lib.bk_decode_info_get(self.h, file_slot, C.byref(d)))
        names = {0: "none", 1: "plain text, parsed on the device", 2: "gzip inflated by zlib on the host, parsed on the device",
                 3: "BGZF inflated by the GPU decompression engine, parsed on the device"}
        return {"mode": d.mode, "how": names.get(d.mode, "?"), "segments": d.segments, "compressed_bytes": d.compressed_bytes,
                "text_bytes": d.text_bytes, "n_reads": d.n_reads, "n_bases": d.n_bases}

    def push_decoded(self, file_slot, reads: DecodedReads):
        self._check(self._lib.bk_reads_push_decoded(self.h, file_slot, reads.h))

    def finish(self):
        res = L.SampleResult()
        self._check(self._lib.bk_sample_finish(self.h, C.byref(res)))
        return Sample(self, res)

    def call_sample(self, files, args: CallArgs = None):
        """files: [(bases, offsets)] (single-end) or [(r1 bases, r1 off), (r2 bases, r2 off)] — the body of
        the loops at reference src/call.rs:213-292 / 298-387."""
        self.begin(args)
        for slot, (bases, off) in enumerate(files):
            self.push(slot, bases, off)
        return self.finish()

    def set_stage_timing(self, on: bool):
        """Per-stage CUDA events on / off (bk_stage_timing): off saves ~20 driver calls per sample."""
        self._check(self._lib.bk_stage_timing(self.h, 1 if on else 0))

    def stage_times(self):
        t = L.StageTimes()
        self._check(self._lib.bk_stage_times_get(self.h, C.byref(t)))
        return {f: getattr(t, f) for f, _ in L.StageTimes._fields_}


class Sample:
    """Results of one finished sample (valid until the next begin() on the same context)."""

    def __init__(self, ctx: Bronko, res):
        self.ctx, self.res = ctx, res
        self.best_genome = res.best_genome
        self.n_files = res.n_files
        self.num_major_variants = res.num_major_variants
        self.num_minor_variants = res.num_minor_variants
        self.breadth_coverage = res.breadth_coverage
        self.depth_coverage = res.depth_coverage
        self.num_perfect_kmers = res.num_perfect_kmers
        self.num_variant_kmers = res.num_variant_kmers
        self.num_unmapped_kmers = res.num_unmapped_kmers
        v = np.zeros(res.n_variants, dtype=L.VARIANT_DTYPE)
        if res.n_variants:
            ctx._check(ctx._lib.bk_sample_variants(ctx.h, L.ptr(v), res.n_variants))
        self.variants = v

    def kmc_stats(self, file=0):
        s = self.res.kmc[file]
        return (s.total_reads, s.total_kmers, s.unique_kmers, s.unique_counted)

    def kmers(self, file=0):
        """(k-mers u64 ascending, counts u32) — what load_kmers returns (reference src/call.rs:1241-1255)."""
        n = C.c_uint64(0)
        self.ctx._check(self.ctx._lib.bk_kmer_counts_get(self.ctx.h, file, None, None, C.byref(n)))
        km = np.zeros(n.value, dtype=np.uint64)
        ct = np.zeros(n.value, dtype=np.uint32)
        self.ctx._check(self.ctx._lib.bk_kmer_counts_get(self.ctx.h, file, L.ptr(km), L.ptr(ct), C.byref(n)))
        return km, ct

    def mapping_data(self, file=0):
        ng = self.ctx.index_info()["n_genomes"]
        o = np.zeros(ng, dtype=L.GENOME_STATS_DTYPE)
        self.ctx._check(self.ctx._lib.bk_sample_genome_stats(self.ctx.h, file, L.ptr(o)))
        return o

    def pileup(self):
        """(4, rows, 4) u64: fwd depth, rev depth, fwd support, rev support of the selected genome."""
        g = self.ctx.genomes()[self.best_genome]
        rows = sum(n for _, n, _ in g[1])
        out = np.zeros((4, rows, 4), dtype=np.uint64)
        for a in range(4):
            self.ctx._check(self.ctx._lib.bk_sample_pileup(self.ctx.h, a, L.ptr(out[a]), rows))
        return out

    def noise_max(self):
        g = self.ctx.genomes()[self.best_genome]
        rows = sum(n for _, n, _ in g[1])
        out = np.zeros(rows, dtype=np.float64)
        self.ctx._check(self.ctx._lib.bk_sample_noise_max(self.ctx.h, L.ptr(out), rows))
        return out

    def write_vcf(self, reads_path, out_path):
        self.ctx._check(self.ctx._lib.bk_write_vcf(self.ctx.h, reads_path.encode(), out_path.encode()))

    def write_pileup(self, out_path):
        self.ctx._check(self.ctx._lib.bk_write_pileup(self.ctx.h, out_path.encode()))
