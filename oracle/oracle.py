"""ctypes loader for the CPU oracle (oracle/bronko_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs.  The product package (bronko_b200/) never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libbronko_oracle.so")

u64 = C.c_uint64
P = C.c_void_p


def build(force=False):
    src = os.path.join(_HERE, "bronko_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class Params(C.Structure):
    """Mirror of the CallArgs fields the hot path consumes (src/cli.rs:61-166, src/consts.rs)."""
    _fields_ = [
        ("k", u64), ("min_kmers", u64), ("use_full_kmer", C.c_int), ("n_fixed", u64),
        ("min_af", C.c_double), ("no_end_filter", C.c_int), ("no_strand_filter", C.c_int),
        ("no_strand_balance_filter", C.c_int), ("strand_balance_ratio", C.c_double),
        ("n_per_strand", u64), ("strand_odds_max", C.c_double), ("min_depth", u64),
        ("min_variant_depth", u64), ("variant_multiplier", C.c_double),
    ]

    @classmethod
    def defaults(cls, k=21, **kw):
        p = cls(k=k, min_kmers=3, use_full_kmer=0, n_fixed=2, min_af=0.03, no_end_filter=0,
                no_strand_filter=0, no_strand_balance_filter=0, strand_balance_ratio=0.1,
                n_per_strand=2, strand_odds_max=6.0, min_depth=300, min_variant_depth=3,
                variant_multiplier=1.5)
        for a, b in kw.items():
            setattr(p, a, b)
        return p


VCF_DTYPE = np.dtype([("seq", "<u4"), ("pos", "<u4"), ("ref_base", "u1"), ("alt_base", "u1"),
                      ("pad", "u1", 6), ("fwd_ref", "<u8"), ("rev_ref", "<u8"), ("fwd_alt", "<u8"),
                      ("rev_alt", "<u8"), ("depth", "<u8"), ("af", "<f8"), ("sor", "<f8")])
BUCKETINFO_DTYPE = np.dtype([("file_id", "<u2"), ("seq_id", "u1"), ("pad0", "u1"), ("location", "<u4"),
                             ("idx", "u1"), ("canonical", "u1"), ("pad1", "u1", 2)])
assert VCF_DTYPE.itemsize == 72 and BUCKETINFO_DTYPE.itemsize == 12

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    sig = {
        "orc_assign_buckets": (None, [u64, C.c_int, P]),
        "orc_canonical_kmer": (u64, [C.c_char_p, C.c_int, P]),
        "orc_reverse_complement": (u64, [u64, C.c_int]),
        "orc_students_t_inverse_cdf": (C.c_double, [C.c_double, C.c_double]),
        "orc_thompson_tau": (C.c_double, [u64]),
        "orc_index_build": (P, [C.c_int, P]),
        "orc_index_load": (P, [C.c_char_p, P, P]),
        "orc_index_decode": (P, [P, u64, P]),
        "orc_index_save": (C.c_int, [P, C.c_char_p]),
        "orc_index_free": (None, [P]),
        "orc_index_k": (u64, [P]), "orc_index_meta_k": (u64, [P]),
        "orc_index_n_keys": (u64, [P]), "orc_index_n_entries": (u64, [P]),
        "orc_index_export": (None, [P, P, P, P]),
        "orc_index_n_genomes": (u64, [P]),
        "orc_genome_name": (C.c_char_p, [P, u64]), "orc_genome_n_seqs": (u64, [P, u64]),
        "orc_seq_name": (C.c_char_p, [P, u64, u64]), "orc_seq_len": (u64, [P, u64, u64]),
        "orc_seq_nbases": (u64, [P, u64, u64]), "orc_seq_bases": (P, [P, u64, u64]),
        "orc_fastq_read": (P, [C.c_char_p]), "orc_reads_n": (u64, [P]), "orc_reads_nbases": (u64, [P]),
        "orc_reads_bases": (P, [P]), "orc_reads_off": (P, [P]), "orc_reads_free": (None, [P]),
        "orc_count": (P, [C.c_int, P, P, u64, u64, u64, C.c_int]),
        "orc_counts_n": (u64, [P]), "orc_counts_get": (None, [P, P, P]),
        "orc_counts_stats": (None, [P, P]), "orc_counts_from_list": (P, [P, P, u64]),
        "orc_counts_free": (None, [P]),
        "orc_sample_run": (P, [P, P, C.c_int, P]), "orc_sample_free": (None, [P]),
        "orc_sample_map_only": (P, [P, P, C.c_int, P]), "orc_pick_best": (C.c_int, [P]),
        "orc_sample_call_with": (None, [P, P, C.c_int, P, P, P, u64, u64]),
        "orc_sample_best": (C.c_int, [P]), "orc_sample_stats": (None, [P, C.c_int, P]),
        "orc_sample_genome_rows": (u64, [P, C.c_int]), "orc_sample_pileup": (None, [P, C.c_int, C.c_int, P]),
        "orc_sample_n_variants": (u64, [P]), "orc_sample_variants": (None, [P, P]),
        "orc_sample_summary": (None, [P, P, P, P, P]), "orc_sample_noise_max": (None, [P, P]),
        "orc_sample_unmapped": (u64, [P]),
        "orc_write_vcf": (C.c_int, [P, C.c_char_p, C.c_char_p]), "orc_write_pileup": (C.c_int, [P, C.c_char_p]),
        "orc_vcf_text": (u64, [P, C.c_char_p, C.c_char_p, u64]), "orc_pileup_text": (u64, [P, C.c_char_p, u64]),
        "orc_clean_sample_id": (u64, [C.c_char_p, C.c_char_p, u64]),
        "orc_baseline_noise": (C.c_int, [P, P, u64, P, P, P]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(P)


def assign_buckets(kmer, k):
    out = np.zeros(k, dtype=np.uint64)
    lib().orc_assign_buckets(int(kmer), k, _ptr(out))
    return [int(x) for x in out]


def canonical_kmer(kmer: str, k=None):
    rc = C.c_int(0)
    v = lib().orc_canonical_kmer(kmer.encode(), k or len(kmer), C.byref(rc))
    return int(v), bool(rc.value)


def thompson_tau(n):
    return lib().orc_thompson_tau(n)


def students_t_inverse_cdf(df, x):
    return lib().orc_students_t_inverse_cdf(df, x)


def clean_sample_id(path):
    buf = C.create_string_buffer(4096)
    lib().orc_clean_sample_id(path.encode(), buf, 4096)
    return buf.value.decode()


class Index:
    def __init__(self, h):
        if not h:
            raise RuntimeError("oracle: index handle is NULL")
        self.h = h

    @classmethod
    def build(cls, k, paths):
        arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
        return cls(lib().orc_index_build(k, len(paths), arr))

    @classmethod
    def load(cls, path):
        consumed, size = u64(0), u64(0)
        ix = cls(lib().orc_index_load(path.encode(), C.byref(consumed), C.byref(size)))
        ix.consumed, ix.file_size = consumed.value, size.value
        return ix

    @classmethod
    def decode(cls, data: bytes):
        consumed = u64(0)
        buf = np.frombuffer(data, dtype=np.uint8)
        ix = cls(lib().orc_index_decode(_ptr(buf), len(data), C.byref(consumed)))
        ix.consumed, ix.file_size = consumed.value, len(data)
        return ix

    def save(self, path):
        if lib().orc_index_save(self.h, path.encode()) != 0:
            raise IOError(path)

    def __del__(self):
        try:
            lib().orc_index_free(self.h)
        except Exception:
            pass

    @property
    def k(self):
        return lib().orc_index_k(self.h)

    @property
    def meta_k(self):
        return lib().orc_index_meta_k(self.h)

    @property
    def n_keys(self):
        return lib().orc_index_n_keys(self.h)

    @property
    def n_entries(self):
        return lib().orc_index_n_entries(self.h)

    def export(self):
        """(keys ascending u64[n], entry_off u64[n+1], entries BUCKETINFO_DTYPE[m])"""
        keys = np.zeros(self.n_keys, dtype=np.uint64)
        off = np.zeros(self.n_keys + 1, dtype=np.uint64)
        ent = np.zeros(self.n_entries, dtype=BUCKETINFO_DTYPE)
        lib().orc_index_export(self.h, _ptr(keys), _ptr(off), _ptr(ent))
        return keys, off, ent

    def genomes(self):
        L = lib()
        out = []
        for g in range(L.orc_index_n_genomes(self.h)):
            seqs = []
            for s in range(L.orc_genome_n_seqs(self.h, g)):
                n = L.orc_seq_nbases(self.h, g, s)
                raw = C.string_at(L.orc_seq_bases(self.h, g, s), n)
                seqs.append((L.orc_seq_name(self.h, g, s).decode(), L.orc_seq_len(self.h, g, s), raw))
            out.append((L.orc_genome_name(self.h, g).decode(), seqs))
        return out


class Counts:
    def __init__(self, h):
        self.h = h

    @classmethod
    def count(cls, k, bases: np.ndarray, off: np.ndarray, ci=3, cs=1000000, threads=1):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        return cls(lib().orc_count(k, _ptr(bases), _ptr(off), len(off) - 1, ci, cs, threads))

    @classmethod
    def from_list(cls, kmers, counts):
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        return cls(lib().orc_counts_from_list(_ptr(kmers), _ptr(counts), len(kmers)))

    def __del__(self):
        try:
            lib().orc_counts_free(self.h)
        except Exception:
            pass

    def get(self):
        n = lib().orc_counts_n(self.h)
        km = np.zeros(n, dtype=np.uint64)
        ct = np.zeros(n, dtype=np.uint64)
        lib().orc_counts_get(self.h, _ptr(km), _ptr(ct))
        return km, ct

    def stats(self):
        """(total_reads, total_kmers, unique_kmers, unique_counted) — the four KMC stdout numbers."""
        o = np.zeros(4, dtype=np.uint64)
        lib().orc_counts_stats(self.h, _ptr(o))
        return tuple(int(x) for x in o)


class Sample:
    """initialize_output_maps → map_kmers per file → pick_best_genome(_paired) → call_variants."""

    def __init__(self, index: Index, params: Params, counts, map_only=False):
        self.index = index
        self.n_files = len(counts)
        arr = (P * len(counts))(*[c.h for c in counts])
        self._keep = (counts, params)
        self.params = params
        run = lib().orc_sample_map_only if map_only else lib().orc_sample_run
        self.h = run(index.h, C.byref(params), len(counts), arr)

    def call_with(self, best, arrays, stats, unique_counted):
        """Replace genome `best`'s pileups / the tallies with externally combined ones, then call_variants."""
        arrays = np.ascontiguousarray(arrays, dtype=np.uint64)
        st = [np.ascontiguousarray(x, dtype=np.uint64) for x in stats] + [np.zeros(1, dtype=np.uint64)]
        uc = list(unique_counted) + [0]
        lib().orc_sample_call_with(self.h, C.byref(self.params), best, _ptr(arrays), _ptr(st[0]), _ptr(st[1]), uc[0], uc[1])

    def __del__(self):
        try:
            lib().orc_sample_free(self.h)
        except Exception:
            pass

    @property
    def best(self):
        return lib().orc_sample_best(self.h)

    def stats(self, file=0):
        ng = lib().orc_index_n_genomes(self.index.h)
        o = np.zeros(ng * 4, dtype=np.uint64)
        lib().orc_sample_stats(self.h, file, _ptr(o))
        return o.reshape(ng, 4)

    def pileup(self, g=None):
        """(4, rows, 4) u64: fwd depth, rev depth, fwd support, rev support of genome g (default best)."""
        g = self.best if g is None else g
        rows = lib().orc_sample_genome_rows(self.h, g)
        out = np.zeros((4, rows, 4), dtype=np.uint64)
        for a in range(4):
            lib().orc_sample_pileup(self.h, g, a, _ptr(out[a]))
        return out

    def variants(self):
        n = lib().orc_sample_n_variants(self.h)
        v = np.zeros(n, dtype=VCF_DTYPE)
        if n:
            lib().orc_sample_variants(self.h, _ptr(v))
        return v

    def summary(self):
        a, b, c, d = u64(0), u64(0), C.c_double(0), C.c_double(0)
        lib().orc_sample_summary(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        return a.value, b.value, c.value, d.value

    def noise_max(self):
        rows = lib().orc_sample_genome_rows(self.h, self.best)
        o = np.zeros(rows, dtype=np.float64)
        lib().orc_sample_noise_max(self.h, _ptr(o))
        return o

    def unmapped(self):
        return lib().orc_sample_unmapped(self.h)

    def vcf_text(self, reads_path):
        n = lib().orc_vcf_text(self.h, reads_path.encode(), None, 0)
        buf = C.create_string_buffer(n)
        lib().orc_vcf_text(self.h, reads_path.encode(), buf, n)
        return buf.value.decode()

    def pileup_text(self):
        n = lib().orc_pileup_text(self.h, None, 0)
        buf = C.create_string_buffer(n)
        lib().orc_pileup_text(self.h, buf, n)
        return buf.value.decode()


def baseline_noise(fwd: np.ndarray, rev: np.ndarray):
    fwd = np.ascontiguousarray(fwd, dtype=np.uint64)
    rev = np.ascontiguousarray(rev, dtype=np.uint64)
    n = fwd.shape[0]
    mx, mean, sd = (np.zeros(n) for _ in range(3))
    rc = lib().orc_baseline_noise(_ptr(fwd), _ptr(rev), n, _ptr(mx), _ptr(mean), _ptr(sd))
    if rc != 0:
        raise ValueError("sequence shorter than the noise window (the reference panics here)")
    return mx, mean, sd


def read_fastq(path):
    h = lib().orc_fastq_read(path.encode())
    if not h:
        raise IOError(path)
    n, nb = lib().orc_reads_n(h), lib().orc_reads_nbases(h)
    bases = np.frombuffer(C.string_at(lib().orc_reads_bases(h), nb), dtype=np.uint8).copy()
    off = np.frombuffer(C.string_at(lib().orc_reads_off(h), (n + 1) * 8), dtype=np.uint64).copy()
    lib().orc_reads_free(h)
    return bases, off
