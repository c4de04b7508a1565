"""ctypes binding of libbronko_b200.so (include/bronko_b200.h).  No fallback: if the shared library
or an sm_100 device is missing, loading / context creation raises."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
# BRONKO_B200_LIB: another build of the same library (tools/build_variant.sh: A/B runs of kernel variants)
SO_PATH = os.environ.get("BRONKO_B200_LIB") or os.path.join(CSRC, "libbronko_b200.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "bronko_b200.h")

u8p, u32, u64, P = C.POINTER(C.c_uint8), C.c_uint32, C.c_uint64, C.c_void_p


class BkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libbronko_b200 error %d: %s" % (code, msg))
        self.code = code


class Params(C.Structure):
    _fields_ = [("k", u32), ("min_kmers", u32), ("counter_max", u32), ("use_full_kmer", u32), ("n_fixed", u32),
                ("no_end_filter", u32), ("no_strand_filter", u32), ("no_strand_balance_filter", u32),
                ("n_per_strand", u32), ("table_log2", u32), ("min_depth", u64), ("min_variant_depth", u64),
                ("min_af", C.c_double), ("strand_balance_ratio", C.c_double), ("strand_odds_max", C.c_double),
                ("variant_multiplier", C.c_double)]


class KmcStats(C.Structure):
    _fields_ = [("total_reads", u64), ("total_kmers", u64), ("unique_kmers", u64), ("unique_counted", u64)]


class SampleResult(C.Structure):
    _fields_ = [("best_genome", C.c_int32), ("n_files", u32), ("n_variants", u64), ("num_major_variants", u64),
                ("num_minor_variants", u64), ("breadth_coverage", C.c_double), ("depth_coverage", C.c_double),
                ("num_perfect_kmers", u64), ("num_variant_kmers", u64), ("num_unmapped_kmers", u64),
                ("kmc", KmcStats * 2)]


class StageTimes(C.Structure):
    _fields_ = [("scan_ms", C.c_float), ("leftover_ms", C.c_float), ("finalize_ms", C.c_float), ("map_ms", C.c_float),
                ("score_ms", C.c_float), ("total_ms", C.c_float), ("launches", u32), ("scan_launches", u32),
                ("coll_ms", C.c_float), ("coll_calls", u32), ("decode_ms", C.c_float)]


class DecodeInfo(C.Structure):
    _fields_ = [("mode", u32), ("segments", u32), ("compressed_bytes", u64), ("text_bytes", u64), ("n_reads", u64), ("n_bases", u64)]


VARIANT_DTYPE = np.dtype([("seq", "<u4"), ("pos", "<u4"), ("ref_base", "u1"), ("alt_base", "u1"), ("pad", "u1", 6),
                          ("fwd_ref", "<u8"), ("rev_ref", "<u8"), ("fwd_alt", "<u8"), ("rev_alt", "<u8"),
                          ("depth", "<u8"), ("af", "<f8"), ("sor", "<f8")])
GENOME_STATS_DTYPE = np.dtype([("perfect", "<u8"), ("variant", "<u8"), ("unique_perfect", "<u8"),
                               ("present", "<u4"), ("pad", "<u4")])
BUCKETINFO_DTYPE = np.dtype([("file_id", "<u2"), ("seq_id", "u1"), ("pad0", "u1"), ("location", "<u4"),
                             ("idx", "u1"), ("canonical", "u1"), ("pad1", "u1", 2)])

# every symbol include/bronko_b200.h declares: name → (restype, argtypes)
SIGNATURES = {
    "bk_create": (C.c_int, [C.POINTER(P), C.c_int]),
    "bk_destroy": (None, [P]),
    "bk_last_error": (C.c_char_p, [P]),
    "bk_stream": (P, [P]),
    "bk_stream_slot": (P, [P, C.c_int]),
    "bk_version": (C.c_char_p, []),
    "bk_index_load": (C.c_int, [P, u32, u64, P, P, P, u32, P, P, P, P]),
    "bk_index_load_file": (C.c_int, [P, C.c_char_p]),
    "bk_index_build": (C.c_int, [P, u32, u32, P]),
    "bk_index_save": (C.c_int, [P, C.c_char_p]),
    "bk_index_share": (C.c_int, [P, P]),
    "bk_index_info": (C.c_int, [P, P, P, P, P]),
    "bk_genome_name": (C.c_char_p, [P, u32]),
    "bk_genome_n_seqs": (u32, [P, u32]),
    "bk_seq_name": (C.c_char_p, [P, u32, u32]),
    "bk_seq_len": (u64, [P, u32, u32]),
    "bk_seq_bases": (P, [P, u32, u32]),
    "bk_index_export": (C.c_int, [P, P, P, P]),
    "bk_params_default": (None, [C.POINTER(Params)]),
    "bk_sample_begin": (C.c_int, [P, C.POINTER(Params)]),
    "bk_reads_push": (C.c_int, [P, C.c_int, P, P, u64]),
    "bk_reads_push_device": (C.c_int, [P, C.c_int, P, P, u64, u64, u32]),
    "bk_reads_push_packed": (C.c_int, [P, C.c_int, P, P, u64]),
    "bk_reads_pack": (C.c_int, [P, P, u64, P, P, P, P, P, P]),
    "bk_reads_push_fastq": (C.c_int, [P, C.c_int, C.c_char_p]),
    "bk_reads_push_fastq_mem": (C.c_int, [P, C.c_int, P, u64]),
    "bk_decode_info_get": (C.c_int, [P, C.c_int, C.POINTER(DecodeInfo)]),
    "bk_fastq_decode": (C.c_int, [C.c_char_p, C.POINTER(P), C.c_char_p, u64]),
    "bk_reads_n_chunks": (u64, [P]),
    "bk_reads_chunk": (C.c_int, [P, u64, C.POINTER(P), C.POINTER(P), C.POINTER(u64), C.POINTER(u64)]),
    "bk_reads_push_decoded": (C.c_int, [P, C.c_int, P]),
    "bk_reads_free": (None, [P]),
    "bk_sample_finish": (C.c_int, [P, C.POINTER(SampleResult)]),
    "bk_sample_result_get": (C.c_int, [P, C.POINTER(SampleResult)]),
    "bk_sample_variants": (C.c_int, [P, P, u64]),
    "bk_sample_genome_stats": (C.c_int, [P, C.c_int, P]),
    "bk_sample_pileup": (C.c_int, [P, C.c_int, P, u64]),
    "bk_sample_noise_max": (C.c_int, [P, P, u64]),
    "bk_kmer_counts_get": (C.c_int, [P, C.c_int, P, P, P]),
    "bk_stage_times_get": (C.c_int, [P, C.POINTER(StageTimes)]),
    "bk_stage_timing": (C.c_int, [P, C.c_int]),
    "bk_write_vcf": (C.c_int, [P, C.c_char_p, C.c_char_p]),
    "bk_write_pileup": (C.c_int, [P, C.c_char_p]),
    "bk_clean_sample_id": (u64, [C.c_char_p, C.c_char_p, u64]),
    "bk_shard_unique_id": (C.c_int, [P]),
    "bk_shard_init": (C.c_int, [P, u32, u32, P]),
    "bk_shard_info": (C.c_int, [P, P, P]),
    "bk_shard_local": (C.c_int, [P, u32]),
    "bk_shard_finish_local": (C.c_int, [P, u32, C.POINTER(SampleResult)]),
    "bk_host_alloc": (P, [u64]),
    "bk_host_free": (None, [P]),
}

_lib = None


def build(force=False):
    """Compile libbronko_b200.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    if force and os.path.exists(SO_PATH):
        os.remove(SO_PATH)
    subprocess.check_call(["make", "-C", CSRC, "-s", "libbronko_b200.so"])
    return SO_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError("%s is missing: run `make -C bronko_b200/csrc` (or __graft_entry__.build()); "
                              "bronko_b200 has no fallback path" % SO_PATH)
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def ptr(a):
    return a.ctypes.data_as(P)
