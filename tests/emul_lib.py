"""ctypes loader of tests/emul/libbk_emul.so — the product's HOST code (index builder, .bkdb reader/writer,
tau table, sample-id) plus a CPU stepping of the counting-stage logic (bk_core.cuh).  Tests only."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "emul", "libbk_emul.so")
P, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", os.path.join(HERE, "emul"), "-s"])
        L = C.CDLL(SO)
        sig = {
            "emul_create_bkdb": (P, [C.c_char_p]), "emul_create_fasta": (P, [u32, u32, P]), "emul_free": (None, [P]),
            "emul_n_keys": (u64, [P]), "emul_n_entries": (u64, [P]), "emul_export": (None, [P, P, P, P]),
            "emul_save": (C.c_int, [P, C.c_char_p]), "emul_assign_buckets": (None, [u64, C.c_int, P]),
            "emul_revcomp": (u64, [u64, C.c_int]), "emul_tau_table": (None, [P]),
            "emul_clean_sample_id": (u64, [C.c_char_p, C.c_char_p, u64]),
            "emul_count": (u64, [P, P, P, u64, u32, u32, u32, u32, P, P]), "emul_count_get": (None, [P, P, P]),
            "emul_noise": (C.c_int, [P, P, u32, P, P]), "emul_hint": (None, [C.c_double, P, u32, P]), "emul_group_check": (u64, [P, P, u64, u32, u32]),
        }
        for n, (r, a) in sig.items():
            f = getattr(L, n)
            f.restype, f.argtypes = r, a
        _lib = L
    return _lib


def ptr(a):
    return a.ctypes.data_as(P)


class Emul:
    def __init__(self, h):
        assert h, "emul: NULL handle"
        self.h = h

    @classmethod
    def from_fasta(cls, k, paths):
        arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
        return cls(lib().emul_create_fasta(k, len(paths), arr))

    @classmethod
    def from_bkdb(cls, path):
        return cls(lib().emul_create_bkdb(path.encode()))

    def __del__(self):
        try:
            lib().emul_free(self.h)
        except Exception:
            pass

    def export(self):
        from oracle.oracle import BUCKETINFO_DTYPE
        nk, ne = lib().emul_n_keys(self.h), lib().emul_n_entries(self.h)
        keys, off = np.zeros(nk, dtype=np.uint64), np.zeros(nk + 1, dtype=np.uint64)
        ent = np.zeros(ne, dtype=BUCKETINFO_DTYPE)
        lib().emul_export(self.h, ptr(keys), ptr(off), ptr(ent))
        return keys, off, ent

    def group_check(self, canonical_kmers, b0, b1):
        """Queries the grouped map tables disagree on with the per-bucket table (0 = exact)."""
        q = np.ascontiguousarray(canonical_kmers, dtype=np.uint64)
        return int(lib().emul_group_check(self.h, ptr(q), len(q), b0, b1))

    def save(self, path):
        assert lib().emul_save(self.h, path.encode()) == 0

    def count(self, bases, off, ci=3, cs=1000000, gen_log2=20, desc_cap=None, dense=False):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint32)
        n_reads = len(off) - 1
        st, dbg = np.zeros(4, dtype=np.uint64), np.zeros(3, dtype=np.uint64)
        cap = desc_cap if desc_cap is not None else 2 * n_reads + 64
        n = lib().emul_count(self.h, ptr(bases), ptr(off), n_reads, gen_log2 | (0x80000000 if dense else 0), cap, ci, cs, ptr(st), ptr(dbg))
        assert n < 2 ** 63, "emulation reported a logic error (code %d)" % (2 ** 64 - 1 - n)
        km, ct = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint32)
        lib().emul_count_get(self.h, ptr(km), ptr(ct))
        return km, ct, tuple(int(x) for x in st), tuple(int(x) for x in dbg)


def noise(fwd, rev):
    """Noise.max per position with the device logic of bk_noise.cuh stepped on the CPU.  fwd/rev: (len, 4) depths.
    Returns (max, stats) with stats = (chunks replayed, iterations replayed, chain rounds, chain stops, serial its)."""
    fwd = np.ascontiguousarray(fwd, dtype=np.uint32)
    rev = np.ascontiguousarray(rev, dtype=np.uint32)
    n = fwd.shape[0]
    out, st = np.zeros(n, dtype=np.float64), np.zeros(5, dtype=np.uint32)
    lib().emul_noise(ptr(fwd), ptr(rev), n, ptr(out), ptr(st))
    return out, tuple(int(x) for x in st)


def hint(s, ahead):
    """(iterations ahead that fit three zones, exponent field of the lowest zone) — the look-ahead of nz_chain_block."""
    a = np.ascontiguousarray(ahead, dtype=np.float64)
    out = np.zeros(2, np.uint32)
    lib().emul_hint(float(s), ptr(a), len(a), ptr(out))
    return int(out[0]), int(out[1])
