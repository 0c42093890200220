"""ctypes binding of oracle/libfkm_oracle.so — TEST INFRASTRUCTURE (see oracle/fkm_oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  The product package never does.
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
SO = os.path.join(ORACLE_DIR, "libfkm_oracle.so")


def build(force=False):
    src = os.path.join(ORACLE_DIR, "fkm_oracle.cpp")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return SO


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        u8p = C.POINTER(C.c_uint8)
        lib.fkmo_hash_to_bucket.restype = C.c_int32
        lib.fkmo_hash_to_bucket.argtypes = [C.c_int32, C.c_int32]
        lib.fkmo_is_allowed.restype = C.c_int32
        lib.fkmo_is_allowed.argtypes = [C.c_int32, C.c_int32]
        lib.fkmo_reverse_complement.restype = C.c_int64
        lib.fkmo_reverse_complement.argtypes = [C.c_int64, C.c_int32]
        lib.fkmo_fill_norm.argtypes = [C.c_int32, C.c_void_p]
        lib.fkmo_signature.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        lib.fkmo_orientation.restype = C.c_int32
        lib.fkmo_orientation.argtypes = [C.c_char_p, C.c_int32]
        lib.fkmo_superkmers.restype = C.c_int32
        lib.fkmo_superkmers.argtypes = [C.c_char_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]
        lib.fkmo_window_bins.argtypes = [C.c_char_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        lib.fkmo_count.restype = C.c_void_p
        lib.fkmo_count.argtypes = [C.c_void_p, C.c_uint64] + [C.c_int32] * 7
        lib.fkmo_result_size.restype = C.c_uint64
        lib.fkmo_result_size.argtypes = [C.c_void_p]
        lib.fkmo_result_copy.argtypes = [C.c_void_p] * 5
        lib.fkmo_result_stats.argtypes = [C.c_void_p, C.c_void_p]
        lib.fkmo_result_free.argtypes = [C.c_void_p]
        lib.fkmo_gen_lcg_fasta.restype = C.c_uint64
        lib.fkmo_gen_lcg_fasta.argtypes = [C.c_uint64] * 4 + [C.c_void_p, C.c_uint64]
        lib.fkmo_synth_fasta.restype = C.c_uint64
        lib.fkmo_synth_fasta.argtypes = [C.c_uint64] * 7 + [C.c_void_p, C.c_uint64, C.c_int32]
        lib.fkmo_synth_long_fasta.restype = C.c_uint64
        lib.fkmo_synth_long_fasta.argtypes = [C.c_uint64] * 5 + [C.c_void_p, C.c_uint64, C.c_int32]

    def hash_to_bucket(self, s, B):
        return self.lib.fkmo_hash_to_bucket(s, B)

    def fill_norm(self, m):
        out = np.empty(4 ** m, dtype=np.int32)
        self.lib.fkmo_fill_norm(m, out.ctypes.data)
        return out

    def signature(self, window: bytes, m):
        s, p = C.c_int32(), C.c_int32()
        self.lib.fkmo_signature(window, len(window), m, C.byref(s), C.byref(p))
        return s.value, p.value

    def superkmers(self, rec: bytes, k, m, max_b):
        cap = max(16, len(rec))
        bins = np.empty(cap, dtype=np.int32)
        lens = np.empty(cap, dtype=np.int32)
        n = self.lib.fkmo_superkmers(rec, len(rec), k, m, max_b, bins.ctypes.data, lens.ctypes.data, cap)
        return list(zip(bins[:n].tolist(), lens[:n].tolist()))

    def window_bins(self, rec: bytes, k, m, max_b):
        out = np.full(max(len(rec) - k + 1, 0), -2, dtype=np.int32)
        if out.size:
            self.lib.fkmo_window_bins(rec, len(rec), k, m, max_b, out.ctypes.data)
        return out

    def gen_lcg_fasta(self, seed, G, R, L) -> bytes:
        n = self.lib.fkmo_gen_lcg_fasta(seed, G, R, L, None, 0)
        buf = np.empty(n, dtype=np.uint8)
        self.lib.fkmo_gen_lcg_fasta(seed, G, R, L, buf.ctypes.data, n)
        return buf.tobytes()

    def synth_fasta(self, spec, threads=8) -> np.ndarray:
        """SURVEY §8(d) synthetic reads as FASTA text; spec as for fastkmer_b200.synth_fasta (the oracle's own generator)."""
        a = list(spec["seeds"]) + [spec["genome_len"], spec["n_reads"], spec["read_len"], spec.get("first_read", 0)]
        n = self.lib.fkmo_synth_fasta(*a, None, 0, threads)
        buf = np.empty(n, dtype=np.uint8)
        assert self.lib.fkmo_synth_fasta(*a, buf.ctypes.data, n, threads) == n
        return buf

    def synth_long_fasta(self, spec, threads=8) -> np.ndarray:
        a = list(spec["seeds"]) + [spec.get("first_pos", 0), spec["n_bases"]]
        n = self.lib.fkmo_synth_long_fasta(*a, None, 0, threads)
        buf = np.empty(n, dtype=np.uint8)
        assert self.lib.fkmo_synth_long_fasta(*a, buf.ctypes.data, n, threads) == n
        return buf

    def count(self, fasta, k, m, x, max_b, use_ht, threads=4, sorted_=True):
        """-> dict(bin, hi, lo, cnt numpy arrays sorted by (bin, key); stats dict)."""
        if isinstance(fasta, (bytes, bytearray)):
            arr = np.frombuffer(fasta, dtype=np.uint8)
        else:
            arr = np.ascontiguousarray(fasta, dtype=np.uint8)
        h = self.lib.fkmo_count(arr.ctypes.data, arr.size, k, m, x, max_b, int(use_ht), threads, int(sorted_))
        if not h:
            raise ValueError("oracle rejected the configuration")
        try:
            n = self.lib.fkmo_result_size(h)
            bin_ = np.empty(n, dtype=np.int32)
            hi = np.empty(n, dtype=np.uint64)
            lo = np.empty(n, dtype=np.uint64)
            cnt = np.empty(n, dtype=np.uint32)
            self.lib.fkmo_result_copy(h, bin_.ctypes.data, hi.ctypes.data, lo.ctypes.data, cnt.ctypes.data)
            st = np.zeros(12, dtype=np.uint64)
            self.lib.fkmo_result_stats(h, st.ctypes.data)
        finally:
            self.lib.fkmo_result_free(h)
        names = ["n_bases", "n_kmers", "n_superkmers", "superkmer_bases", "n_records", "n_distinct",
                 "total_count", "digest_sum", "digest_xor", "us_map", "us_reduce"]
        stats = {k_: int(v) for k_, v in zip(names, st.tolist())}
        return {"bin": bin_, "hi": hi, "lo": lo, "cnt": cnt, "stats": stats}


def load():
    build()
    return Oracle(C.CDLL(SO))


def kmer_str(hi, lo, k):
    v = (int(hi) << 64) | int(lo)
    return "".join("ACGT"[(v >> (2 * (k - 1 - i))) & 3] for i in range(k))


def lines_of(res, k):
    """'<bin>\\t<kmer>\\t<count>\\n' lines (SURVEY App. C.4 digest form)."""
    return [b"%d\t%s\t%d\n" % (int(b), kmer_str(h, l, k).encode(), int(c))
            for b, h, l, c in zip(res["bin"], res["hi"], res["lo"], res["cnt"])]


def digest_of(res, k):
    return hashlib.sha256(b"".join(sorted(lines_of(res, k)))).hexdigest()
