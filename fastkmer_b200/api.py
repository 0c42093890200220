"""ctypes binding of libfastkmer_b200.so (C ABI: include/fastkmer_b200.h)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfastkmer_b200.so")

FKM_OK, FKM_EINVAL, FKM_ECUDA, FKM_EIO, FKM_ENOMEM, FKM_EOVERFLOW = 0, -1, -2, -3, -4, -5


class FkmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("fastkmer_b200 error %d: %s" % (code, msg))
        self.code = code


class fkm_config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("k", "m", "x", "max_b", "sequence_type", "use_ht", "write",
                                         "use_kryo_serializer", "use_custom_partitioner", "num_partition_tasks")] + \
               [("dataset", C.c_char_p), ("output_directory", C.c_char_p), ("prefix", C.c_char_p)]


class fkm_stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_positions", "n_bases", "n_kmers", "n_superkmers", "superkmer_bytes",
                                          "n_distinct", "total_count", "digest_sum", "digest_xor", "n_nonempty_bins",
                                          "h2d_bytes", "d2h_bytes", "gpu_launches", "n_batches", "n_fallbacks")] + \
               [("ms_total", C.c_double), ("ms_stage", C.c_double * 8), ("n_folded_records", C.c_uint64), ("ms_fold", C.c_double),
                ("n_mid_bins", C.c_uint64), ("n_slow_bins", C.c_uint64), ("ms_partition", C.c_double)]


class fkm_synth(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("seed_genome", "seed_reads", "seed_errors", "genome_len", "n_reads",
                                          "read_len", "first_read")]


class fkm_synth_long(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("seed_genome", "seed_repeats", "seed_n", "first_pos", "n_bases")]


class Stats(dict):
    """fkm_stats as a dict (plus attribute access)."""
    __getattr__ = dict.__getitem__


_lib = None


def lib_path():
    return _SO


def load_library():
    """Loads the CUDA library.  Fails loudly when it has not been built — there is no other path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise FkmError(FKM_ECUDA, "%s is missing: build it with `make -C fastkmer_b200/csrc` "
                                  "(or __graft_entry__.build()); there is no CPU fallback" % _SO)
    lib = C.CDLL(_SO)
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int32
    cfgp, stp = C.POINTER(fkm_config), C.POINTER(fkm_stats)
    sig = {
        "fkm_last_error": (C.c_char_p, []),
        "fkm_ctx_create": (C.c_int, [C.c_int, vp, C.POINTER(vp)]),
        "fkm_ctx_destroy": (None, [vp]),
        "fkm_ctx_sync": (C.c_int, [vp]),
        "fkm_ctx_set": (C.c_int, [vp, C.c_char_p, C.c_double]),
        "fkm_derive": (C.c_int, [cfgp, C.POINTER(i32), C.c_char_p, C.c_size_t]),
        "fkm_execute_job": (C.c_int, [vp, cfgp, stp]),
        "fkm_execute_job_multi": (C.c_int, [vp, i32, cfgp, stp]),
        "fkm_count_fasta": (C.c_int, [vp, cfgp, vp, u64, C.POINTER(vp), stp]),
        "fkm_count_packed_host": (C.c_int, [vp, cfgp, vp, vp, u64, C.POINTER(vp), stp]),
        "fkm_count_packed_device": (C.c_int, [vp, cfgp, vp, vp, u64, C.POINTER(vp), stp]),
        "fkm_pack_fasta": (C.c_int, [vp, u64, vp, vp, u64, C.POINTER(u64), C.POINTER(u64)]),
        "fkm_pack_fasta_mt": (C.c_int, [vp, u64, vp, vp, u64, C.POINTER(u64), C.POINTER(u64), i32]),
        "fkm_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(vp)]),
        "fkm_host_free": (None, [vp]),
        "fkm_result_size": (u64, [vp]),
        "fkm_result_num_bins": (i32, [vp]),
        "fkm_result_sorted": (i32, [vp]),
        "fkm_result_bin_offsets": (C.c_int, [vp, vp]),
        "fkm_result_copy": (C.c_int, [vp, vp, vp, vp, vp]),
        "fkm_result_write": (C.c_int, [vp, C.c_char_p]),
        "fkm_result_free": (None, [vp]),
        "fkm_result_clone": (C.c_int, [vp, C.POINTER(vp)]),
        "fkm_result_dot": (C.c_int, [vp, vp, vp, C.POINTER(u64)]),
        "fkm_synth_fasta_host": (C.c_int, [C.POINTER(fkm_synth), vp, u64, C.POINTER(u64)]),
        "fkm_synth_packed_device": (C.c_int, [vp, C.POINTER(fkm_synth), C.POINTER(vp), C.POINTER(vp), C.POINTER(u64)]),
        "fkm_synth_long_fasta_host": (C.c_int, [C.POINTER(fkm_synth_long), vp, u64, C.POINTER(u64)]),
        "fkm_synth_long_packed_device": (C.c_int, [vp, C.POINTER(fkm_synth_long), C.POINTER(vp), C.POINTER(vp), C.POINTER(u64)]),
        "fkm_device_free": (C.c_int, [vp, vp]),
        "fkm_debug_window_bins": (C.c_int, [vp, cfgp, vp, vp, u64, vp]),
        "fkm_record_bytes": (i32, [cfgp]),
        "fkm_job_bins": (C.c_int, [vp, cfgp, u64, C.POINTER(i32)]),
        "fkm_debug_multi_plan": (C.c_int, [i32, i32, i32, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, u64, C.POINTER(u64)]),
        "fkm_mg_scan": (C.c_int, [vp, cfgp, vp, vp, u64, vp, vp]),
        "fkm_mg_scan_fasta": (C.c_int, [vp, cfgp, vp, u64, vp, vp, C.POINTER(u64)]),
        "fkm_mg_scatter": (C.c_int, [vp, vp, vp]),
        "fkm_mg_regroup": (C.c_int, [vp, cfgp, vp, u64, vp, vp, u64, C.POINTER(vp)]),
        "fkm_mg_count": (C.c_int, [vp, cfgp, vp, vp, vp, C.POINTER(vp), stp]),
        "fkm_multiseq_fasta": (C.c_int, [vp, cfgp, vp, u64, i32, C.POINTER(i32), C.c_char_p, vp, C.POINTER(vp), stp]),
        "fkm_debug_pack_fasta_device": (C.c_int, [vp, vp, u64, vp, vp, u64, C.POINTER(u64), C.POINTER(u64)]),
        "fkm_total_launches": (u64, []),
    }
    for name, (res, args) in sig.items():
        f = getattr(lib, name)
        f.restype, f.argtypes = res, args
    _lib = lib
    return lib


def _check(rc):
    if rc != FKM_OK:
        raise FkmError(rc, load_library().fkm_last_error().decode("utf-8", "replace"))


def _cfg(configuration):
    """TestConfiguration -> fkm_config (strings kept alive on the returned struct)."""
    c = fkm_config()
    c.k, c.m, c.x, c.max_b = configuration.k, configuration.m, configuration.x, configuration.max_b
    c.sequence_type = configuration.sequenceType
    c.use_ht, c.write = int(bool(configuration.useHT)), int(bool(configuration.write))
    c.use_kryo_serializer = int(bool(configuration.useKryoSerializer))
    c.use_custom_partitioner = int(bool(configuration.useCustomPartitioner))
    c.num_partition_tasks = configuration.numPartitionTasks
    c.dataset = (configuration.dataset or "").encode()
    c.output_directory = (configuration.outputDirectory or "").encode()
    c.prefix = (configuration.prefix or "").encode()
    return c


def _stats(st):
    d = Stats()
    for name, _ in fkm_stats._fields_:
        v = getattr(st, name)
        d[name] = list(v) if name == "ms_stage" else v
    return d


def record_bytes(configuration):
    return int(load_library().fkm_record_bytes(C.byref(_cfg(configuration))))


def derive(configuration):
    """(b, outputDir) as the library derives them (test/package.scala:32-33)."""
    lib = load_library()
    b = C.c_int32()
    buf = C.create_string_buffer(8192)
    _check(lib.fkm_derive(C.byref(_cfg(configuration)), C.byref(b), buf, len(buf)))
    return b.value, buf.value.decode()


def pack_fasta(fasta: bytes):
    """Host-side FASTA -> (bases u64[], invalid u32[], n_positions, n_bases); no GPU needed."""
    lib = load_library()
    arr = np.frombuffer(fasta, dtype=np.uint8)
    n_pos, n_bases = C.c_uint64(), C.c_uint64()
    _check(lib.fkm_pack_fasta(arr.ctypes.data, arr.size, None, None, 0, C.byref(n_pos), C.byref(n_bases)))
    nw = (n_pos.value + 31) // 32
    bases = np.zeros(max(nw, 1), dtype=np.uint64)
    inv = np.zeros(max(nw, 1), dtype=np.uint32)
    _check(lib.fkm_pack_fasta(arr.ctypes.data, arr.size, bases.ctypes.data, inv.ctypes.data, nw * 32,
                              C.byref(n_pos), C.byref(n_bases)))
    return bases[:nw], inv[:nw], n_pos.value, n_bases.value


def _synth(spec):
    s = fkm_synth()
    s.seed_genome, s.seed_reads, s.seed_errors = spec["seeds"]
    s.genome_len, s.n_reads, s.read_len = spec["genome_len"], spec["n_reads"], spec["read_len"]
    s.first_read = spec.get("first_read", 0)
    return s


def synth_fasta(spec, out=None) -> np.ndarray:
    """SURVEY §8(d) synthetic reads as FASTA text (uint8 array).  spec: dict(seeds=(G,R,E), genome_len, n_reads,
    read_len[, first_read]).  `out` may be a preallocated uint8 array (e.g. pinned)."""
    lib = load_library()
    s = _synth(spec)
    n = C.c_uint64()
    _check(lib.fkm_synth_fasta_host(C.byref(s), None, 0, C.byref(n)))
    if out is None:
        out = np.empty(n.value, dtype=np.uint8)
    _check(lib.fkm_synth_fasta_host(C.byref(s), out.ctypes.data, out.size, C.byref(n)))
    return out[:n.value]


def _synth_long(spec):
    s = fkm_synth_long()
    s.seed_genome, s.seed_repeats, s.seed_n = spec["seeds"]
    s.first_pos, s.n_bases = spec.get("first_pos", 0), spec["n_bases"]
    return s


def synth_long_fasta(spec, out=None) -> np.ndarray:
    """One long synthetic record (BASELINE config 3).  spec: dict(seeds=(G,rep,N), n_bases[, first_pos])."""
    lib = load_library()
    s = _synth_long(spec)
    n = C.c_uint64()
    _check(lib.fkm_synth_long_fasta_host(C.byref(s), None, 0, C.byref(n)))
    if out is None:
        out = np.empty(n.value, dtype=np.uint8)
    _check(lib.fkm_synth_long_fasta_host(C.byref(s), out.ctypes.data, out.size, C.byref(n)))
    return out[:n.value]


class CountResult:
    """Per-bin (canonical k-mer, count) arrays held on the device; numpy copies on demand.

    The device arrays belong to the context's job arena and stay valid until the next job on
    that context: call arrays()/write() before counting again (arrays() caches host copies)."""

    def __init__(self, handle, k, ctx=None):
        self._h = handle
        self.k = k
        self._cache = None
        self._ctx = ctx            # the arrays live in the context's arena: keep it alive

    def __len__(self):
        return int(load_library().fkm_result_size(self._h))

    @property
    def num_bins(self):
        return int(load_library().fkm_result_num_bins(self._h))

    @property
    def sorted(self):
        return bool(load_library().fkm_result_sorted(self._h))

    def bin_offsets(self):
        off = np.zeros(self.num_bins + 1, dtype=np.uint64)
        _check(load_library().fkm_result_bin_offsets(self._h, off.ctypes.data))
        return off

    def arrays(self):
        """dict(bin int32[], hi u64[], lo u64[], cnt u32[]) in the library's bin-major order."""
        if self._cache is None:
            n = len(self)
            bin_ = np.empty(n, dtype=np.int32)
            hi = np.empty(n, dtype=np.uint64)
            lo = np.empty(n, dtype=np.uint64)
            cnt = np.empty(n, dtype=np.uint32)
            _check(load_library().fkm_result_copy(self._h, bin_.ctypes.data, hi.ctypes.data, lo.ctypes.data, cnt.ctypes.data))
            self._cache = {"bin": bin_, "hi": hi, "lo": lo, "cnt": cnt}
        return self._cache

    def sorted_arrays(self):
        """Same, sorted by (bin, k-mer): the order parity is defined on."""
        a = self.arrays()
        order = np.lexsort((a["lo"], a["hi"], a["bin"]))
        return {k_: v[order] for k_, v in a.items()}

    def write(self, out_dir: str):
        _check(load_library().fkm_result_write(self._h, out_dir.encode()))

    def clone(self):
        """A copy in plain device memory that survives later jobs on the context."""
        h = C.c_void_p()
        _check(load_library().fkm_result_clone(self._h, C.byref(h)))
        return CountResult(h, self.k, self._ctx)

    def dot(self, other):
        """sum over common (bin, k-mer) of count * other's count — both must be clones of sorted results."""
        v = C.c_uint64()
        _check(load_library().fkm_result_dot(self._ctx._h, self._h, other._h, C.byref(v)))
        return v.value

    def free(self):
        if self._h:
            load_library().fkm_result_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One per (process, GPU).  stream: a raw cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream)."""

    def __init__(self, device=-1, stream=None):
        lib = load_library()
        h = C.c_void_p()
        _check(lib.fkm_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(h)))
        self._h = h
        self._dev_bufs = []

    def set(self, name, value):
        _check(load_library().fkm_ctx_set(self._h, name.encode(), float(value)))

    def sync(self):
        _check(load_library().fkm_ctx_sync(self._h))

    def close(self):
        if self._h:
            for p in self._dev_bufs:
                load_library().fkm_device_free(self._h, p)
            self._dev_bufs = []
            load_library().fkm_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the drop-in call (SparkBinKmerCounter.scala:989)
    def execute_job(self, configuration) -> Stats:
        if getattr(configuration, "debug", False) and configuration.write:
            # debug = true moves the output to /tmp/<stem> without the _s<type> suffix (test/package.scala:33): count here, write there
            with open(configuration.dataset, "rb") as f:
                text = f.read()
            res, st = self.count_fasta(configuration, text)
            res.write(configuration.outputDir)
            return st
        st = fkm_stats()
        cfg = _cfg(configuration)
        _check(load_library().fkm_execute_job(self._h, C.byref(cfg), C.byref(st)))
        return _stats(st)

    @staticmethod
    def execute_job_multi(configuration, devices) -> Stats:
        """The drop-in call on several GPUs of this node (fkm_execute_job_multi): no context needed, the library makes its own."""
        st = fkm_stats()
        cfg = _cfg(configuration)
        dev = (C.c_int32 * len(devices))(*devices)
        _check(load_library().fkm_execute_job_multi(dev, len(devices), C.byref(cfg), C.byref(st)))
        return _stats(st)

    # ---- in-memory variants
    def count_fasta(self, configuration, fasta, want_result=True):
        """fasta: bytes or uint8 numpy array (host).  -> (CountResult | None, Stats)"""
        arr = np.frombuffer(fasta, dtype=np.uint8) if isinstance(fasta, (bytes, bytearray)) else fasta
        st, cfg, h = fkm_stats(), _cfg(configuration), C.c_void_p()
        _check(load_library().fkm_count_fasta(self._h, C.byref(cfg), arr.ctypes.data, arr.size,
                                              C.byref(h) if want_result else None, C.byref(st)))
        return (CountResult(h, configuration.k, self) if want_result else None), _stats(st)

    def count_packed_host(self, configuration, bases, inv, n_positions, want_result=True):
        st, cfg, h = fkm_stats(), _cfg(configuration), C.c_void_p()
        _check(load_library().fkm_count_packed_host(self._h, C.byref(cfg), bases.ctypes.data, inv.ctypes.data, n_positions,
                                                    C.byref(h) if want_result else None, C.byref(st)))
        return (CountResult(h, configuration.k, self) if want_result else None), _stats(st)

    def count_packed_device(self, configuration, d_bases, d_inv, n_positions, want_result=True):
        """d_bases / d_inv: raw device pointers (ints) in the library's packed layout."""
        st, cfg, h = fkm_stats(), _cfg(configuration), C.c_void_p()
        _check(load_library().fkm_count_packed_device(self._h, C.byref(cfg), C.c_void_p(d_bases), C.c_void_p(d_inv), n_positions,
                                                      C.byref(h) if want_result else None, C.byref(st)))
        return (CountResult(h, configuration.k, self) if want_result else None), _stats(st)

    def synth_packed_device(self, spec):
        """-> (d_bases, d_inv, n_positions): synthetic reads generated in device memory (freed with the context)."""
        s = _synth(spec)
        b, i, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        _check(load_library().fkm_synth_packed_device(self._h, C.byref(s), C.byref(b), C.byref(i), C.byref(n)))
        self._dev_bufs += [b, i]
        return b.value, i.value, n.value

    def synth_long_packed_device(self, spec):
        s = _synth_long(spec)
        b, i, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        _check(load_library().fkm_synth_long_packed_device(self._h, C.byref(s), C.byref(b), C.byref(i), C.byref(n)))
        self._dev_bufs += [b, i]
        return b.value, i.value, n.value

    def free_device(self, ptr):
        for p in list(self._dev_bufs):
            if p.value == ptr:
                self._dev_bufs.remove(p)
        _check(load_library().fkm_device_free(self._h, C.c_void_p(ptr)))

    # ---- staged entry points (multi-GPU; see fastkmer_b200/multigpu.py)
    def job_bins(self, configuration, n_positions=0):
        """Internal bins of a job on this context: the configuration's bins times the job's bin split (knob bin_split)."""
        n = C.c_int32()
        cfg = _cfg(configuration)
        _check(load_library().fkm_job_bins(self._h, C.byref(cfg), n_positions, C.byref(n)))
        return n.value

    def mg_scan(self, configuration, d_bases, d_inv, n_positions):
        """-> (hist_rec, hist_kmer) uint64[internal bins]: records / k-mers this shard puts into every (internal) bin."""
        b = self.job_bins(configuration, n_positions)
        rec = np.zeros(b, dtype=np.uint64)
        kmer = np.zeros(b, dtype=np.uint64)
        cfg = _cfg(configuration)
        _check(load_library().fkm_mg_scan(self._h, C.byref(cfg), C.c_void_p(d_bases), C.c_void_p(d_inv), n_positions,
                                          rec.ctypes.data, kmer.ctypes.data))
        return rec, kmer

    def mg_scan_fasta(self, configuration, fasta):
        arr = np.frombuffer(fasta, dtype=np.uint8) if isinstance(fasta, (bytes, bytearray)) else fasta
        b = self.job_bins(configuration, arr.size)
        rec = np.zeros(b, dtype=np.uint64)
        kmer = np.zeros(b, dtype=np.uint64)
        nb = C.c_uint64()
        cfg = _cfg(configuration)
        _check(load_library().fkm_mg_scan_fasta(self._h, C.byref(cfg), arr.ctypes.data, arr.size, rec.ctypes.data, kmer.ctypes.data,
                                                C.byref(nb)))
        return rec, kmer, nb.value

    def mg_scatter(self, bin_base, d_send):
        bin_base = np.ascontiguousarray(bin_base, dtype=np.uint64)
        _check(load_library().fkm_mg_scatter(self._h, bin_base.ctypes.data, C.c_void_p(d_send)))

    def mg_regroup(self, configuration, d_recv, n_records, seg_src, seg_dst):
        seg_src = np.ascontiguousarray(seg_src, dtype=np.uint64)
        seg_dst = np.ascontiguousarray(seg_dst, dtype=np.uint64)
        out = C.c_void_p()
        cfg = _cfg(configuration)
        _check(load_library().fkm_mg_regroup(self._h, C.byref(cfg), C.c_void_p(d_recv), n_records, seg_src.ctypes.data,
                                             seg_dst.ctypes.data, len(seg_dst), C.byref(out)))
        return out.value or 0

    def mg_count(self, configuration, d_records, bin_rec, bin_kmer, want_result=True):
        bin_rec = np.ascontiguousarray(bin_rec, dtype=np.uint64)
        bin_kmer = np.ascontiguousarray(bin_kmer, dtype=np.uint64)
        st, cfg, h = fkm_stats(), _cfg(configuration), C.c_void_p()
        _check(load_library().fkm_mg_count(self._h, C.byref(cfg), C.c_void_p(d_records), bin_rec.ctypes.data, bin_kmer.ctypes.data,
                                           C.byref(h) if want_result else None, C.byref(st)))
        return (CountResult(h, configuration.k, self) if want_result else None), _stats(st)

    def multiseq_fasta(self, configuration, fasta, max_samples=64, want_result=False):
        """Multi-sample job (skc.multisequence).  -> (sample names, S x S float64 squared-euclidean distances,
        merged CountResult | None, Stats)"""
        arr = np.frombuffer(fasta, dtype=np.uint8) if isinstance(fasta, (bytes, bytearray)) else fasta
        st, cfg, h, ns = fkm_stats(), _cfg(configuration), C.c_void_p(), C.c_int32()
        names = C.create_string_buffer(64 * max_samples)
        dist = np.zeros((max_samples, max_samples), dtype=np.float64)
        _check(load_library().fkm_multiseq_fasta(self._h, C.byref(cfg), arr.ctypes.data, arr.size, max_samples, C.byref(ns), names,
                                                 dist.ctypes.data, C.byref(h) if want_result else None, C.byref(st)))
        S = ns.value
        tags = [names.raw[64 * i:64 * (i + 1)].split(b"\0", 1)[0].decode() for i in range(S)]
        return tags, dist[:S, :S].copy(), (CountResult(h, configuration.k, self) if want_result else None), _stats(st)

    def pack_fasta_device(self, fasta: bytes):
        """Device ingest of FASTA text, copied back: (bases, inv, n_positions, n_bases) — test hook."""
        arr = np.frombuffer(fasta, dtype=np.uint8)
        n_pos, n_bases = C.c_uint64(), C.c_uint64()
        nw = (arr.size + 1 + 31) // 32 + 1
        bases = np.zeros(nw, dtype=np.uint64)
        inv = np.zeros(nw, dtype=np.uint32)
        _check(load_library().fkm_debug_pack_fasta_device(self._h, arr.ctypes.data, arr.size, bases.ctypes.data, inv.ctypes.data,
                                                          nw * 32, C.byref(n_pos), C.byref(n_bases)))
        nw = (n_pos.value + 31) // 32
        return bases[:nw], inv[:nw], n_pos.value, n_bases.value

    def window_bins(self, configuration, bases, inv, n_positions):
        out = np.empty(max(n_positions, 1), dtype=np.int32)
        cfg = _cfg(configuration)
        _check(load_library().fkm_debug_window_bins(self._h, C.byref(cfg), bases.ctypes.data, inv.ctypes.data, n_positions, out.ctypes.data))
        return out[:n_positions]


def host_alloc(nbytes) -> np.ndarray:
    """Pinned host memory as a uint8 numpy array (kept alive by the returned array's base)."""
    lib = load_library()
    p = C.c_void_p()
    _check(lib.fkm_host_alloc(nbytes, C.byref(p)))
    buf = (C.c_uint8 * nbytes).from_address(p.value)
    arr = np.frombuffer(buf, dtype=np.uint8)
    arr_ptr = p.value

    class _Owner:
        def __del__(self_inner):
            lib.fkm_host_free(C.c_void_p(arr_ptr))
    host_alloc._owners.append((_Owner(), buf))
    return arr


host_alloc._owners = []
