"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
inputs — bit-exact (bin, canonical k-mer, count) sets.  Run with -m gpu on a B200."""
import hashlib
import os
import random
import subprocess

import numpy as np
import pytest

import clean_spec
import fastkmer_b200 as fk
import oracle_lib
from test_oracle_kat import GOLD, _rand_fasta

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx():
    c = fk.Context(0)
    yield c
    c.close()


def cfg(k, m, x, B, ht, **kw):
    return fk.TestConfiguration("", "", k, m, x, max_b=B, useHT=bool(ht), write=False, **kw)


def assert_same(got, want, what=""):
    for key in ("bin", "hi", "lo", "cnt"):
        assert got[key].shape == want[key].shape, "%s: %d entries vs oracle %d" % (what, got[key].size, want[key].size)
        assert np.array_equal(got[key], want[key]), "%s: %s differs" % (what, key)


def check(ctx, oracle, fasta: bytes, k, m, x, B, ht, what=""):
    res, st = ctx.count_fasta(cfg(k, m, x, B, ht), fasta)
    want = oracle.count(fasta, k, m, x, B, ht)
    assert_same(res.sorted_arrays(), want, what)
    ws = want["stats"]
    assert st["n_kmers"] == ws["n_kmers"] == st["total_count"]
    assert st["n_distinct"] == ws["n_distinct"] and st["n_bases"] == ws["n_bases"]
    assert (st["digest_sum"], st["digest_xor"]) == (ws["digest_sum"], ws["digest_xor"])
    if not ht:                                  # sort path: ascending inside each bin, as the reference writes them
        a = res.arrays()
        assert_same(a, want, what + " (unsorted order)")
    return res, st


# ---------------------------------------------------------------- K1: per-window bins
@pytest.mark.parametrize("k,m,B", [(28, 10, 2048), (31, 11, 4096), (55, 13, 2048), (5, 3, 64), (10, 10, 999), (64, 15, 5000),
                                    (33, 4, 256), (32, 3, 64), (20, 5, 2000)])
def test_window_bins_match_oracle(ctx, oracle, k, m, B):
    rng = random.Random(k * 100 + m)
    seq = "".join("N" if rng.random() < 0.01 else rng.choice("ACGT") for _ in range(30000))
    seq = seq[:5000] + "A" * 300 + "ACACACAC" * 40 + "T" * 200 + seq[5000:]
    bases, inv, n_pos, _ = fk.pack_fasta((">s\n%s\n" % seq).encode())
    got = ctx.window_bins(cfg(k, m, 3, B, 1), bases, inv, n_pos)
    want = oracle.window_bins(seq.encode(), k, m, B)
    assert np.array_equal(got[:want.size], want)
    assert (got[want.size:] == -1).all()


# ---------------------------------------------------------------- golden vectors
@pytest.mark.parametrize("gid", sorted(GOLD))
@pytest.mark.parametrize("ht", [0, 1])
def test_golden_digests(ctx, oracle, gid, ht):
    (seed, G, R, L), (k, m, x, B), nk, dist, nbins, maxc, _, _, sha = GOLD[gid]
    fasta = oracle.gen_lcg_fasta(seed, G, R, L)
    res, st = check(ctx, oracle, fasta, k, m, x, B, ht, gid)
    assert (st["n_kmers"], st["n_distinct"], st["n_nonempty_bins"]) == (nk, dist, nbins)
    assert oracle_lib.digest_of(res.sorted_arrays(), k) == sha


# ---------------------------------------------------------------- edge cases of the reference's input handling
@pytest.mark.parametrize("k,m,x,B", [(5, 3, 1, 64), (12, 4, 2, 100), (20, 5, 3, 2000), (28, 10, 3, 2048), (31, 11, 3, 4096),
                                      (32, 7, 2, 333), (33, 8, 3, 512), (55, 13, 3, 2048), (60, 9, 4, 77), (61, 6, 3, 4096),
                                      (64, 15, 1, 5000), (15, 15, 2, 3), (7, 7, 1, 1 << 20)])
@pytest.mark.parametrize("ht", [0, 1])
def test_random_reads_with_invalid_bytes(ctx, oracle, k, m, x, B, ht):
    if not ht and k + x > 64:
        pytest.skip("oracle's sort path holds (k+x)-mers in 128 bits")
    rng = random.Random(k * 1000 + m + ht)
    for alphabet, width in (("ACGT", None), ("AC", 17), ("ACGT", 70), ("A", None)):
        fasta = _rand_fasta(rng, 40, 0, 3 * k + 60, alphabet=alphabet, width=width).encode()
        check(ctx, oracle, fasta, k, m, x, B, ht, "%s/%s" % (alphabet, width))


@pytest.mark.parametrize("ht", [0, 1])
def test_degenerate_inputs(ctx, oracle, ht):
    for text in (b"", b">empty\n", b">short\nACGT\n", b">allN\n" + b"N" * 500 + b"\n", b">lower\n" + b"acgt" * 50 + b"\n",
                 b">one\n" + b"ACGTTGCATGCAGGCTTAACCGGTAAGC" + b"\n", b"no header\nACGTACGTACGTACGTACGTACGTACGTACGTACGT\n",
                 b">crlf\r\n" + b"ACGGTCAGGT" * 9 + b"\r\n" + b"ACGGTCAGGT" * 9 + b"\r\n"):
        res, st = check(ctx, oracle, text, 28, 10, 3, 2048, ht, repr(text[:12]))
    assert len(res) == 0 or st["n_kmers"] > 0


def test_long_homopolymer_and_repeats(ctx, oracle):          # runs far longer than one record's capacity
    rng = random.Random(5)
    unit = "".join(rng.choice("ACGT") for _ in range(37))
    seq = "A" * 5000 + unit * 300 + "C" * 3000 + "AC" * 2000 + "".join(rng.choice("ACGT") for _ in range(20000))
    fasta = (">g\n" + "\n".join(seq[i:i + 70] for i in range(0, len(seq), 70)) + "\n").encode()
    for (k, m, x, B) in ((28, 10, 3, 2048), (31, 11, 3, 4096), (55, 13, 3, 2048), (64, 12, 1, 512)):
        for ht in (0, 1):
            if not ht and k + x > 64:
                continue
            check(ctx, oracle, fasta, k, m, x, B, ht, "k%d ht%d" % (k, ht))


# ---------------------------------------------------------------- BASELINE configs at reduced scale
SYNTH = {
    "config1": (dict(seeds=(1001, 1002, 1003), genome_len=50000, n_reads=10000, read_len=100), (28, 10, 3, 2048, 0, 0)),
    "config2": (dict(seeds=(2001, 2002, 2003), genome_len=100000, n_reads=20000, read_len=150), (28, 10, 3, 2048, 1, 0)),
    "config4": (dict(seeds=(4001, 4002, 4003), genome_len=100000, n_reads=20000, read_len=150), (55, 13, 3, 2048, 0, 0)),
    "config4ht": (dict(seeds=(4001, 4002, 4003), genome_len=100000, n_reads=8000, read_len=150), (55, 13, 3, 2048, 1, 0)),
}


@pytest.mark.parametrize("name", sorted(SYNTH))
def test_baseline_configs_reduced(ctx, oracle, name):
    spec, (k, m, x, B, ht, seqtype) = SYNTH[name]
    fasta = fk.synth_fasta(spec).tobytes()
    res, st = check(ctx, oracle, fasta, k, m, x, B, ht, name)
    # the device generator lays out the same reads: identical digest without any host text
    d_b, d_i, n_pos = ctx.synth_packed_device(spec)
    _, st2 = ctx.count_packed_device(cfg(k, m, x, B, ht), d_b, d_i, n_pos, want_result=False)
    assert (st2["digest_sum"], st2["digest_xor"], st2["n_kmers"], st2["n_distinct"]) == \
           (st["digest_sum"], st["digest_xor"], st["n_kmers"], st["n_distinct"])
    ctx.free_device(d_b)
    ctx.free_device(d_i)


def _long_genome(n, seed):
    rng = random.Random(seed)
    g = [rng.choice("ACGT") for _ in range(n)]
    for _ in range(max(1, n // 100000)):                       # 5-kb repeats
        src = rng.randrange(0, n - 5000)
        for c in range(10):
            dst = rng.randrange(0, n - 5000)
            g[dst:dst + 5000] = g[src:src + 5000]
    for _ in range(max(1, n // 200000)):                       # N runs of 1000
        p = rng.randrange(0, n - 1000)
        g[p:p + 1000] = "N" * 1000
    return "".join(g)


def test_config3_long_sequence_reduced(ctx, oracle):          # sequenceType=1: one record, 70-column lines
    seq = _long_genome(1_200_000, 3001)
    fasta = (">chr1 synthetic\n" + "\n".join(seq[i:i + 70] for i in range(0, len(seq), 70)) + "\n").encode()
    for ht in (0, 1):
        res, st = ctx.count_fasta(cfg(31, 11, 3, 4096, ht, sequenceType=1), fasta)
        want = oracle.count(fasta, 31, 11, 3, 4096, ht, threads=8)
        assert_same(res.sorted_arrays(), want, "config3 ht%d" % ht)
        assert st["n_kmers"] == want["stats"]["n_kmers"]


# ---------------------------------------------------------------- size-independent properties at larger scale
def test_properties_at_scale(ctx):
    spec = dict(seeds=(2001, 2002, 2003), genome_len=2_000_000, n_reads=400_000, read_len=150)    # 60 Mbases
    d_b, d_i, n_pos = ctx.synth_packed_device(spec)
    _, ht = ctx.count_packed_device(cfg(28, 10, 3, 2048, 1), d_b, d_i, n_pos, want_result=False)
    _, so = ctx.count_packed_device(cfg(28, 10, 3, 2048, 0), d_b, d_i, n_pos, want_result=False)
    for key in ("n_kmers", "n_distinct", "total_count", "digest_sum", "digest_xor", "n_nonempty_bins"):
        assert ht[key] == so[key], key                         # hash path == sort path (the reference's own cross-check)
    assert ht["total_count"] == ht["n_kmers"] > 40_000_000
    # the bin count changes where k-mers land, never which k-mers exist
    _, b1 = ctx.count_packed_device(cfg(28, 10, 3, 1, 1), d_b, d_i, n_pos, want_result=False)
    assert (b1["n_kmers"], b1["n_distinct"]) == (ht["n_kmers"], ht["n_distinct"])
    ctx.free_device(d_b)
    ctx.free_device(d_i)


def test_strand_symmetry_and_split_invariance(ctx):
    rng = random.Random(11)
    seq = "".join(rng.choice("ACGT") for _ in range(200000))
    a, sa = ctx.count_fasta(cfg(28, 10, 3, 2048, 1), (">a\n%s\n" % seq).encode())
    b, sb = ctx.count_fasta(cfg(28, 10, 3, 2048, 1), (">a\n%s\n" % clean_spec.revcomp(seq)).encode())
    assert (sa["digest_sum"], sa["digest_xor"]) == (sb["digest_sum"], sb["digest_xor"])
    # overlapping chunks with a (k-1)-base halo count the same k-mers as the whole sequence
    k = 28
    cuts = [0, 50000, 123457, 200000]
    text = "".join(">c%d\n%s\n" % (i, seq[cuts[i]:min(len(seq), cuts[i + 1] + k - 1)]) for i in range(3))
    c, sc = ctx.count_fasta(cfg(28, 10, 3, 2048, 0), text.encode())
    assert (sa["digest_sum"], sa["digest_xor"], sa["n_kmers"]) == (sc["digest_sum"], sc["digest_xor"], sc["n_kmers"])


def test_small_table_budget_batches_and_retry(ctx, oracle):
    spec = dict(seeds=(7, 8, 9), genome_len=30000, n_reads=30000, read_len=100)      # 100x coverage: few distinct per k-mer
    fasta = fk.synth_fasta(spec).tobytes()
    want = oracle.count(fasta, 28, 10, 3, 2048, 1, threads=8)
    c2 = fk.Context(0)
    try:
        c2.set("count_mode", 0)                               # the global-table pipeline (fallback of the shared-memory path)
        c2.set("table_budget_bytes", 1 << 20)
        c2.set("async_table_bytes", 0)                        # synchronous batches only
        res, st = c2.count_fasta(cfg(28, 10, 3, 2048, 1), fasta)
        assert st["n_batches"] > 4
        assert_same(res.sorted_arrays(), want, "tiny budget")
        c2.set("sort_budget_keys", 50000)
        res, st = c2.count_fasta(cfg(28, 10, 3, 2048, 0), fasta)
        assert st["n_batches"] > 4
        assert_same(res.arrays(), want, "tiny sort budget")
    finally:
        c2.close()


def test_async_phase_overflow_falls_back(oracle):
    """A wrong distinct/k-mer estimate overflows the pre-planned tables of the asynchronous batches: the job must notice and redo the
    bins synchronously with the same result."""
    spec = dict(seeds=(17, 18, 19), genome_len=400000, n_reads=40000, read_len=100)
    fasta = fk.synth_fasta(spec).tobytes()
    want = oracle.count(fasta, 28, 10, 3, 2048, 1, threads=8)
    c2 = fk.Context(0)
    try:
        c2.set("count_mode", 0)                               # the global-table pipeline (fallback of the shared-memory path)
        res, st = c2.count_fasta(cfg(28, 10, 3, 2048, 1), fasta)
        assert st["n_fallbacks"] == 0 and st["n_batches"] >= 2
        assert_same(res.sorted_arrays(), want, "async phase")
        c2.set("debug_rho_scale", 0.02)
        res, st = c2.count_fasta(cfg(28, 10, 3, 2048, 1), fasta)
        assert st["n_fallbacks"] == 1
        assert_same(res.sorted_arrays(), want, "after fallback")
        assert (st["digest_sum"], st["digest_xor"]) == (want["stats"]["digest_sum"], want["stats"]["digest_xor"])
        c2.set("debug_rho_scale", 1.0)
        c2.set("debug_event_scale", 0.001)                 # run-event list too small: second scan writes the records directly
        for ht in (1, 0):
            res, st = c2.count_fasta(cfg(28, 10, 3, 2048, ht), fasta)
            assert st["n_fallbacks"] == 1
            assert_same(res.sorted_arrays(), want, "event overflow ht%d" % ht)
        c2.set("debug_event_scale", 1.0)
        c2.set("async_table_bytes", 1 << 16)                  # many tiny asynchronous batches
        res, st = c2.count_fasta(cfg(28, 10, 3, 2048, 1), fasta)
        assert st["n_fallbacks"] == 0 and st["n_batches"] > 100
        assert_same(res.sorted_arrays(), want, "tiny async batches")
        for k, m in ((55, 13), (33, 9)):                   # wide keys through the same phases
            w2 = oracle.count(fasta, k, m, 3, 2048, 1, threads=8)
            res, st = c2.count_fasta(cfg(k, m, 3, 2048, 1), fasta)
            assert_same(res.sorted_arrays(), w2, "wide async k=%d" % k)
    finally:
        c2.close()


@pytest.mark.parametrize("mode", [2, 1])
def test_smem_path_tiers(oracle, mode):
    configs = ((28, 10, 2048), (55, 13, 2048), (33, 9, 64), (31, 11, 1), (12, 4, 100), (22, 8, 300)) if mode == 2 else ((28, 10, 2048), (55, 13, 2048), (33, 9, 64))
    """useHT=1 counts in shared-memory tables (mode 2: k-mers hash-partitioned into sub-buckets, k_count_keys; mode 1: dual-minimizer
    mid bins, k_count_smem).  A sub-bucket / mid bin with more distinct k-mers than the table holds is
    redone by its CTA in a private global table (slow path); if that overflows too, or the output estimate is too small, or
    (mode 1) the run-event lists overflow, the whole job is redone by the global-table pipeline.  Every tier gives the oracle's result."""
    deep = fk.synth_fasta(dict(seeds=(31, 32, 33), genome_len=50000, n_reads=50000, read_len=100)).tobytes()      # 100x coverage
    flat = fk.synth_fasta(dict(seeds=(34, 35, 36), genome_len=30000000, n_reads=40000, read_len=150)).tobytes()   # nearly all distinct
    c2 = fk.Context(0)
    try:
        c2.set("count_mode", mode)
        if mode == 2:
            c2.set("part_max_subs", 1e9)      # (the tiny tables below need thousands of sub-buckets per bin: keep the size guard out of the way)
        for text, label in ((deep, "deep"), (flat, "flat")):
            for k, m, B in configs:
                want = oracle.count(text, k, m, 3, B, 1, threads=8)
                ws = want["stats"]
                # (mode 2 plans its sub-buckets for the table it has: a small table alone never overflows, a wrong distinct / k-mer estimate does)
                small = {"smem_table_slots": 256} if mode == 1 else {"smem_table_slots": 64, "debug_rho_scale": 0.02}
                for knobs, tier in (({}, "fast"), (small, "slow"),
                                    (dict(small, smem_slow_slots=1024), "fallback"),
                                    ({"debug_rho_scale": 0.02}, "rho"), ({"debug_event_scale": 0.001}, "events")):
                    for name in ("smem_table_slots", "debug_rho_scale", "debug_event_scale"):
                        c2.set(name, {"smem_table_slots": 0, "debug_rho_scale": 1.0, "debug_event_scale": 1.0}[name])
                    c2.set("smem_slow_slots", 1 << 20)
                    for name, v in knobs.items():
                        c2.set(name, v)
                    res, st = c2.count_fasta(cfg(k, m, 3, B, 1), text)
                    what = "smem mode %d %s %s k=%d B=%d" % (mode, tier, label, k, B)
                    assert_same(res.sorted_arrays(), want, what)
                    assert (st["digest_sum"], st["digest_xor"]) == (ws["digest_sum"], ws["digest_xor"]), what
                    assert st["n_kmers"] == ws["n_kmers"] == st["total_count"] and st["n_distinct"] == ws["n_distinct"], what
                    if tier == "fast":      # (with few bins the one-bin sample may misjudge the distinct share: a few slow mid bins are fine)
                        assert st["n_mid_bins"] > 0 and st["n_fallbacks"] == 0 and st["n_slow_bins"] <= (0 if B == 2048 else st["n_mid_bins"] // 4), what
                    if tier == "slow":
                        assert st["n_fallbacks"] == 0 and st["n_mid_bins"] > 0, what
                        if B >= 64 and k >= 22 and (mode == 1 or label == "flat"):
                            assert st["n_slow_bins"] > 0, what
                    if tier == "fallback" and mode == 2 and label == "flat" and B >= 64 and k >= 22:
                        assert st["n_fallbacks"] >= 1 and st["n_mid_bins"] == 0, what
                    if tier == "events":      # (mode 2 keeps its tables: only the scatter stage falls back to a second scan)
                        assert st["n_fallbacks"] >= 1 and (st["n_mid_bins"] == 0) == (mode == 1), what
    finally:
        c2.close()


@pytest.mark.parametrize("mode", [2])
def test_smem_path_edge_cases(oracle, mode):
    """The reference's input edge cases (empty, short, all-N, lowercase, CRLF, homopolymers and short-period repeats far longer than
    a record, k = m, k = 64, B = 1, B > 4096) through the shared-memory tables."""
    rng = random.Random(11)
    unit = "".join(rng.choice("ACGT") for _ in range(37))
    seq = "A" * 5000 + unit * 300 + "C" * 3000 + "AC" * 2000 + "".join(rng.choice("ACGT") for _ in range(20000))
    long_text = (">g\n" + "\n".join(seq[i:i + 70] for i in range(0, len(seq), 70)) + "\n").encode()
    texts = [b"", b">empty\n", b">short\nACGT\n", b">allN\n" + b"N" * 500 + b"\n", b">lower\n" + b"acgt" * 50 + b"\n",
             b">one\n" + b"ACGTTGCATGCAGGCTTAACCGGTAAGC" + b"\n", b">crlf\r\n" + b"ACGGTCAGGT" * 9 + b"\r\n" + b"ACGGTCAGGT" * 9 + b"\r\n", long_text]
    for alphabet, width in (("ACGT", None), ("AC", 17), ("A", None)):
        texts.append(_rand_fasta(rng, 40, 0, 250, alphabet=alphabet, width=width).encode())
    c2 = fk.Context(0)
    try:
        c2.set("count_mode", mode)
        for (k, m, B) in ((28, 10, 2048), (5, 3, 64), (31, 11, 4096), (32, 7, 333), (33, 8, 512), (55, 13, 2048), (64, 12, 5000), (13, 13, 3), (20, 5, 1), (7, 7, 1 << 20)):      # (m = 15, whose norm table costs the oracle 20 s a call, is covered by test_random_reads_with_invalid_bytes)
            for t in texts:
                res, st = check(c2, oracle, t, k, m, 3, B, 1, "smem mode %d k=%d B=%d %r" % (mode, k, B, t[:10]))
                assert st["n_fallbacks"] == 0
    finally:
        c2.close()


def test_count_overflow_is_reported():
    """Counts are 32-bit like the reference's (Int: SBKC:562,676).  One k-mer seen more than 2^32 - 1 times (a homopolymer of
    4.3 G bases) must end in FKM_EOVERFLOW on every count path, never in a wrapped count."""
    from fastkmer_b200 import api
    n_pos = (1 << 32) + 5000
    nw = (n_pos + 31) // 32
    bases = np.zeros(nw, dtype=np.uint64)                       # all 'A'
    inv = np.zeros(nw, dtype=np.uint32)
    c2 = fk.Context(0)
    try:
        for ht, mode in ((1, 0), (1, 1), (1, 2), (0, 0)):
            c2.set("count_mode", mode)
            with pytest.raises(api.FkmError) as e:
                c2.count_packed_host(cfg(28, 10, 3, 2048, ht), bases, inv, n_pos, want_result=False)
            assert e.value.code == api.FKM_EOVERFLOW, (ht, mode, str(e.value))
        # just below the limit the count is exact
        n_ok = (1 << 32) - 1 + 27
        c2.set("count_mode", 0)
        res, st = c2.count_packed_host(cfg(28, 10, 3, 2048, 1), bases, inv, n_ok)
        a = res.arrays()
        assert a["cnt"].tolist() == [0xFFFFFFFF] and a["lo"].tolist() == [0] and st["n_kmers"] == 0xFFFFFFFF
    finally:
        c2.close()


@pytest.mark.parametrize("mode", [2, 1])
def test_smem_and_global_tables_agree_at_scale(ctx, mode):
    """2 M reads: the shared-memory tables and the global tables give the same digest, and the sum of counts is the number of windows."""
    spec = dict(seeds=(41, 42, 43), genome_len=10_000_000, n_reads=2_000_000, read_len=150)
    d_b, d_i, n_pos = ctx.synth_packed_device(spec)
    try:
        for k, m in ((28, 10), (55, 13)):
            c = cfg(k, m, 3, 2048, 1)
            ctx.set("count_mode", mode)
            _, a = ctx.count_packed_device(c, d_b, d_i, n_pos, want_result=False)
            ctx.set("count_mode", 0)
            _, b = ctx.count_packed_device(c, d_b, d_i, n_pos, want_result=False)
            assert a["n_mid_bins"] > 0 and b["n_mid_bins"] == 0
            for key in ("n_kmers", "n_distinct", "total_count", "digest_sum", "digest_xor"):
                assert a[key] == b[key], (k, key)
            assert a["total_count"] == a["n_kmers"] and a["n_fallbacks"] == 0
    finally:
        ctx.set("count_mode", 0)
        ctx.free_device(d_b)
        ctx.free_device(d_i)


def test_internal_bins(oracle):
    """Deep inputs / multi-GPU ranks cut every bin into 2^j internal bins by a second hash of the signature (knob bin_split; chosen
    from the input size otherwise): same (bin, k-mer, count) as the oracle, the configuration's bins outside, on the hash path with
    shared-memory and with global-memory tables; the sort path and the dual-minimizer mode ignore the knob."""
    deep = fk.synth_fasta(dict(seeds=(81, 82, 83), genome_len=60000, n_reads=40000, read_len=100)).tobytes()
    flat = fk.synth_fasta(dict(seeds=(84, 85, 86), genome_len=20000000, n_reads=30000, read_len=150)).tobytes()
    c2 = fk.Context(0)
    try:
        for text, label in ((deep, "deep"), (flat, "flat")):
            for k, m, B in ((28, 10, 2048), (55, 13, 300), (31, 11, 1), (12, 4, 100)):
                want = oracle.count(text, k, m, 3, B, 1, threads=8)
                for mode in (2, 0, 1):
                    c2.set("count_mode", mode)
                    for split in (4, 64, 1):
                        c2.set("bin_split", split)
                        res, st = c2.count_fasta(cfg(k, m, 3, B, 1), text)
                        what = "%s k=%d B=%d mode %d split %d" % (label, k, B, mode, split)
                        assert_same(res.sorted_arrays(), want, what)
                        assert (st["digest_sum"], st["digest_xor"], st["n_nonempty_bins"]) == \
                               (want["stats"]["digest_sum"], want["stats"]["digest_xor"], np.unique(want["bin"]).size), what
                        a = res.arrays()
                        assert np.all(np.diff(a["bin"]) >= 0), what              # a bin's entries are contiguous
                c2.set("count_mode", 2)
                c2.set("bin_split", 8)
                res, st = c2.count_fasta(cfg(k, m, 3, B, 0), text)               # sort path: one ascending list per bin
                assert_same(res.arrays(), oracle.count(text, k, m, 3, B, 0, threads=8), "sort path ignores bin_split")
    finally:
        c2.close()


# ---------------------------------------------------------------- the drop-in call and its files
def test_record_folding_gives_identical_counts(oracle):
    """fold_records=1 (hash path, k <= 32): identical super-k-mer records, either strand, are folded into one
    weighted record before counting.  Same (bin, k-mer, count) as the oracle; deep coverage folds a lot, an input
    without repeats is left alone after the sampled first batch; 128-bit k-mers and the sort path ignore the knob."""
    rng = random.Random(99)
    deep = fk.synth_fasta(dict(seeds=(71, 72, 73), genome_len=30000, n_reads=60000, read_len=100)).tobytes()     # 200x coverage
    flat = fk.synth_fasta(dict(seeds=(74, 75, 76), genome_len=40000000, n_reads=60000, read_len=100)).tobytes()  # 0.15x: no repeats
    poly = (">a\n" + "A" * 5000 + "\n>t\n" + "T" * 5000 + "\n>ac\n" + "AC" * 3000 + "\n>r\n" +
            "".join(rng.choice("ACGT") for _ in range(3000)) * 3 + "\n").encode()
    c2 = fk.Context(0)
    try:
        c2.set("count_mode", 0)                               # folding belongs to the global-table pipeline
        c2.set("fold_records", 1)
        for text, label in ((deep, "deep"), (poly, "poly"), (flat, "flat")):
            for k, m, B in ((28, 10, 2048), (31, 11, 64), (12, 4, 100), (32, 7, 1)):
                want = oracle.count(text, k, m, 3, B, 1, threads=8)
                for tbl in (1 << 30, 1 << 16):                     # one fold batch / many small ones (sample + asynchronous rest)
                    c2.set("fold_table_bytes", tbl)
                    res, st = c2.count_fasta(cfg(k, m, 3, B, 1), text)
                    assert_same(res.sorted_arrays(), want, "fold %s k=%d tbl=%d" % (label, k, tbl))
                    assert (st["digest_sum"], st["digest_xor"]) == (want["stats"]["digest_sum"], want["stats"]["digest_xor"])
                    if st["n_superkmers"] >= 4096:
                        if label == "deep" and tbl == 1 << 30:          # (small tables: sized from a one-bin sample, may give up)
                            assert 0 < st["n_folded_records"] < 0.6 * st["n_superkmers"]
                        if label == "flat" and tbl == 1 << 16 and B > 64:
                            assert st["n_folded_records"] == 0      # the sample said: not worth it
        c2.set("fold_table_bytes", 1 << 30)
        want = oracle.count(deep, 55, 13, 3, 2048, 1, threads=8)
        res, st = c2.count_fasta(cfg(55, 13, 3, 2048, 1), deep)
        assert_same(res.sorted_arrays(), want, "fold knob, 128-bit k-mers")
        assert st["n_folded_records"] == 0
        want = oracle.count(deep, 28, 10, 3, 2048, 0, threads=8)
        res, st = c2.count_fasta(cfg(28, 10, 3, 2048, 0), deep)
        assert_same(res.arrays(), want, "fold knob, sort path")
        assert st["n_folded_records"] == 0
    finally:
        c2.close()


@pytest.mark.parametrize("ht", [0, 1])
def test_execute_job_writes_reference_layout(ctx, oracle, tmp_path, ht):
    fasta = oracle.gen_lcg_fasta(42, 2000, 200, 100)
    inp = tmp_path / "reads.fasta"
    inp.write_bytes(fasta)
    tc = fk.TestConfiguration(str(inp), str(tmp_path) + "/out/", 28, 10, 3, max_b=2048, prefix="run_", useHT=bool(ht), write=True)
    st = fk.SparkBinKmerCounter.executeJob(ctx, tc)
    out = tmp_path / "out" / "run_k28_m10_x3_b2048_s0"                    # test/package.scala:33
    assert out.is_dir()
    want = oracle.count(fasta, 28, 10, 3, 2048, ht)
    by_bin = {}
    for b, h, l, c in zip(want["bin"], want["hi"], want["lo"], want["cnt"]):
        by_bin.setdefault(int(b), []).append(b"%s\t%d\n" % (oracle_lib.kmer_str(h, l, 28).encode(), int(c)))
    assert sorted(os.listdir(out)) == sorted("bin%d" % b for b in by_bin)  # one file per non-empty bin
    for b, lines in by_bin.items():
        data = (out / ("bin%d" % b)).read_bytes()
        if ht:                                                             # hash order, no trailer (SBKC:723-734)
            assert not data.endswith(b"EOF")
            assert sorted(data.splitlines(keepends=True)) == sorted(lines)
        else:                                                              # ascending, then "EOF" without newline (SBKC:598-606)
            assert data == b"".join(lines) + b"EOF"
    assert st["n_kmers"] == 10628
    # write=0 creates nothing (lazy writers, SBKC:552-554)
    tc0 = fk.TestConfiguration(str(inp), str(tmp_path) + "/none/", 28, 10, 3, max_b=2048, useHT=bool(ht), write=False)
    fk.LocalTestKmerCounter.run(tc0)
    assert not (tmp_path / "none").exists()


def test_debug_configuration_writes_to_the_debug_directory(ctx, oracle, tmp_path, monkeypatch):
    """TestConfiguration.debug moves the output to debugDirectory + stem, without the _s<type> suffix (test/package.scala:33)."""
    from fastkmer_b200 import config
    monkeypatch.setattr(config, "debugDirectory", str(tmp_path) + "/dbg_")
    fasta = oracle.gen_lcg_fasta(7, 3000, 150, 90)
    inp = tmp_path / "r.fasta"
    inp.write_bytes(fasta)
    tc = fk.TestConfiguration(str(inp), str(tmp_path) + "/out/", 20, 6, 3, max_b=64, useHT=True, write=True, debug=True)
    st = fk.SparkBinKmerCounter.executeJob(ctx, tc)
    assert tc.outputDir == str(tmp_path) + "/dbg_k20_m6_x3_b64" and os.path.isdir(tc.outputDir) and not (tmp_path / "out").exists()
    want = oracle.count(fasta, 20, 6, 3, 64, 1)
    assert st["n_kmers"] == want["stats"]["n_kmers"] and len(os.listdir(tc.outputDir)) == np.unique(want["bin"]).size


@pytest.mark.parametrize("k,m,ht", [(28, 10, 0), (28, 10, 1), (55, 13, 1)])
def test_execute_job_on_several_gpus(oracle, tmp_path, k, m, ht):
    """fkm_execute_job_multi: the drop-in job over the GPUs of the node through the C ABI alone (byte ranges of the file, LPT bin owners,
    GPU-to-GPU copies): the oracle's bin files.  With one visible GPU the same device is used twice (the exchange is then a local copy)."""
    import torch
    from fastkmer_b200 import api
    n_dev = torch.cuda.device_count()
    devices = [0, 1] if n_dev >= 2 else [0, 0]
    fasta = fk.synth_fasta(dict(seeds=(91, 92, 93), genome_len=40000, n_reads=20000, read_len=120)).tobytes()
    inp = tmp_path / "reads.fasta"
    inp.write_bytes(fasta)
    tc = fk.TestConfiguration(str(inp), str(tmp_path) + "/out/", k, m, 3, max_b=512, prefix="mg_", useHT=bool(ht), write=True)
    st = api.Context.execute_job_multi(tc, devices)
    want = oracle.count(fasta, k, m, 3, 512, ht, threads=8)
    assert (st["n_kmers"], st["n_distinct"], st["total_count"]) == (want["stats"]["n_kmers"], want["stats"]["n_distinct"], want["stats"]["n_kmers"])
    assert (st["digest_sum"], st["digest_xor"]) == (want["stats"]["digest_sum"], want["stats"]["digest_xor"])
    out = tmp_path / "out" / ("mg_k%d_m%d_x3_b512_s0" % (k, m))
    by_bin = {}
    for b, h, l, c in zip(want["bin"], want["hi"], want["lo"], want["cnt"]):
        by_bin.setdefault(int(b), []).append(b"%s\t%d\n" % (oracle_lib.kmer_str(h, l, k).encode(), int(c)))
    assert sorted(os.listdir(out)) == sorted("bin%d" % b for b in by_bin)
    for b, lines in by_bin.items():
        data = (out / ("bin%d" % b)).read_bytes()
        if ht:
            assert sorted(data.splitlines(keepends=True)) == sorted(lines)
        else:
            assert data == b"".join(lines) + b"EOF"


@pytest.mark.parametrize("k,m,ht", [(28, 10, 0), (55, 13, 1), (33, 9, 0)])
def test_device_formatted_files_at_depth(ctx, oracle, tmp_path, k, m, ht):
    """Counts with 1-3 digits, 64- and 128-bit k-mers, several bins per device pass: every file byte-identical
    to the reference layout built from the oracle."""
    spec = dict(seeds=(71, 72, 73), genome_len=3000, n_reads=20000, read_len=150)       # 1000x coverage
    fasta = fk.synth_fasta(spec).tobytes()
    want = oracle.count(fasta, k, m, 3, 512, ht, threads=8)
    assert int(want["cnt"].max()) >= 100
    res, st = ctx.count_fasta(cfg(k, m, 3, 512, ht), fasta)
    out = tmp_path / "o"
    res.write(str(out))
    by_bin = {}
    for b, h, l, c in zip(want["bin"], want["hi"], want["lo"], want["cnt"]):
        by_bin.setdefault(int(b), []).append(b"%s\t%d\n" % (oracle_lib.kmer_str(h, l, k).encode(), int(c)))
    assert sorted(os.listdir(out)) == sorted("bin%d" % b for b in by_bin)
    for b, lines in by_bin.items():
        data = (out / ("bin%d" % b)).read_bytes()
        if ht:
            assert sorted(data.splitlines(keepends=True)) == sorted(lines)
        else:
            assert data == b"".join(lines) + b"EOF"


def test_cli_matches_reference_argv(oracle, tmp_path):
    fasta = oracle.gen_lcg_fasta(45, 300, 60, 40)
    inp = tmp_path / "g4.fasta"
    inp.write_bytes(fasta)
    cli = os.path.join(ROOT, "fastkmer_b200", "fastkmer_cli")
    r = subprocess.run([cli, "5", "3", "1", "64", "0", "0", str(inp), str(tmp_path) + "/", "g4", "1", "0", "0"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = tmp_path / "g4k5_m3_x1_b64_s0"
    first = (out / "bin0").read_bytes().splitlines()
    assert first[:5] == [b"CTGAC\t3", b"CTGCA\t10", b"CTGGA\t13", b"CTGTC\t18", b"GGAAC\t16"]     # SURVEY App. C.4 anchor
    lines = []
    for name in os.listdir(out):
        b = int(name[3:])
        body = (out / name).read_bytes()
        assert body.endswith(b"EOF")
        lines += [b"%d\t%s\n" % (b, ln) for ln in body[:-3].splitlines()]
    assert hashlib.sha256(b"".join(sorted(lines))).hexdigest() == GOLD["G4"][-1]


# ---------------------------------------------------------------- device FASTA ingest == host packer, bit for bit
INGEST_TEXTS = [
    b"", b"no header here\nACGT\n", b">only header", b">h\n", b">a\nACGTN\nAC\n>b x\nTTTT\n", b">a\r\nAC\r\nGT\r\n",
    b"junk\n>a\nACGTacgtRYKM\n\n\nAC GT\n>b\n>c\nA", b">x\n" + b"ACGT" * 40 + b"\n>y\n" + b"TTGCA" * 13, b">a\nAC>GT\n>b\nGG\n",
    b"\n\n>a\n\nAC\n\n", b">a\n" + b"ACGT" * 5000,                       # one line longer than a tile, no final newline
    b">" + b"h" * 20000 + b"\nACGT\n>b\nGGCC\n",                       # header longer than two tiles
]


@pytest.mark.parametrize("idx", range(len(INGEST_TEXTS)))
def test_device_ingest_equals_host_pack(ctx, idx):
    text = INGEST_TEXTS[idx]
    hb, hi, hn, hbases = fk.pack_fasta(text)
    db, di, dn, dbases = ctx.pack_fasta_device(text)
    assert (dn, dbases) == (hn, hbases)
    assert np.array_equal(db, hb) and np.array_equal(di, hi)


def test_device_ingest_random_texts(ctx):
    rng = random.Random(99)
    for trial in range(6):
        parts = []
        if trial % 2:
            parts.append("leading junk\nmore junk\n")
        for r in range(rng.randint(1, 400)):
            L = rng.choice([0, 1, 31, 32, 33, 100, 150, 1000, 9000])
            s = "".join(rng.choice("ACGTNacgt>\r ") if rng.random() < 0.03 else rng.choice("ACGT") for _ in range(L))
            width = rng.choice([None, 60, 70, 7])
            if width:
                s = "\n".join(s[i:i + width] for i in range(0, len(s), width))
            parts.append(">r%d %s\n%s%s" % (r, "x" * rng.randint(0, 80), s, "\n" if rng.random() < 0.95 else ""))
            if not parts[-1].endswith("\n"):
                parts[-1] += "\n"
        text = "".join(parts).encode()
        hb, hi, hn, hbases = fk.pack_fasta(text)
        db, di, dn, dbases = ctx.pack_fasta_device(text)
        assert (dn, dbases) == (hn, hbases)
        assert np.array_equal(db, hb) and np.array_equal(di, hi)


def test_sort_path_variants(oracle):
    """useHT=0: sub-buckets by the top key bits + shared-memory chunk sort (default: the partition kernels of fkm_part.cuh; knob 2: the
    older MSD kernels), the no-partition case (small bins), the LSD fallback when one sub-bucket is larger than a chunk, and forced
    LSD passes all give the reference's ordered output."""
    rng = random.Random(77)
    spec = dict(seeds=(61, 62, 63), genome_len=200000, n_reads=30000, read_len=150)
    reads = fk.synth_fasta(spec).tobytes()
    skew = (">poly\n" + "A" * 60000 + "\n>rnd\n" + "".join(rng.choice("ACGT") for _ in range(50000)) + "\n").encode()
    c2 = fk.Context(0)
    try:
        for text, B, expect_fallback in ((reads, 2048, False), (reads, 4, False), (reads, 1, False), (skew, 64, True)):
            for k, m in ((28, 10), (55, 13), (31, 11)):
                want = oracle.count(text, k, m, 3, B, 0, threads=8)
                c2.set("debug_force_lsd", 2)
                res, st = c2.count_fasta(cfg(k, m, 3, B, 0), text)
                assert_same(res.arrays(), want, "msd B=%d k=%d" % (B, k))          # already in file order
                assert (st["n_fallbacks"] > 0) == expect_fallback
                c2.set("debug_force_lsd", 0)
                res, st = c2.count_fasta(cfg(k, m, 3, B, 0), text)
                assert_same(res.arrays(), want, "auto B=%d k=%d" % (B, k))          # expanded once, sub-buckets by the top key bits (fkm_part.cuh kernels)
                assert (st["n_fallbacks"] > 0) == expect_fallback
                c2.set("sort_partition", 2)                                        # the same sub-buckets, counted in ordered shared-memory tables
                res, st = c2.count_fasta(cfg(k, m, 3, B, 0), text)                 # (or, when a key range is too dense for the table, by the kernels above)
                assert_same(res.arrays(), want, "ordered tables B=%d k=%d" % (B, k))
                c2.set("sort_partition", 1)
                c2.set("debug_force_lsd", 1)
                res, st = c2.count_fasta(cfg(k, m, 3, B, 0), text)
                assert_same(res.arrays(), want, "lsd B=%d k=%d" % (B, k))
    finally:
        c2.close()


def test_streamed_fasta_chunks(oracle):
    """FASTA text is copied and parsed in chunks cut at record boundaries: any chunk size gives the same counts."""
    rng = random.Random(123)
    fasta = _rand_fasta(rng, 600, 0, 400, width=60).encode()
    long_rec = (">one long record\n" + "\n".join("".join(rng.choice("ACGT") for _ in range(70)) for _ in range(300)) + "\n").encode()
    c2 = fk.Context(0)
    try:
        for text in (fasta, b"junk before\n" + fasta, long_rec, fasta + long_rec + fasta):
            for ht, k, m in ((1, 28, 10), (0, 31, 11), (1, 55, 13)):
                want = oracle.count(text, k, m, 3, 2048, ht, threads=8)
                for chunk in (4096, 20000, 1 << 30):
                    c2.set("ingest_chunk_bytes", chunk)
                    res, st = c2.count_fasta(cfg(k, m, 3, 2048, ht), text)
                    assert_same(res.sorted_arrays(), want, "chunk %d" % chunk)
                    assert st["n_bases"] == want["stats"]["n_bases"] and st["n_kmers"] == want["stats"]["n_kmers"]
    finally:
        c2.close()


def test_speculative_scatter_under_the_copy(oracle):
    """Hash path, FASTA input in several chunks: after a quarter of the chunks the bins' sizes are forecast and every scanned chunk is
    scattered at once (bin regions with slack; k_expand_hist reads around the gaps).  Same result as the oracle with the forecast
    in use, with it switched off, and when a bin outgrows its region (the first chunks hold one read over and over: they say nothing
    about the bins of the later ones), where the job scatters again with the exact offsets."""
    spec = dict(seeds=(101, 102, 103), genome_len=400000, n_reads=60000, read_len=150)
    shuffled = fk.synth_fasta(spec).tobytes()
    rng = random.Random(7)
    one = "".join(rng.choice("ACGT") for _ in range(150))
    skewed = ("".join(">same%d\n%s\n" % (i, one) for i in range(100000)).encode() +           # the first third of the text fills a handful of bins ...
              fk.synth_fasta(dict(seeds=(104, 105, 106), genome_len=2000000, n_reads=200000, read_len=150)).tobytes())   # ... the rest all of them
    c2 = fk.Context(0)
    try:
        c2.set("ingest_chunk_bytes", 400000)
        for text, label in ((shuffled, "shuffled"), (skewed, "skewed")):
            for k, m in ((28, 10), (55, 13)):
                want = oracle.count(text, k, m, 3, 256, 1, threads=8)
                rb = 16 if k <= 32 else 32
                for on in (1, 0):
                    c2.set("speculative_scatter", on)
                    res, st = c2.count_fasta(cfg(k, m, 3, 256, 1), text)
                    what = "%s k=%d speculative %d" % (label, k, on)
                    assert_same(res.sorted_arrays(), want, what)
                    assert (st["digest_sum"], st["digest_xor"], st["n_kmers"]) == (want["stats"]["digest_sum"], want["stats"]["digest_xor"], want["stats"]["n_kmers"]), what
                    used = st["superkmer_bytes"] > st["n_superkmers"] * rb           # regions with slack
                    assert used == (on == 1 and label == "shuffled"), what
                c2.set("speculative_scatter", 1)
                c2.set("bin_split", 4)                                               # the forecast is per internal bin
                res, st = c2.count_fasta(cfg(k, m, 3, 256, 1), text)
                assert_same(res.sorted_arrays(), want, "%s k=%d speculative, 4 internal bins per bin" % (label, k))
                c2.set("bin_split", 0)
    finally:
        c2.close()


# ---------------------------------------------------------------- multi-GPU stages, emulated rank by rank on one GPU
@pytest.mark.parametrize("world,k,m,ht", [(2, 28, 10, 1), (3, 28, 10, 0), (4, 55, 13, 1), (2, 55, 13, 0)])
def test_multigpu_stages_emulated(ctx, oracle, world, k, m, ht):
    from fastkmer_b200 import multigpu
    reads, L = 6000, 150
    spec = dict(seeds=(31, 32, 33), genome_len=60000, n_reads=reads, read_len=L)
    fasta = fk.synth_fasta(spec).tobytes()
    want = oracle.count(fasta, k, m, 3, 2048, ht, threads=8)
    per = reads // world
    shards, bufs = [], []
    for r in range(world):
        n = per if r < world - 1 else reads - per * (world - 1)
        sh = ctx.synth_packed_device(dict(spec, n_reads=n, first_read=r * per))
        shards.append(sh)
    c = cfg(k, m, 3, 2048, ht)
    got, stats, plans = multigpu.emulate_ranks(ctx, c, shards, world)
    assert_same(got, want, "emulated %d ranks" % world)
    assert sum(s["n_kmers"] for s in stats) == want["stats"]["n_kmers"]
    dsum = sum(s["digest_sum"] for s in stats) & ((1 << 64) - 1)
    dxor = 0
    for s in stats:
        dxor ^= s["digest_xor"]
    assert (dsum, dxor) == (want["stats"]["digest_sum"], want["stats"]["digest_xor"])
    owners = plans[0]["owner"]
    for r, s in enumerate(stats):                            # every rank counted only bins it owns
        assert s["n_kmers"] == int(plans[r]["bin_kmer"].sum())
    assert len(set(owners.tolist())) == world
    for sh in shards:
        ctx.free_device(sh[0])
        ctx.free_device(sh[1])


# ---------------------------------------------------------------- multi-sample distances (BASELINE config 5 shape)
def _multiseq_input(n_samples, reads_per_sample, L, seed):
    """Samples = a common ancestor with per-sample substitutions; reads of all samples interleaved, header 'S<s>.<r> ...'."""
    rng = random.Random(seed)
    anc = [rng.choice("ACGT") for _ in range(8000)]
    recs = []
    for s_ in range(n_samples):
        g = list(anc)
        for i in range(len(g)):
            if rng.random() < 0.02:
                g[i] = rng.choice("ACGT")
        for r in range(reads_per_sample):
            p = rng.randrange(0, len(g) - L)
            read = "".join(g[p:p + L])
            if rng.random() < 0.5:
                read = clean_spec.revcomp(read)
            if rng.random() < 0.1:
                q = rng.randrange(L)
                read = read[:q] + "N" + read[q + 1:]
            recs.append((rng.random(), ">S%d.%d len=%d\n%s\n" % (s_, r, L, read)))
    recs.sort()
    return "".join(t for _, t in recs)


def _multiseq_expected(oracle, fasta, k, m, B):
    import re
    texts, order = {}, []
    for rec in re.split(r"(?m)^(?=>)", fasta):
        if not rec:
            continue
        tag = re.match(r">\W*(\w+)", rec).group(1)
        if tag not in texts:
            texts[tag] = []
            order.append(tag)
        texts[tag].append(rec)
    counts = []
    for tag in order:
        res = oracle.count("".join(texts[tag]).encode(), k, m, 3, B, 1, threads=8)
        counts.append({(int(b), int(h), int(l)): int(c) for b, h, l, c in zip(res["bin"], res["hi"], res["lo"], res["cnt"])})
    S = len(order)
    keys = set().union(*[set(c) for c in counts]) if counts else set()
    dist = np.zeros((S, S))
    for a in range(S):
        for b in range(a + 1, S):
            d = sum((counts[a].get(key, 0) - counts[b].get(key, 0)) ** 2 for key in keys)   # SquaredEuclidean.java:19-27
            dist[a, b] = dist[b, a] = d
    return order, dist


@pytest.mark.parametrize("k,m", [(28, 10), (55, 13), (12, 5)])
def test_multisequence_distances(ctx, oracle, tmp_path, k, m):
    fasta = _multiseq_input(5, 300, 100, 4242 + k)
    c = cfg(k, m, 3, 2048, 0)
    names, dist, res, st = ctx.multiseq_fasta(c, fasta.encode(), want_result=True)
    want_names, want_dist = _multiseq_expected(oracle, fasta, k, m, 2048)
    assert names == want_names
    assert np.array_equal(dist, want_dist) and dist.max() > 0          # integer-valued doubles: exact
    whole = oracle.count(fasta.encode(), k, m, 3, 2048, 0, threads=8)    # the files hold kmer<TAB>sum of the per-sample counts
    assert_same(res.arrays(), whole, "merged counts")
    res.write(str(tmp_path / "ms"))
    some = sorted(os.listdir(tmp_path / "ms"))[0]
    body = (tmp_path / "ms" / some).read_bytes()
    assert body.endswith(b"\n") and not body.endswith(b"EOF")             # no trailer in this writer (MSKC:524-526)


def test_multisequence_mirror(ctx, tmp_path):
    from fastkmer_b200.multisequence import MultisequenceTestConfiguration, SparkMultiSequenceKmerCounter
    fasta = _multiseq_input(3, 100, 80, 99)
    inp = tmp_path / "samples.fasta"
    inp.write_text(fasta)
    tc = MultisequenceTestConfiguration(str(inp), str(tmp_path) + "/", 20, 6, 3, max_b=100, write=True)
    names, dist = SparkMultiSequenceKmerCounter.executeJob(ctx, tc)
    assert names == ["S%d" % i for i in sorted(range(3), key=lambda s_: fasta.index(">S%d." % s_))]
    assert dist.shape == (3, 3) and (dist == dist.T).all() and (np.diag(dist) == 0).all()
    assert (tmp_path / "k20_m6_x3_b100_s0").is_dir()                      # multisequence/package.scala:29


def test_result_clone_survives_later_jobs_and_dot(ctx, oracle):
    fa = _multiseq_input(1, 400, 100, 7).encode()
    fb = _multiseq_input(1, 400, 100, 8).encode()
    c = cfg(28, 10, 3, 2048, 0)
    ra, _ = ctx.count_fasta(c, fa)
    ca = ra.clone()
    rb, _ = ctx.count_fasta(c, fb)                      # a later job on the same context: ra is stale, its clone is not
    cb = rb.clone()
    with pytest.raises(fk.FkmError):
        ra._cache = None
        ra.arrays()
    wa = oracle.count(fa, 28, 10, 3, 2048, 0)
    wb = oracle.count(fb, 28, 10, 3, 2048, 0)
    assert_same(ca.arrays(), wa, "clone a")
    da = {(int(b), int(l)): int(n) for b, l, n in zip(wa["bin"], wa["lo"], wa["cnt"])}
    db = {(int(b), int(l)): int(n) for b, l, n in zip(wb["bin"], wb["lo"], wb["cnt"])}
    assert ca.dot(cb) == sum(n * db.get(key, 0) for key, n in da.items()) == cb.dot(ca)
    assert ca.dot(ca) == sum(n * n for n in da.values())
    ca.free()
    cb.free()


# ---------------------------------------------------------------- N ranks under torchrun / NCCL against the oracle
def test_multigpu_ranks_match_oracle_under_torchrun():
    """scripts/mg_check.py under `python -m torch.distributed.run` with every GPU of the box (at most 8): each rank counts its shard
    through ShardedJob (NCCL all-to-all), rank 0 merges the per-rank results and compares them with the CPU oracle — short reads on
    both count paths, 64- and 128-bit k-mers, and a halo-sharded long sequence.  With one GPU the same script runs as a single rank
    (no exchange partner, same code path)."""
    import sys
    import torch
    n = min(8, torch.cuda.device_count())
    world = 1 if n < 2 else (8 if n >= 8 else 4 if n >= 4 else 2)
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "mg_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("mg_check")]
    assert len(lines) == 6 and all(ln.split(":")[1].strip().startswith("OK") for ln in lines), r.stdout[-3000:]
