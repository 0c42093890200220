"""CPU tests: the C-ABI library loads and exports every symbol of include/fastkmer_b200.h,
host-side packing / synthetic generator / configuration mirror behave as specified.
No kernel runs here."""
import ctypes
import os
import re

import numpy as np
import pytest

import clean_spec
import fastkmer_b200 as fk
from fastkmer_b200 import api
from fastkmer_b200.counter import _parse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "fastkmer_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(fkm_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 25
    lib = ctypes.CDLL(fk.lib_path())
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in the header but not exported" % n


def test_no_cpu_fallback():
    lib = fk.load_library()
    h = ctypes.c_void_p()
    rc = lib.fkm_ctx_create(-1, None, ctypes.byref(h))
    if rc == 0:                               # a GPU is present (GPU box): nothing to check here
        lib.fkm_ctx_destroy(h)
        pytest.skip("CUDA device present")
    assert rc == api.FKM_ECUDA
    assert b"no CPU fallback" in lib.fkm_last_error()
    with pytest.raises(fk.FkmError):
        fk.Context()


def _py_pack(fasta: str):
    bits, inv = [], []
    for rec in clean_spec.records(fasta):
        for c in rec:
            bits.append("ACGT".index(c) if c in "ACGT" else 0)
            inv.append(0 if c in "ACGT" else 1)
        bits.append(0)
        inv.append(1)
    return bits, inv


@pytest.mark.parametrize("text", [
    "", "no header here\nACGT\n", ">only header", ">h\n", ">a\nACGTN\nAC\n>b x\nTTTT\n", ">a\r\nAC\r\nGT\r\n",
    "junk\n>a\nACGTacgtRYKM\n\n\nAC GT\n>b\n>c\nA", ">x\n" + "ACGT" * 40 + "\n>y\n" + "TTGCA" * 13,
    ">a\nAC>GT\n>b\nGG\n",
])
def test_pack_fasta_matches_record_spec(text):
    bases, inv, n_pos, n_bases = fk.pack_fasta(text.encode())
    bits, flags = _py_pack(text)
    assert n_pos == len(bits)
    assert n_bases == sum(len(r) for r in clean_spec.records(text))
    for p in range(n_pos):
        assert (int(bases[p >> 5]) >> (62 - 2 * (p & 31))) & 3 == bits[p]
        assert (int(inv[p >> 5]) >> (31 - (p & 31))) & 1 == flags[p]
    for p in range(n_pos, ((n_pos + 31) // 32) * 32):          # tail of the last word is invalid
        assert (int(inv[p >> 5]) >> (31 - (p & 31))) & 1 == 1


M64 = (1 << 64) - 1


def _sm(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


def py_synth(seeds, G, R, L, first=0):
    """SURVEY §8(d) generator, pure Python."""
    sg, sr, se = seeds
    out = []
    for r in range(first, first + R):
        pos = _sm(sr + 2 * r) % (G - L + 1)
        strand = _sm(sr + 2 * r + 1) & 1
        s = []
        for j in range(L):
            gi = pos + (L - 1 - j) if strand else pos + j
            b = _sm((sg + gi) & M64) >> 62
            if strand:
                b = 3 - b
            e = _sm((se + r * L + j) & M64)
            if e % 1000 == 0:
                s.append("N")
                continue
            if e % 100 == 1:
                b = (b + 1 + ((e >> 32) % 3)) % 4
            s.append("ACGT"[b])
        out.append(">r%d\n%s\n" % (r, "".join(s)))
    return "".join(out)


def test_synth_fasta_matches_spec():
    spec = dict(seeds=(1001, 1002, 1003), genome_len=5000, n_reads=137, read_len=100)
    got = fk.synth_fasta(spec).tobytes().decode()
    assert got == py_synth((1001, 1002, 1003), 5000, 137, 100)
    spec2 = dict(spec, n_reads=20, first_read=95)              # shard of the same read set, digit-count boundary inside
    assert fk.synth_fasta(spec2).tobytes().decode() == py_synth((1001, 1002, 1003), 5000, 20, 100, first=95)
    assert "N" in got


def test_configuration_mirror():
    tc = fk.TestConfiguration("in.fa", "/out/", 28, 10, 3, max_b=2048, prefix="pfx")
    assert tc.b == 2048 and tc.outputDir == "/out/pfxk28_m10_x3_b2048_s0"          # test/package.scala:32-33
    assert fk.TestConfiguration("i", "o", 5, 3, 1, max_b=2048, sequenceType=1).outputDir == "ok5_m3_x1_b64_s1"
    assert fk.TestConfiguration("i", "o", 5, 3, 1, debug=True, prefix="p").outputDir == "/tmp/pk5_m3_x1_b64"
    assert api.derive(tc) == (2048, "/out/pfxk28_m10_x3_b2048_s0")
    assert "Using HT:  false" in str(tc)


def test_argv_order_is_the_codes_not_the_readmes():
    # LocalTestKmerCounter.scala:35-48: k m x B useHT sequenceType input output prefix write kryo customPart [tasks]
    tc = _parse("28 10 3 2048 1 0 in.fa /o/ pfx 1 0 0".split())
    assert (tc.k, tc.m, tc.x, tc.max_b, tc.useHT, tc.sequenceType) == (28, 10, 3, 2048, True, 0)
    assert (tc.dataset, tc.outputDirectory, tc.prefix, tc.write, tc.useKryoSerializer) == ("in.fa", "/o/", "pfx", True, False)
    tc = _parse("28 10 3 2048 0 1 a b c 0 1 1 7".split())
    assert tc.useCustomPartitioner and tc.numPartitionTasks == 7 and tc.useKryoSerializer and not tc.write
    with pytest.raises(IndexError):
        _parse(["28", "10"])


@pytest.mark.parametrize("bad", [dict(m=2), dict(m=16), dict(k=65), dict(k=9, m=10), dict(max_b=0), dict(x=0, useHT=False)])
def test_rejected_configurations(bad):
    kw = dict(k=28, m=10, x=3, max_b=2048, useHT=True)
    kw.update(bad)
    tc = fk.TestConfiguration("i", "o", kw.pop("k"), kw.pop("m"), kw.pop("x"), **kw)
    with pytest.raises(fk.FkmError) as e:
        api.derive(tc)
    assert e.value.code == api.FKM_EINVAL


def test_header_is_plain_c_and_ctypes_mirrors_match(tmp_path):
    """include/fastkmer_b200.h compiles as C (what cgo / JNI would include) and the ctypes mirrors in api.py have the
    same size and field offsets as the C structs."""
    import ctypes
    import subprocess
    from fastkmer_b200 import api
    mirrors = {"fkm_config": api.fkm_config, "fkm_stats": api.fkm_stats, "fkm_synth": api.fkm_synth, "fkm_synth_long": api.fkm_synth_long}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "fastkmer_b200.h"', 'int main(void) {']
    for name, cls in mirrors.items():
        lines.append('  printf("%s.sizeof %%zu\\n", sizeof(%s));' % (name, name))
        for field, _ in cls._fields_:
            lines.append('  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, field, name, field))
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for name, cls in mirrors.items():
        assert int(got[name + ".sizeof"]) == ctypes.sizeof(cls), name
        for field, _ in cls._fields_:
            assert int(got["%s.%s" % (name, field)]) == getattr(cls, field).offset, "%s.%s" % (name, field)


def test_jni_shim_builds_and_reports_errors(tmp_path):
    """integration/fkm_jni.c (the reference-side binding of INTEGRATION.md) compiles against a minimal stand-in for
    <jni.h>, links the C-ABI library and, driven through a fake JNIEnv, returns the library's error code and message
    (no CUDA device here; on a GPU box the missing dataset is the error) and releases every string it took."""
    import subprocess
    exe = tmp_path / "jni_harness"
    libdir = os.path.join(ROOT, "fastkmer_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "integration", "fkm_jni.c"), os.path.join(ROOT, "tests", "jni_harness.c"),
                           "-L", libdir, "-lfastkmer_b200", "-Wl,-rpath," + libdir, "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120).stdout.strip()
    rc = int(out.split()[0].split("=")[1])
    assert rc != 0 and "live=0" in out and len(out.split("err=", 1)[1]) > 0, out


def test_threaded_packer_equals_the_plain_loop():
    """fkm_pack_fasta_mt cuts the text at record starts, counts, packs the ranges in place and merges the words two ranges share:
    bit-identical to the one-thread loop for any thread count, on ragged records, junk before the first header, CRLF, no final newline."""
    import ctypes as C
    import random
    import numpy as np
    from fastkmer_b200 import api
    lib = api.load_library()

    def pack(text, threads):
        arr = np.frombuffer(text, dtype=np.uint8)
        n_pos, n_bases = C.c_uint64(), C.c_uint64()
        assert lib.fkm_pack_fasta_mt(arr.ctypes.data, arr.size, None, None, 0, C.byref(n_pos), C.byref(n_bases), threads) == 0
        nw = (n_pos.value + 31) // 32
        bases = np.full(max(nw, 1), 0xDEADBEEF, dtype=np.uint64); inv = np.full(max(nw, 1), 0xABCD, dtype=np.uint32)
        assert lib.fkm_pack_fasta_mt(arr.ctypes.data, arr.size, bases.ctypes.data, inv.ctypes.data, nw * 32, C.byref(n_pos), C.byref(n_bases), threads) == 0
        return bases[:nw].tolist(), inv[:nw].tolist(), n_pos.value, n_bases.value

    rng = random.Random(2026)
    for case in range(6):
        recs = []
        for r in range(rng.choice([3, 50, 4000, 20000])):
            n = rng.choice([0, 1, 31, 32, 33, 100, 150, 1000]) if case % 2 else rng.randrange(0, 300)
            seq = "".join(rng.choice("ACGTNacgt-" if case == 3 else "ACGT") for _ in range(n))
            width = rng.choice([0, 60, 70])
            body = seq if not width else "\n".join(seq[i:i + width] for i in range(0, len(seq), width))
            recs.append(">r%d some text > inside\n%s%s" % (r, body, "\r\n" if case == 4 else "\n"))
        text = ("junk before the first header\nACGT\n" if case >= 2 else "") + "".join(recs)
        if case == 5:
            text = text.rstrip("\n")
        text = text.encode()
        want = pack(text, 1)
        for threads in (2, 3, 7, 16, 0):
            assert pack(text, threads) == want, (case, threads)
    assert pack(b"", 4) == ([], [], 0, 0) and pack(b"no header at all\nACGT\n" * 10000, 4) == ([], [], 0, 0)
