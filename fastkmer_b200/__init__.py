"""fastkmer_b200 — B200-native exact k-mer counting behind fastkmer's own entry points.

Host-side mirror of the reference's interface for this path (the JVM toolchain is not
in this image, so the mirror is Python over the C ABI of include/fastkmer_b200.h):

    TestConfiguration        skc.test.testutil.TestConfiguration     (test/package.scala:16-42)
    SparkBinKmerCounter      skc.SparkBinKmerCounter.executeJob      (SparkBinKmerCounter.scala:989)
    LocalTestKmerCounter     skc.test.LocalTestKmerCounter.main      (LocalTestKmerCounter.scala:18)
    TestKmerCounter          skc.test.TestKmerCounter.main           (TestKmerCounter.scala:15)

All compute runs in libfastkmer_b200.so (hand-written sm_100a kernels).  There is no CPU
fallback: importing works anywhere, creating a Context without a CUDA device raises.
"""
from .config import TestConfiguration
from .api import Context, CountResult, FkmError, Stats, lib_path, load_library, pack_fasta, synth_fasta, synth_long_fasta
from .counter import LocalTestKmerCounter, SparkBinKmerCounter, TestKmerCounter

__all__ = ["TestConfiguration", "Context", "CountResult", "FkmError", "Stats", "lib_path", "load_library",
           "pack_fasta", "synth_fasta", "synth_long_fasta", "SparkBinKmerCounter", "LocalTestKmerCounter", "TestKmerCounter"]
