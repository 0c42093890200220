"""Mirrors of the reference's multisequence prototype (skc.multisequence): k-mer based squared-euclidean
distances between the samples of one FASTA file.

    MultisequenceTestConfiguration     multisequence/package.scala:25-34
    SparkMultiSequenceKmerCounter.executeJob   multisequence/SparkMultiSequenceKmerCounter.scala:549-591
    TestMultisequenceKmerCounter.main  multisequence/TestMultisequenceKmerCounter.scala:12-104

The reference job never runs (its RDD has no action, SparkMultiSequenceKmerCounter.scala:585-588); the semantics
implemented are the intended ones: a read's sample is the leading \\w+ of its header, every pair of
samples gets sum over distinct canonical k-mers of (c_a - c_b)^2 (multiseq/SquaredEuclidean.java:19-32), and the bin
files hold `kmer<TAB>sum of counts` in ascending order.
"""
import sys
from dataclasses import dataclass

from .api import Context
from .config import TestConfiguration


@dataclass
class MultisequenceTestConfiguration:                  # multisequence/package.scala:25-34
    dataset: str
    outputDirectory: str
    k: int
    m: int
    x: int
    max_b: int = 2000
    sequenceType: int = 0
    canonical: bool = True
    debug: bool = False
    write: bool = True
    useCustomPartitioner: bool = False
    numPartitionTasks: int = 0

    @property
    def b(self):
        return int(min(4 ** self.m, self.max_b))

    @property
    def outputDir(self):                               # package.scala:29 — no prefix in this variant
        return self.outputDirectory + "k%d_m%d_x%d_b%d_s%d" % (self.k, self.m, self.x, self.b, self.sequenceType)

    def counting_configuration(self):
        return TestConfiguration(self.dataset, self.outputDirectory, self.k, self.m, self.x, max_b=self.max_b,
                                 sequenceType=self.sequenceType, write=self.write, useHT=False)


class SparkMultiSequenceKmerCounter:
    @staticmethod
    def executeJob(spark, configuration: MultisequenceTestConfiguration, max_samples=64):
        """-> (sample names, distance matrix).  `spark` is a fastkmer_b200.Context (None: a temporary one)."""
        print("SparkMultiSequenceKmerCounter")
        ctx = spark if spark is not None else Context()
        try:
            with open(configuration.dataset, "rb") as f:
                fasta = f.read()
            names, dist, res, _ = ctx.multiseq_fasta(configuration.counting_configuration(), fasta, max_samples,
                                                     want_result=configuration.write)
            if configuration.write:
                res.write(configuration.outputDir)
            return names, dist
        finally:
            if spark is None:
                ctx.close()


class TestMultisequenceKmerCounter:
    __test__ = False

    @staticmethod
    def main(args):
        # same positional arguments as the k-mer counting mains (TestMultisequenceKmerCounter.scala:29-41)
        k, m, x, b = int(args[0]), int(args[1]), int(args[2]), int(args[3])
        sequenceType = int(args[5])
        write = int(args[9]) == 1
        useCustomPartitioner = int(args[11]) == 1
        numPartitionTasks = int(args[12]) if useCustomPartitioner else 0
        tc = MultisequenceTestConfiguration(args[6], args[7], k, m, x, max_b=b, sequenceType=sequenceType, write=write,
                                            useCustomPartitioner=useCustomPartitioner, numPartitionTasks=numPartitionTasks)
        return SparkMultiSequenceKmerCounter.executeJob(None, tc)


if __name__ == "__main__":
    names, dist = TestMultisequenceKmerCounter.main(sys.argv[1:])
    for i, a in enumerate(names):
        for j in range(i + 1, len(names)):
            print("%s\t%s\t%.1f" % (a, names[j], dist[i, j]))
