"""Mirrors of the reference's job driver and runnable mains for the k-mer counting path.

SparkBinKmerCounter.executeJob(spark, configuration)   SparkBinKmerCounter.scala:989-1046
LocalTestKmerCounter.main(args) / TestKmerCounter.main(args)   LocalTestKmerCounter.scala:18-56 / TestKmerCounter.scala:15-54
"""
import sys

from .api import Context
from .config import TestConfiguration


class SparkBinKmerCounter:
    @staticmethod
    def executeJob(spark, configuration: TestConfiguration):
        """`spark` is a fastkmer_b200.Context standing where the SparkSession stood (None: a temporary one)."""
        print("SparkBinKmerCounter")                      # SparkBinKmerCounter.scala:997
        print(configuration)                               # :998
        ctx = spark if spark is not None else Context()
        try:
            return ctx.execute_job(configuration)
        finally:
            if spark is None:
                ctx.close()


def _parse(args):
    """Positional arguments in the CODE's order (LocalTestKmerCounter.scala:35-48):
    k m x B useHT sequenceType inputPath outputPath prefix write enableKryo useCustomPartitioner [numPartitionTasks]"""
    if len(args) < 12:
        raise IndexError("expected: k m x B useHT sequenceType inputPath outputPath prefix write enableKryo "
                         "useCustomPartitioner [numPartitionTasks]")       # the JVM raises ArrayIndexOutOfBounds
    k, m, x, b = int(args[0]), int(args[1]), int(args[2]), int(args[3])
    useHT = int(args[4]) == 1
    sequenceType = int(args[5])
    inputDatasetPath, outputDatasetPath, prefix = args[6], args[7], args[8]
    write = int(args[9]) == 1
    useKryo = int(args[10]) == 1
    useCustomPartitioner = int(args[11]) == 1
    numPartitionTasks = int(args[12]) if useCustomPartitioner else 0
    return TestConfiguration(inputDatasetPath, outputDatasetPath, k, m, x, max_b=b, prefix=prefix, useHT=useHT,
                             sequenceType=sequenceType, write=write, useCustomPartitioner=useCustomPartitioner,
                             numPartitionTasks=numPartitionTasks, useKryoSerializer=useKryo)


class LocalTestKmerCounter:
    @staticmethod
    def main(args):
        return LocalTestKmerCounter.run(_parse(args))

    @staticmethod
    def run(configuration: TestConfiguration):
        # the reference pins local[4] here (LocalTestKmerCounter.scala:62); one GPU stands in for it
        return SparkBinKmerCounter.executeJob(None, configuration)


class TestKmerCounter:
    __test__ = False

    @staticmethod
    def main(args):
        return TestKmerCounter.run(_parse(args))

    @staticmethod
    def run(configuration: TestConfiguration):
        return SparkBinKmerCounter.executeJob(None, configuration)


if __name__ == "__main__":
    LocalTestKmerCounter.main(sys.argv[1:])
