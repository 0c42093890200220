"""TestConfiguration — mirror of skc.test.testutil.TestConfiguration (test/package.scala:16-42)."""
from dataclasses import dataclass

debugDirectory = "/tmp/"                        # test/package.scala:11


@dataclass
class TestConfiguration:
    __test__ = False                            # not a pytest class
    dataset: str
    outputDirectory: str
    k: int
    m: int
    x: int
    max_b: int = 2000
    sequenceType: int = 0
    canonical: bool = True                      # never read by the hot path (SparkBinKmerCounter.scala:34 `bothStrands`)
    debug: bool = False
    write: bool = True
    useKryoSerializer: bool = False
    useHT: bool = False
    useCustomPartitioner: bool = False
    numPartitionTasks: int = 0
    prefix: str = ""

    @property
    def b(self) -> int:                         # test/package.scala:32
        return int(min(4 ** self.m, self.max_b))

    @property
    def outputDir(self) -> str:                 # test/package.scala:33 — plain concatenation, no separator added
        stem = self.prefix + "k%d_m%d_x%d_b%d" % (self.k, self.m, self.x, self.b)
        if self.debug:
            return debugDirectory + stem
        return self.outputDirectory + stem + "_s%d" % self.sequenceType

    def __str__(self) -> str:                   # test/package.scala:37-41
        s = ("Kmer counting on Spark. \nTest parameters:\nDataset: %s\nk: %d\nm: %d\nx: %d\nb: %d\nSequence type: %d"
             "\nUsing HT:  %s\nWriting: %s\nUsing Kryo Serializer: %s\nMultiprocessor Scheduliong Partitioning: %s"
             % (self.dataset, self.k, self.m, self.x, self.b, self.sequenceType, str(self.useHT).lower(),
                str(self.write).lower(), str(self.useKryoSerializer).lower(), str(self.useCustomPartitioner).lower()))
        if self.useCustomPartitioner:
            s += "\t no. partition tasks: %d" % self.numPartitionTasks
        return s
