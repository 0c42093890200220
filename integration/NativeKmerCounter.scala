// Source only — see integration/README.md.  Drop into src/main/scala/skc/.
package skc

import skc.test.testutil.TestConfiguration

object NativeKmerCounter {
  System.loadLibrary("fastkmer_b200_jni") // libfastkmer_b200_jni.so, linked against libfastkmer_b200.so

  @native def executeJob(dataset: String, outputDirectory: String, prefix: String,
                         k: Int, m: Int, x: Int, maxB: Int, sequenceType: Int,
                         useHT: Boolean, write: Boolean, useKryo: Boolean,
                         useCustomPartitioner: Boolean, numPartitionTasks: Int): Int

  /** The same job on the first nGpus GPUs of this node (fkm_execute_job_multi). */
  @native def executeJobMulti(nGpus: Int, dataset: String, outputDirectory: String, prefix: String,
                              k: Int, m: Int, x: Int, maxB: Int, sequenceType: Int,
                              useHT: Boolean, write: Boolean, useKryo: Boolean,
                              useCustomPartitioner: Boolean, numPartitionTasks: Int): Int

  @native def lastError(): String

  /** Body for SparkBinKmerCounter.executeJob(spark, configuration) (SparkBinKmerCounter.scala:989). */
  def run(c: TestConfiguration): Unit = {
    val rc = executeJob(c.dataset, c.outputDirectory, c.prefix, c.k, c.m, c.x, c.max_b, c.sequenceType,
      c.useHT, c.write, c.useKryoSerializer, c.useCustomPartitioner, c.numPartitionTasks)
    if (rc != 0) throw new RuntimeException("fastkmer_b200: " + lastError())
  }
}
