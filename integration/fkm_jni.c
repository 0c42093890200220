/* Source only — see integration/README.md.
 * gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../include fkm_jni.c \
 *     -L../fastkmer_b200 -lfastkmer_b200 -o libfastkmer_b200_jni.so                                  */
#include <jni.h>
#include <string.h>
#include "fastkmer_b200.h"

static fkm_ctx* g_ctx; /* one context per process, created on first use */

JNIEXPORT jint JNICALL Java_skc_NativeKmerCounter_00024_executeJob(JNIEnv* env, jobject self,
        jstring dataset, jstring outDir, jstring prefix, jint k, jint m, jint x, jint maxB, jint seqType,
        jboolean useHT, jboolean write, jboolean useKryo, jboolean useCustomPartitioner, jint numPartitionTasks) {
    (void)self;
    if (!g_ctx && fkm_ctx_create(-1, NULL, &g_ctx) != FKM_OK) return FKM_ECUDA;
    fkm_config c;
    memset(&c, 0, sizeof c);
    c.k = k; c.m = m; c.x = x; c.max_b = maxB; c.sequence_type = seqType;
    c.use_ht = useHT; c.write = write; c.use_kryo_serializer = useKryo;
    c.use_custom_partitioner = useCustomPartitioner; c.num_partition_tasks = numPartitionTasks;
    c.dataset = (*env)->GetStringUTFChars(env, dataset, NULL);
    c.output_directory = (*env)->GetStringUTFChars(env, outDir, NULL);
    c.prefix = (*env)->GetStringUTFChars(env, prefix, NULL);
    fkm_stats st;
    int rc = fkm_execute_job(g_ctx, &c, &st); /* SparkBinKmerCounter.scala:989-1046 */
    (*env)->ReleaseStringUTFChars(env, dataset, c.dataset);
    (*env)->ReleaseStringUTFChars(env, outDir, c.output_directory);
    (*env)->ReleaseStringUTFChars(env, prefix, c.prefix);
    return rc;
}

/* The same job on the first nGpus GPUs of the node (fkm_execute_job_multi): what a driver with several GPUs calls instead. */
JNIEXPORT jint JNICALL Java_skc_NativeKmerCounter_00024_executeJobMulti(JNIEnv* env, jobject self, jint nGpus,
        jstring dataset, jstring outDir, jstring prefix, jint k, jint m, jint x, jint maxB, jint seqType,
        jboolean useHT, jboolean write, jboolean useKryo, jboolean useCustomPartitioner, jint numPartitionTasks) {
    (void)self;
    int32_t dev[64];
    int n = nGpus < 1 ? 1 : nGpus > 64 ? 64 : nGpus;
    for (int i = 0; i < n; i++) dev[i] = i;
    fkm_config c;
    memset(&c, 0, sizeof c);
    c.k = k; c.m = m; c.x = x; c.max_b = maxB; c.sequence_type = seqType;
    c.use_ht = useHT; c.write = write; c.use_kryo_serializer = useKryo;
    c.use_custom_partitioner = useCustomPartitioner; c.num_partition_tasks = numPartitionTasks;
    c.dataset = (*env)->GetStringUTFChars(env, dataset, NULL);
    c.output_directory = (*env)->GetStringUTFChars(env, outDir, NULL);
    c.prefix = (*env)->GetStringUTFChars(env, prefix, NULL);
    fkm_stats st;
    int rc = fkm_execute_job_multi(dev, n, &c, &st);
    (*env)->ReleaseStringUTFChars(env, dataset, c.dataset);
    (*env)->ReleaseStringUTFChars(env, outDir, c.output_directory);
    (*env)->ReleaseStringUTFChars(env, prefix, c.prefix);
    return rc;
}

JNIEXPORT jstring JNICALL Java_skc_NativeKmerCounter_00024_lastError(JNIEnv* env, jobject self) {
    (void)self;
    return (*env)->NewStringUTF(env, fkm_last_error());
}
