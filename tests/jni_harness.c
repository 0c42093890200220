/* Calls the JNI shim (integration/fkm_jni.c) the way a JVM would, through a fake JNIEnv whose strings are plain C
 * strings.  Prints "rc=<code> err=<message>".  Test infrastructure (tests/test_host_abi.py). */
#include <stdio.h>
#include <string.h>
#include <jni.h>

jint Java_skc_NativeKmerCounter_00024_executeJob(JNIEnv* env, jobject self, jstring dataset, jstring outDir, jstring prefix,
        jint k, jint m, jint x, jint maxB, jint seqType, jboolean useHT, jboolean write, jboolean useKryo,
        jboolean useCustomPartitioner, jint numPartitionTasks);
jstring Java_skc_NativeKmerCounter_00024_lastError(JNIEnv* env, jobject self);

static int g_live = 0;
static const char* get_chars(JNIEnv* env, jstring s, jboolean* is_copy) { (void)env; if (is_copy) *is_copy = 0; g_live++; return (const char*)s; }
static void release_chars(JNIEnv* env, jstring s, const char* chars) { (void)env; if ((const char*)s == chars) g_live--; }
static jstring new_string(JNIEnv* env, const char* chars) { (void)env; return (jstring)chars; }

int main(int argc, char** argv) {
    static const struct JNINativeInterface_ table = {get_chars, release_chars, new_string};
    JNIEnv env = &table;
    const char* dataset = argc > 1 ? argv[1] : "/nonexistent/input.fasta";
    jint rc = Java_skc_NativeKmerCounter_00024_executeJob(&env, NULL, (jstring)dataset, (jstring)"/tmp/fkm_jni_out/", (jstring)"p_",
                                                          28, 10, 3, 2048, 0, 1, 0, 0, 0, 0);
    const char* err = (const char*)Java_skc_NativeKmerCounter_00024_lastError(&env, NULL);
    printf("rc=%d live=%d err=%s\n", (int)rc, g_live, err ? err : "");
    return 0;
}
