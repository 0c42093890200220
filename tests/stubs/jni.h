/* Minimal stand-in for the JDK's <jni.h>: only what integration/fkm_jni.c uses, with the JNI calling convention of
 * C code ((*env)->Fn(env, ...)).  Test infrastructure — the image has no JDK; a real build uses $JAVA_HOME/include. */
#ifndef FKM_TEST_JNI_H
#define FKM_TEST_JNI_H
#include <stdint.h>
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
typedef int32_t jint;
typedef uint8_t jboolean;
typedef void* jobject;
typedef jobject jstring;
struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
    const char* (*GetStringUTFChars)(JNIEnv* env, jstring s, jboolean* is_copy);
    void (*ReleaseStringUTFChars)(JNIEnv* env, jstring s, const char* chars);
    jstring (*NewStringUTF)(JNIEnv* env, const char* chars);
};
#endif
