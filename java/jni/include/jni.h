/*
 * jni.h -- MINIMAL DECLARATION SET FOR SYNTAX CHECKING ONLY.
 *
 * The build image has no JDK, hence no real <jni.h>.  This file declares just the JNI types and the
 * JNIEnv members that java/jni/dge_jni.c uses, so that tests/test_jni_glue.py can run
 *     gcc -fsyntax-only -Wall -Wextra -Werror -Ijava/jni/include -Iinclude java/jni/dge_jni.c
 * on the CPU box and catch typos, wrong argument counts and missing bodies.  The member ORDER of JNINativeInterface_
 * below is NOT the real one: never compile a binary against this header.  A maintainer builds the shim with the
 * JDK's own headers (-I$JAVA_HOME/include -I$JAVA_HOME/include/linux), which take precedence over this directory.
 * Written from the public JNI specification (type names and function signatures); no JDK source was copied.
 */
#ifndef DGE_SYNTAX_ONLY_JNI_H
#define DGE_SYNTAX_ONLY_JNI_H
#include <stdint.h>
#define DGE_JNI_SYNTAX_ONLY 1
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
#define JNI_FALSE 0
#define JNI_TRUE 1
typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef uint16_t jchar;
typedef int16_t jshort;
typedef float jfloat;
typedef double jdouble;
typedef jint jsize;
struct _jobject;
typedef struct _jobject *jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jthrowable;
typedef jobject jarray;
typedef jarray jintArray;
typedef jarray jlongArray;
typedef jarray jshortArray;
typedef jarray jbyteArray;
typedef jarray jfloatArray;
typedef jarray jdoubleArray;
struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;
struct JNINativeInterface_ {
    jclass (*FindClass)(JNIEnv *, const char *);
    jint (*ThrowNew)(JNIEnv *, jclass, const char *);
    jboolean (*ExceptionCheck)(JNIEnv *);
    jsize (*GetArrayLength)(JNIEnv *, jarray);
    const char *(*GetStringUTFChars)(JNIEnv *, jstring, jboolean *);
    void (*ReleaseStringUTFChars)(JNIEnv *, jstring, const char *);
    jint *(*GetIntArrayElements)(JNIEnv *, jintArray, jboolean *);
    void (*ReleaseIntArrayElements)(JNIEnv *, jintArray, jint *, jint);
    jlong *(*GetLongArrayElements)(JNIEnv *, jlongArray, jboolean *);
    void (*ReleaseLongArrayElements)(JNIEnv *, jlongArray, jlong *, jint);
    jshort *(*GetShortArrayElements)(JNIEnv *, jshortArray, jboolean *);
    void (*ReleaseShortArrayElements)(JNIEnv *, jshortArray, jshort *, jint);
    jbyte *(*GetByteArrayElements)(JNIEnv *, jbyteArray, jboolean *);
    void (*ReleaseByteArrayElements)(JNIEnv *, jbyteArray, jbyte *, jint);
    jfloat *(*GetFloatArrayElements)(JNIEnv *, jfloatArray, jboolean *);
    void (*ReleaseFloatArrayElements)(JNIEnv *, jfloatArray, jfloat *, jint);
    jdouble *(*GetDoubleArrayElements)(JNIEnv *, jdoubleArray, jboolean *);
    void (*ReleaseDoubleArrayElements)(JNIEnv *, jdoubleArray, jdouble *, jint);
    jdoubleArray (*NewDoubleArray)(JNIEnv *, jsize);
    void (*SetDoubleArrayRegion)(JNIEnv *, jdoubleArray, jsize, jsize, const jdouble *);
    jbyteArray (*NewByteArray)(JNIEnv *, jsize);
    void (*SetByteArrayRegion)(JNIEnv *, jbyteArray, jsize, jsize, const jbyte *);
};
#endif
