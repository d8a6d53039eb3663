/*
 * Minimal stand-in for the JDK's <jni.h>, for SYNTAX and SIGNATURE checks of java/jni/rapid_jni.c in an image without a JDK
 * (tests/test_java_seam.py runs `gcc -fsyntax-only` with this directory on the include path).  Only what the glue uses is
 * declared; the types and the members of the JNIEnv function table carry the names and prototypes the JNI specification
 * gives them (Java Native Interface Specification, ch. 4 "JNI Functions").  Test infrastructure: never shipped, never linked.
 */
#ifndef RAPID_B200_TEST_JNI_STUB_H
#define RAPID_B200_TEST_JNI_STUB_H

#include <stdint.h>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
#define JNI_COMMIT 1
#define JNI_OK 0
#define JNI_FALSE 0
#define JNI_TRUE 1

typedef uint8_t jboolean;
typedef int8_t jbyte;
typedef uint16_t jchar;
typedef int16_t jshort;
typedef int32_t jint;
typedef int64_t jlong;
typedef float jfloat;
typedef double jdouble;
typedef jint jsize;

struct _jobject;
typedef struct _jobject* jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jarray;
typedef jarray jbyteArray;
typedef jarray jintArray;
typedef jarray jlongArray;

struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;

struct JNINativeInterface_ {
    jstring (*NewStringUTF)(JNIEnv* env, const char* utf);
    jsize (*GetArrayLength)(JNIEnv* env, jarray array);
    jlongArray (*NewLongArray)(JNIEnv* env, jsize len);
    jbyte* (*GetByteArrayElements)(JNIEnv* env, jbyteArray array, jboolean* isCopy);
    jint* (*GetIntArrayElements)(JNIEnv* env, jintArray array, jboolean* isCopy);
    jlong* (*GetLongArrayElements)(JNIEnv* env, jlongArray array, jboolean* isCopy);
    void (*ReleaseByteArrayElements)(JNIEnv* env, jbyteArray array, jbyte* elems, jint mode);
    void (*ReleaseIntArrayElements)(JNIEnv* env, jintArray array, jint* elems, jint mode);
    void (*ReleaseLongArrayElements)(JNIEnv* env, jlongArray array, jlong* elems, jint mode);
    void (*SetLongArrayRegion)(JNIEnv* env, jlongArray array, jsize start, jsize len, const jlong* buf);
    void* (*GetDirectBufferAddress)(JNIEnv* env, jobject buf);
    jlong (*GetDirectBufferCapacity)(JNIEnv* env, jobject buf);
};

#endif
