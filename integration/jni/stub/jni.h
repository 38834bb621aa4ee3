/*
 * stub jni.h — NOT the JDK header.  The build image has no JDK, so integration/jni/ps_jni.c cannot be compiled against the real
 * <jni.h>; this file declares just the types and the JNINativeInterface_ entries the shim uses, with the signatures the JNI
 * specification (Java SE, "JNI Functions") gives them, so that `gcc -fsyntax-only -Iintegration/jni/stub -Iinclude
 * integration/jni/ps_jni.c` type-checks every call the shim makes into include/ps_b200.h (tests/test_capi.py::test_jni_shim_typechecks).
 * With a real JDK use -I$JAVA_HOME/include instead.
 */
#ifndef PS_STUB_JNI_H_
#define PS_STUB_JNI_H_
#include <stdint.h>

typedef int32_t jint;
typedef int64_t jlong;
typedef float jfloat;
typedef unsigned char jboolean;
typedef jint jsize;
struct _jobject;
typedef struct _jobject* jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jarray;
typedef jarray jfloatArray;
typedef jarray jintArray;
typedef jarray jbyteArray;
typedef jarray jobjectArray;
typedef signed char jbyte;
typedef jobject jthrowable;
#define JNI_FALSE 0
#define JNI_TRUE 1
#define JNI_ABORT 2
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL

struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;

struct JNINativeInterface_ {
  jclass (*FindClass)(JNIEnv* env, const char* name);
  jint (*ThrowNew)(JNIEnv* env, jclass clazz, const char* msg);
  jstring (*NewStringUTF)(JNIEnv* env, const char* utf);
  const char* (*GetStringUTFChars)(JNIEnv* env, jstring str, jboolean* isCopy);
  void (*ReleaseStringUTFChars)(JNIEnv* env, jstring str, const char* chars);
  jsize (*GetArrayLength)(JNIEnv* env, jarray array);
  jfloatArray (*NewFloatArray)(JNIEnv* env, jsize len);
  jfloat* (*GetFloatArrayElements)(JNIEnv* env, jfloatArray array, jboolean* isCopy);
  void (*ReleaseFloatArrayElements)(JNIEnv* env, jfloatArray array, jfloat* elems, jint mode);
  jint* (*GetIntArrayElements)(JNIEnv* env, jintArray array, jboolean* isCopy);
  void (*ReleaseIntArrayElements)(JNIEnv* env, jintArray array, jint* elems, jint mode);
  jobjectArray (*NewObjectArray)(JNIEnv* env, jsize length, jclass elementClass, jobject initialElement);
  jobject (*GetObjectArrayElement)(JNIEnv* env, jobjectArray array, jsize index);
  void (*SetObjectArrayElement)(JNIEnv* env, jobjectArray array, jsize index, jobject value);
  void (*SetFloatArrayRegion)(JNIEnv* env, jfloatArray array, jsize start, jsize len, const jfloat* buf);
  void (*DeleteLocalRef)(JNIEnv* env, jobject localRef);
  void* (*GetPrimitiveArrayCritical)(JNIEnv* env, jarray array, jboolean* isCopy);
  void (*ReleasePrimitiveArrayCritical)(JNIEnv* env, jarray array, void* carray, jint mode);
};
#endif
