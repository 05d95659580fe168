/* jni/stub/jni.h — COMPILE-CHECK STUB, not the JDK's header.
 *
 * This image has no JDK.  The JNI shim (jni/mmplace_jni.c) is compiled in CI against this hand-written subset of the JNI
 * declarations so that its calls into include/mmplace.h stay type-correct and complete; the object it produces is never
 * loaded.  On a box with a JDK build with -I$JAVA_HOME/include -I$JAVA_HOME/include/linux instead (jni/Makefile does that
 * when JAVA_HOME is set).  Only the types and JNIEnv members the shim uses are declared; the order of the function table
 * is NOT the real one. */
#ifndef MMPLACE_STUB_JNI_H
#define MMPLACE_STUB_JNI_H
#include <stdint.h>
#define MMPLACE_STUB_JNI 1
typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef double jdouble;
typedef jint jsize;
typedef void *jobject;
typedef jobject jclass, jstring, jarray, jobjectArray, jintArray, jbyteArray, jlongArray, jdoubleArray, jthrowable;
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;
struct JNINativeInterface_ {
  jclass (*FindClass)(JNIEnv *, const char *);
  jint (*ThrowNew)(JNIEnv *, jclass, const char *);
  jobject (*NewDirectByteBuffer)(JNIEnv *, void *, jlong);
  void *(*GetDirectBufferAddress)(JNIEnv *, jobject);
  jlong (*GetDirectBufferCapacity)(JNIEnv *, jobject);
  const char *(*GetStringUTFChars)(JNIEnv *, jstring, jboolean *);
  void (*ReleaseStringUTFChars)(JNIEnv *, jstring, const char *);
  jstring (*NewStringUTF)(JNIEnv *, const char *);
  jsize (*GetArrayLength)(JNIEnv *, jarray);
  jobject (*GetObjectArrayElement)(JNIEnv *, jobjectArray, jsize);
  void (*DeleteLocalRef)(JNIEnv *, jobject);
  void *(*GetPrimitiveArrayCritical)(JNIEnv *, jarray, jboolean *);
  void (*ReleasePrimitiveArrayCritical)(JNIEnv *, jarray, void *, jint);
  void (*SetByteArrayRegion)(JNIEnv *, jbyteArray, jsize, jsize, const jbyte *);
  void (*GetByteArrayRegion)(JNIEnv *, jbyteArray, jsize, jsize, jbyte *);
  void (*SetIntArrayRegion)(JNIEnv *, jintArray, jsize, jsize, const jint *);
  void (*SetLongArrayRegion)(JNIEnv *, jlongArray, jsize, jsize, const jlong *);
  void (*SetDoubleArrayRegion)(JNIEnv *, jdoubleArray, jsize, jsize, const jdouble *);
};
#endif
