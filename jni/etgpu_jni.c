/*
 * etgpu_jni.c -- JNI shim between the Scala facade (scala/lamp/extratrees/gpu/EtGpu.scala) and libetgpu.so.
 *
 * The reference targets JDK 17 (Dockerfile:1), where the Panama FFM API is still a preview, so the binding a
 * maintainer ships is JNI.  Build next to libetgpu.so:
 *     cc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../include etgpu_jni.c \
 *        -L../lamp_b200 -letgpu -o libetgpu_jni.so
 * This image has no JDK (no jni.h): the file is syntax-checked against jni/stub/jni.h (tests/test_host_logic.py),
 * a declaration-only subset of the JNI header, and cannot be run here.
 *
 * Every native method mirrors one call of include/etgpu.h; Java arrays are pinned with
 * GetPrimitiveArrayCritical for the duration of the call (the library copies to the device and never keeps a
 * host pointer).  ET_EINVAL becomes IllegalArgumentException -- the reference's require(...) failures
 * (pkg:624-633, 715-718, 779) -- every other code a RuntimeException carrying et_last_error().
 */
#include <jni.h>
#include <stddef.h>
#include <stdint.h>

#include "etgpu.h"

static void throw_et(JNIEnv *env, int rc) {
  jclass c = (*env)->FindClass(env, rc == ET_EINVAL ? "java/lang/IllegalArgumentException" : "java/lang/RuntimeException");
  if (c) (*env)->ThrowNew(env, c, et_last_error());
}

#define CTX(h) ((et_ctx *)(intptr_t)(h))
#define FOREST(h) ((et_forest *)(intptr_t)(h))

/* ---- context: et_init / et_init_multi / et_shutdown ---------------------------------------------------------- */
JNIEXPORT jlong JNICALL Java_lamp_extratrees_gpu_Native_init(JNIEnv *env, jclass cls, jintArray devices) {
  (void)cls;
  et_ctx *ctx = NULL;
  jsize n = (*env)->GetArrayLength(env, devices);
  jint *dev = (*env)->GetIntArrayElements(env, devices, NULL);
  /* one device: a plain context; several: one context over all of them (trees sharded by tree id, NCCL inside) */
  int rc = (n == 1) ? et_init(dev[0], &ctx) : et_init_multi((const int32_t *)dev, (int32_t)n, &ctx);
  (*env)->ReleaseIntArrayElements(env, devices, dev, JNI_ABORT);
  if (rc != ET_OK) {
    throw_et(env, rc);
    return 0;
  }
  return (jlong)(intptr_t)ctx;
}

JNIEXPORT void JNICALL Java_lamp_extratrees_gpu_Native_shutdown(JNIEnv *env, jclass cls, jlong ctx) {
  (void)env;
  (void)cls;
  et_shutdown(CTX(ctx));
}

/* ---- build: buildForestClassification pkg:611-681 / buildForestRegression pkg:704-764 ------------------------------- */
JNIEXPORT jlong JNICALL Java_lamp_extratrees_gpu_Native_buildClassification(
    JNIEnv *env, jclass cls, jlong ctx, jdoubleArray data, jlong n, jint d, jintArray target, jdoubleArray weights,
    jint numClasses, jint nMin, jint k, jint m, jint parallelism, jboolean bestSplit, jint maxDepth, jlong seed) {
  (void)cls;
  et_data *D = NULL;
  et_forest *F = NULL;
  jdouble *x = (*env)->GetPrimitiveArrayCritical(env, data, NULL); /* Mat.toArray: row-major, pkg:936 */
  int rc = et_data_dense_rowmajor(CTX(ctx), x, n, d, &D);
  (*env)->ReleasePrimitiveArrayCritical(env, data, x, JNI_ABORT);
  if (rc == ET_OK) {
    jint *y = (*env)->GetPrimitiveArrayCritical(env, target, NULL);
    jdouble *w = weights ? (*env)->GetPrimitiveArrayCritical(env, weights, NULL) : NULL;
    rc = et_build_classification(CTX(ctx), D, (const int32_t *)y, (*env)->GetArrayLength(env, target), w, numClasses, nMin,
                                 k, m, parallelism, bestSplit ? 1 : 0, maxDepth, seed, NULL, NULL, &F, NULL);
    if (w) (*env)->ReleasePrimitiveArrayCritical(env, weights, w, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, target, y, JNI_ABORT);
  }
  et_data_free(D);
  if (rc != ET_OK) {
    throw_et(env, rc);
    return 0;
  }
  return (jlong)(intptr_t)F;
}

JNIEXPORT jlong JNICALL Java_lamp_extratrees_gpu_Native_buildRegression(
    JNIEnv *env, jclass cls, jlong ctx, jdoubleArray data, jlong n, jint d, jdoubleArray target, jint nMin, jint k, jint m,
    jint parallelism, jboolean bestSplit, jint maxDepth, jlong seed) {
  (void)cls;
  et_data *D = NULL;
  et_forest *F = NULL;
  jdouble *x = (*env)->GetPrimitiveArrayCritical(env, data, NULL);
  int rc = et_data_dense_rowmajor(CTX(ctx), x, n, d, &D);
  (*env)->ReleasePrimitiveArrayCritical(env, data, x, JNI_ABORT);
  if (rc == ET_OK) {
    jdouble *y = (*env)->GetPrimitiveArrayCritical(env, target, NULL);
    rc = et_build_regression(CTX(ctx), D, y, (*env)->GetArrayLength(env, target), nMin, k, m, parallelism,
                             bestSplit ? 1 : 0, maxDepth, seed, NULL, NULL, &F, NULL);
    (*env)->ReleasePrimitiveArrayCritical(env, target, y, JNI_ABORT);
  }
  et_data_free(D);
  if (rc != ET_OK) {
    throw_et(env, rc);
    return 0;
  }
  return (jlong)(intptr_t)F;
}

/* ---- forest: Seq[ClassificationTree] / Seq[RegressionTree] (extratrees.scala:3-63) as pre-order arrays --------------- */
/* out[0] = trees, out[1] = leaf width, out[2] = regression flag, out[3] = nodes of the whole forest */
JNIEXPORT void JNICALL Java_lamp_extratrees_gpu_Native_forestDims(JNIEnv *env, jclass cls, jlong forest, jlongArray out) {
  (void)cls;
  int32_t m = 0, lw = 0, reg = 0;
  int64_t total = 0;
  int rc = et_forest_dims(FOREST(forest), &m, &lw, &reg, &total);
  if (rc != ET_OK) {
    throw_et(env, rc);
    return;
  }
  jlong v[4] = {m, lw, reg, total};
  (*env)->SetLongArrayRegion(env, out, 0, 4, v);
}

JNIEXPORT void JNICALL Java_lamp_extratrees_gpu_Native_exportAll(JNIEnv *env, jclass cls, jlong forest, jintArray treeSizes,
                                                                 jintArray feature, jdoubleArray cut, jbyteArray mil,
                                                                 jintArray left, jintArray right, jdoubleArray leaf) {
  (void)cls;
  jint *ts = (*env)->GetPrimitiveArrayCritical(env, treeSizes, NULL);
  jint *fe = (*env)->GetPrimitiveArrayCritical(env, feature, NULL);
  jdouble *cu = (*env)->GetPrimitiveArrayCritical(env, cut, NULL);
  jbyte *mi = (*env)->GetPrimitiveArrayCritical(env, mil, NULL);
  jint *le = (*env)->GetPrimitiveArrayCritical(env, left, NULL);
  jint *ri = (*env)->GetPrimitiveArrayCritical(env, right, NULL);
  jdouble *lf = (*env)->GetPrimitiveArrayCritical(env, leaf, NULL);
  int rc = et_forest_export_all(FOREST(forest), (int32_t *)ts, (int32_t *)fe, cu, (uint8_t *)mi, (int32_t *)le,
                                (int32_t *)ri, lf);
  (*env)->ReleasePrimitiveArrayCritical(env, leaf, lf, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, right, ri, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, left, le, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, mil, mi, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, cut, cu, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, feature, fe, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, treeSizes, ts, 0);
  if (rc != ET_OK) throw_et(env, rc);
}

/* JVM-held trees (flattened by the facade) -> a forest handle for predict */
JNIEXPORT jlong JNICALL Java_lamp_extratrees_gpu_Native_importForest(JNIEnv *env, jclass cls, jlong ctx, jint leafWidth,
                                                                     jboolean regression, jintArray treeSizes,
                                                                     jintArray feature, jdoubleArray cut, jbyteArray mil,
                                                                     jintArray left, jintArray right, jdoubleArray leaf) {
  (void)cls;
  et_forest *F = NULL;
  jsize m = (*env)->GetArrayLength(env, treeSizes);
  jint *ts = (*env)->GetPrimitiveArrayCritical(env, treeSizes, NULL);
  jint *fe = (*env)->GetPrimitiveArrayCritical(env, feature, NULL);
  jdouble *cu = (*env)->GetPrimitiveArrayCritical(env, cut, NULL);
  jbyte *mi = (*env)->GetPrimitiveArrayCritical(env, mil, NULL);
  jint *le = (*env)->GetPrimitiveArrayCritical(env, left, NULL);
  jint *ri = (*env)->GetPrimitiveArrayCritical(env, right, NULL);
  jdouble *lf = (*env)->GetPrimitiveArrayCritical(env, leaf, NULL);
  int rc = et_forest_import(CTX(ctx), (int32_t)m, leafWidth, regression ? 1 : 0, (const int32_t *)ts, (const int32_t *)fe,
                            cu, (const uint8_t *)mi, (const int32_t *)le, (const int32_t *)ri, lf, &F);
  (*env)->ReleasePrimitiveArrayCritical(env, leaf, lf, JNI_ABORT);
  (*env)->ReleasePrimitiveArrayCritical(env, right, ri, JNI_ABORT);
  (*env)->ReleasePrimitiveArrayCritical(env, left, le, JNI_ABORT);
  (*env)->ReleasePrimitiveArrayCritical(env, mil, mi, JNI_ABORT);
  (*env)->ReleasePrimitiveArrayCritical(env, cut, cu, JNI_ABORT);
  (*env)->ReleasePrimitiveArrayCritical(env, feature, fe, JNI_ABORT);
  (*env)->ReleasePrimitiveArrayCritical(env, treeSizes, ts, JNI_ABORT);
  if (rc != ET_OK) {
    throw_et(env, rc);
    return 0;
  }
  return (jlong)(intptr_t)F;
}

JNIEXPORT void JNICALL Java_lamp_extratrees_gpu_Native_freeForest(JNIEnv *env, jclass cls, jlong forest) {
  (void)env;
  (void)cls;
  et_forest_free(FOREST(forest));
}

/* ---- predict: predictClassification pkg:542-551 (out: n x numClasses row-major) / predictRegression pkg:577-586 --- */
JNIEXPORT void JNICALL Java_lamp_extratrees_gpu_Native_predict(JNIEnv *env, jclass cls, jlong ctx, jlong forest,
                                                               jboolean regression, jdoubleArray samples, jlong n, jint d,
                                                               jdoubleArray out) {
  (void)cls;
  jdouble *x = (*env)->GetPrimitiveArrayCritical(env, samples, NULL);
  jdouble *o = (*env)->GetPrimitiveArrayCritical(env, out, NULL);
  int rc = regression ? et_predict_regression(CTX(ctx), FOREST(forest), x, n, d, o, 0)
                      : et_predict_classification(CTX(ctx), FOREST(forest), x, n, d, o, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, out, o, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, samples, x, JNI_ABORT);
  if (rc != ET_OK) throw_et(env, rc);
}
