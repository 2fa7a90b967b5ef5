/*
 * snappy_gpu_jni.c -- JNI shim between the reference's JVM operators and libsnappygpu.so.
 *
 * NOT COMPILED IN THIS REPOSITORY'S CONTAINER (no JDK / jni.h here); kept to the thin pattern of the
 * reference's only native precedent, org.apache.spark.unsafe.Native
 * (/root/reference/aqp/src/main/cpp/io/snappydata/DataOptimizations.c:26-68): static natives taking raw
 * addresses and sizes as jlong/jint, returning primitives, no JNI object access beyond array pinning.
 *
 * Build on a box with a JDK:
 *   gcc -O2 -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../include \
 *       snappy_gpu_jni.c -L../../snappydata_b200/csrc -lsnappygpu -o libsnappygpujni.so
 *
 * Scala side: jvm/src/main/scala/io/snappydata/gpu/SnappyGpuNative.scala.
 * Errors: non-zero sd_status -> RuntimeException(sd_last_error()); there is no CPU fallback.
 */
#include <jni.h>
#include <stdint.h>
#include <string.h>

#include "snappy_gpu.h"

#define JFN(name) Java_io_snappydata_gpu_SnappyGpuNative_00024_##name

static void throw_last(JNIEnv* env) {
  jclass cls = (*env)->FindClass(env, "java/lang/RuntimeException");
  if (cls) (*env)->ThrowNew(env, cls, sd_last_error());
}

JNIEXPORT jint JNICALL JFN(init)(JNIEnv* env, jobject self, jint device) {
  int rc = sd_init(device);
  if (rc) throw_last(env);
  return rc;
}

/* planDesc: address of a serialized sd_plan_desc built off-heap by GpuPlanSerializer (all pointers inside
 * are absolute addresses into the same off-heap block) */
JNIEXPORT jlong JNICALL JFN(planCreate)(JNIEnv* env, jobject self, jlong planDescAddr) {
  sd_plan* p = NULL;
  if (sd_plan_create((const sd_plan_desc*)(intptr_t)planDescAddr, &p)) { throw_last(env); return 0; }
  return (jlong)(intptr_t)p;
}

JNIEXPORT void JNICALL JFN(planSetLiterals)(JNIEnv* env, jobject self, jlong plan, jlong literalsAddr, jint n) {
  if (sd_plan_set_literals((sd_plan*)(intptr_t)plan, (const sd_literal*)(intptr_t)literalsAddr, n)) throw_last(env);
}

/* One column batch.  addrs/lens: per projected column the address and length of the value buffer; for heap
 * ByteBuffers the Scala side passes the backing byte[] instead (ColumnTableScan.scala:430-437 handles both),
 * pinned here with GetPrimitiveArrayCritical for the duration of the call -- sd_batch_submit has copied the
 * bytes to the device when it returns (ownership rule, ColumnBatchIterator.scala:165-184). */
JNIEXPORT void JNICALL JFN(batchSubmit)(JNIEnv* env, jobject self, jlong plan, jint numRows, jint nCols,
                                        jlongArray colAddrs, jlongArray colLens, jobjectArray heapCols,
                                        jlongArray delta0Addrs, jlongArray delta0Lens, jlongArray delta1Addrs,
                                        jlongArray delta1Lens, jlong deleteAddr, jlong deleteLen, jlong statsAddr,
                                        jlong statsLen, jint statsNCols, jint bucketId, jlong batchId) {
  enum { MAXC = 256 };
  const void* cols[MAXC]; const void* d0[MAXC]; const void* d1[MAXC];
  int64_t lens[MAXC], d0l[MAXC], d1l[MAXC];
  jbyteArray pinned[MAXC]; void* pinnedPtr[MAXC];
  if (nCols > MAXC) { (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/IllegalArgumentException"), "too many columns"); return; }
  jlong* a = (*env)->GetLongArrayElements(env, colAddrs, NULL);
  jlong* l = (*env)->GetLongArrayElements(env, colLens, NULL);
  jlong* a0 = delta0Addrs ? (*env)->GetLongArrayElements(env, delta0Addrs, NULL) : NULL;
  jlong* l0 = delta0Lens ? (*env)->GetLongArrayElements(env, delta0Lens, NULL) : NULL;
  jlong* a1 = delta1Addrs ? (*env)->GetLongArrayElements(env, delta1Addrs, NULL) : NULL;
  jlong* l1 = delta1Lens ? (*env)->GetLongArrayElements(env, delta1Lens, NULL) : NULL;
  for (int i = 0; i < nCols; i++) {
    pinned[i] = NULL; pinnedPtr[i] = NULL;
    lens[i] = l[i];
    d0[i] = a0 ? (const void*)(intptr_t)a0[i] : NULL; d0l[i] = l0 ? l0[i] : 0;
    d1[i] = a1 ? (const void*)(intptr_t)a1[i] : NULL; d1l[i] = l1 ? l1[i] : 0;
    if (a[i] != 0) cols[i] = (const void*)(intptr_t)a[i];          /* direct buffer: GetDirectBufferAddress done in Scala */
    else {                                                          /* heap buffer: pin the byte[] */
      pinned[i] = (jbyteArray)(*env)->GetObjectArrayElement(env, heapCols, i);
      pinnedPtr[i] = (*env)->GetPrimitiveArrayCritical(env, pinned[i], NULL);
      cols[i] = pinnedPtr[i];
    }
  }
  sd_batch b;
  memset(&b, 0, sizeof(b));
  b.num_rows = numRows; b.ncols = nCols; b.col_bufs = cols; b.col_lens = lens;
  b.delta0 = a0 ? d0 : NULL; b.delta0_lens = d0l; b.delta1 = a1 ? d1 : NULL; b.delta1_lens = d1l;
  b.delete_buf = (const void*)(intptr_t)deleteAddr; b.delete_len = deleteLen;
  b.stats_row = (const void*)(intptr_t)statsAddr; b.stats_len = statsLen; b.stats_ncols = statsNCols;
  b.bucket_id = bucketId; b.batch_id = batchId;
  int rc = sd_batch_submit((sd_plan*)(intptr_t)plan, &b);
  for (int i = 0; i < nCols; i++) if (pinned[i]) (*env)->ReleasePrimitiveArrayCritical(env, pinned[i], pinnedPtr[i], JNI_ABORT);
  (*env)->ReleaseLongArrayElements(env, colAddrs, a, JNI_ABORT);
  (*env)->ReleaseLongArrayElements(env, colLens, l, JNI_ABORT);
  if (a0) (*env)->ReleaseLongArrayElements(env, delta0Addrs, a0, JNI_ABORT);
  if (l0) (*env)->ReleaseLongArrayElements(env, delta0Lens, l0, JNI_ABORT);
  if (a1) (*env)->ReleaseLongArrayElements(env, delta1Addrs, a1, JNI_ABORT);
  if (l1) (*env)->ReleaseLongArrayElements(env, delta1Lens, l1, JNI_ABORT);
  if (rc) throw_last(env);
}

JNIEXPORT void JNICALL JFN(rowsSubmit)(JNIEnv* env, jobject self, jlong plan, jlong rowsAddr, jlong len, jint nrows) {
  if (sd_rows_submit((sd_plan*)(intptr_t)plan, (const void*)(intptr_t)rowsAddr, len, nrows)) throw_last(env);
}

/* returns the number of bytes written to outAddr; a negative value -needed when the buffer is too small */
JNIEXPORT jlong JNICALL JFN(planFinish)(JNIEnv* env, jobject self, jlong plan, jlong outAddr, jlong cap) {
  int64_t len = 0, nrows = 0;
  int rc = sd_plan_finish((sd_plan*)(intptr_t)plan, (void*)(intptr_t)outAddr, cap, &len, &nrows);
  if (rc == SD_ERR_OVERFLOW) return -len;
  if (rc) { throw_last(env); return 0; }
  return len;
}

JNIEXPORT void JNICALL JFN(planReset)(JNIEnv* env, jobject self, jlong plan) {
  if (sd_plan_reset((sd_plan*)(intptr_t)plan)) throw_last(env);
}

JNIEXPORT void JNICALL JFN(planMetrics)(JNIEnv* env, jobject self, jlong plan, jlongArray out) {
  int64_t m[SD_NUM_METRICS];
  if (sd_plan_metrics((sd_plan*)(intptr_t)plan, m)) { throw_last(env); return; }
  (*env)->SetLongArrayRegion(env, out, 0, SD_NUM_METRICS, (const jlong*)m);
}

JNIEXPORT void JNICALL JFN(planDestroy)(JNIEnv* env, jobject self, jlong plan) { sd_plan_destroy((sd_plan*)(intptr_t)plan); }

JNIEXPORT jlong JNICALL JFN(finalMerge)(JNIEnv* env, jobject self, jlong planDescAddr, jlong rowsAddr, jlong len, jlong outAddr, jlong cap) {
  int64_t olen = 0, nrows = 0;
  int rc = sd_final_merge((const sd_plan_desc*)(intptr_t)planDescAddr, (const void*)(intptr_t)rowsAddr, len,
                          (void*)(intptr_t)outAddr, cap, &olen, &nrows);
  if (rc == SD_ERR_OVERFLOW) return -olen;
  if (rc) { throw_last(env); return 0; }
  return olen;
}
