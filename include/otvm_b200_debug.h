/* otvm_b200 — tuning / diagnostic hooks of libotvm_sm100.so.  NOT part of the stable C ABI (include/otvm_b200.h): they
 * exist for the parity tests (to force a kernel variant) and for the measurement scripts under scripts/, may change
 * between builds, and no product code calls them.  All of them set process-wide state. */
#ifndef OTVM_B200_DEBUG_H
#define OTVM_B200_DEBUG_H
#ifdef __cplusplus
extern "C" {
#endif
/* 3x3 stride-1 convolutions: one shared-memory input patch for all 9 taps.  -1 auto (default), 0 off, 1 whenever the shape
 * allows (env OTVM_CONV_HALO).  tests/test_gpu_ops.py::test_conv2d_tcgen05 runs both. */
void otvm_debug_set_conv_halo(int mode);
/* persistent patch-mode kernel: -1 auto (default), 0 off, 1 whenever the shape allows (env OTVM_CONV_PERSIST); and the
 * number of times it has been launched (tests assert that the variant under test really ran) */
void otvm_debug_set_conv_persist(int mode);
long long otvm_debug_conv_persist_launches(void);
/* launches that took the cluster split-K path (K slices of a tile as a thread-block cluster, partial tiles through
 * distributed shared memory; env OTVM_CONV_CLUSTER_K=0 disables it) */
long long otvm_debug_conv_cluster_launches(void);
/* K-chunks per barrier pair on one-wave grids, 1 or 2 (env OTVM_CONV_KSUB); ring budget of multi-wave grids in KB (0 = default) */
void otvm_debug_set_conv_ksub(int n);
void otvm_debug_set_conv_budget_kb(int kb);
/* per-CTA clock64 / globaltimer stamps of the next tcgen05 convolution / Memory.read launches into a device buffer
 * ([grid][64] long long; NULL switches them off).  scripts/conv_ts*.py, scripts/read_ts.py */
void otvm_debug_set_conv_timestamps(long long* device_buffer);
void otvm_debug_set_read_timestamps(long long* device_buffer);
#ifdef __cplusplus
}
#endif
#endif
