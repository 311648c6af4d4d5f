/* c2a_b200 -- unit-test hooks: the device functions of the CCD hot path, one element per thread, so
 * that tests/ can check each against the CPU oracle (and sin/cos against the host libm).
 * Host pointers in, host pointers out; each call copies, launches one kernel and synchronises.
 * Not part of the drop-in boundary. */
#ifndef C2A_B200_TESTING_H
#define C2A_B200_TESTING_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* C2ARectDist (C2A/C2A_RectDist.h:157-934): R [n][9], T [n][3], ab [n][4] = a0,a1,b0,b1 ->
 * dist [n], S [n][3] (pre-filled with NaN where the reference leaves S untouched). */
int c2a_b200_test_rect_dist(const double *R, const double *T, const double *ab, int64_t n, double *dist, double *S);

/* PQP TriDistance (call sites C2A/src/C2A.cpp:1148,1916): R [n][9], T [n][3], t1/t2 [n][9] ->
 * dist [n], pq [n][6]. */
int c2a_b200_test_tri_distance(const double *R, const double *T, const double *t1, const double *t2, int64_t n,
                               double *dist, double *pq);

/* CInterpMotion_Linear::integrate and the two motion bounds: rec [n][24] (one object's motion record),
 * t [n], ang_radius [n], dir [n][3] -> out [n][14] = R(9) T(3) computeTOC_MotionBound computeTOC. */
int c2a_b200_test_motion(const double *rec, const double *t, const double *ang_radius, const double *dir, int64_t n,
                         double *out);

/* device sin/cos that mirror the host libm (c2a_libm.cuh) and their host twins (no GPU needed). */
int c2a_b200_test_sincos(const double *x, int64_t n, double *s, double *c);
int c2a_b200_host_sincos(const double *x, int64_t n, double *s, double *c);

/* Phase statistics of the solve kernel (development aid): see c2a_kernels.cu.  The counters are compiled out of the product
 * build (they cost 6 %): all zeros unless the library was built with -DC2A_SOLVE_STATS=1 (scripts/build_variant.py). */
int c2a_b200_phase_stats(int32_t enable, uint64_t *out20);

/* Durations (ms) of the three kernels of the calling thread's last batch launch (CUDA events on its stream). */
int c2a_b200_kernel_times(double *out3);

/* Counters of the wide traversal kernel (development aid): see c2a_kernels.cu. */
int c2a_b200_wide_stats(int32_t enable, uint64_t *out24);  /* out: 24 words */

/* Per-query timeline (development aid): n > 0 arms a [n][2] device buffer that the next batches of <= n queries fill
 * with the globaltimer (ns) at claim and at result write-out; n == 0 copies it to out; n < 0 frees it. */
int c2a_b200_query_trace(int64_t n, uint64_t *out);

/* FP64 pipe peak of the current device in TFLOP/s: dependent-free DFMA chains, and the same with
 * separate DMUL + DADD (the product is built with -fmad=false, so that is its ceiling). */
int c2a_b200_fp64_peak(double *tflops_fma, double *tflops_mul_add);

/* Wall-clock breakdown (seconds) of this thread's last host-buffer call: stream+arena allocation, motion
 * set-up, claim order, enqueue (H2D + launch), wait for the kernels, D2H of the results, release, total. */
int c2a_b200_host_timing(double *out8);

#ifdef __cplusplus
}
#endif
#endif
