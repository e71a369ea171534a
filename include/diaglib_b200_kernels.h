/*
 * diaglib_b200_kernels.h — kernel-level C-ABI entry points.  These are NOT part of the
 * reference's interface; they expose the individual sm_100a kernels behind the drivers so
 * that tests/ can check each one against the CPU oracle and bench.py can time the dominant
 * kernel for its roofline leg.  All matrix arguments are column-major.  "dev" = device
 * pointer, "host" = host pointer.
 */
#ifndef DIAGLIB_B200_KERNELS_H
#define DIAGLIB_B200_KERNELS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* device memory helpers (so that harnesses need nothing but this library) */
void* diaglib_b200_malloc(int64_t bytes);
void diaglib_b200_free(void* dev);
int32_t diaglib_b200_h2d(void* dev, const void* host, int64_t bytes);
int32_t diaglib_b200_d2h(void* host, const void* dev, int64_t bytes);
int32_t diaglib_b200_d2d(void* dst_dev, const void* src_dev, int64_t bytes);
int32_t diaglib_b200_sync(void);
/* fills an n x m block (ld) with stateless-hash U[0,1) values (test / bench data) */
int32_t diaglib_b200_k_fill_uniform(double* dev, int64_t n, int32_t m, int64_t ld, int64_t seed_row0);
/* CUDA-event stopwatch on the library stream */
void diaglib_b200_timer_start(void);
double diaglib_b200_timer_stop_ms(void);
int32_t diaglib_b200_num_sms(void);

/* C(p x q, dev, ldc) = A(n x p, dev)^T B(n x q, dev), all-reduced over the communicator.
 * replaces dgemm('t','n') at diaglib.f90:313,403,1691,3256,3543 */
int32_t diaglib_b200_k_gram(int64_t n, const double* a_dev, int64_t lda, int32_t p, const double* b_dev,
                            int64_t ldb, int32_t q, double* c_dev, int32_t ldc, int32_t sym_lower);
/* Y(n x q, dev) = alpha V(n x p, dev) C(p x q, dev) + beta Y.  replaces dgemm('n','n') at
 * diaglib.f90:322,420,495,1717,3544 */
int32_t diaglib_b200_k_block_mul(int64_t n, const double* v_dev, int64_t ldv, int32_t p, const double* c_dev,
                                 int32_t ldc, int32_t q, double alpha, double beta, double* y_dev, int64_t ldy);
/* block multiply fused with the metric of its result: Y = alpha V C + beta Y, G(q x q, dev) = Y^T Y.
 * replaces dgemm/dtrmm (diaglib.f90:3544 / 3327) followed by the dgemm('t','n') of the next ortho_cd pass (3256) */
int32_t diaglib_b200_k_block_mul_gram(int64_t n, const double* v_dev, int64_t ldv, int32_t p, const double* c_dev,
                                      int32_t ldc, int32_t q, double alpha, double beta, double* y_dev, int64_t ldy,
                                      int32_t upper_tri, double* g_dev, int32_t ldg);
/* r = ax - theta_j x, norms[0..m) = sum r^2 (all-reduced), norms[m..2m) = max|r| (all-reduced);
 * theta (m, host), active (m, host), norms (2m, host).  diaglib.f90:428-442 */
int32_t diaglib_b200_k_residual(int64_t n, int32_t m, const double* ax_dev, int64_t ldax, const double* x_dev,
                                int64_t ldx, const double* theta_host, const int32_t* active_host, double* r_dev,
                                int64_t ldr, double* norms_host);
/* U (n x m, device) <- U * T in place, T (m x m, ld m, device) upper triangular with explicit zeros below
 * the diagonal: the dtrmm('r','l','t','n') of ortho_cd (diaglib.f90:3327) */
int32_t diaglib_b200_k_trmm(int64_t n, double* u, int64_t ldu, int32_t m, const double* t_dev);
/* u (n x k) <- u - x (n x m) xu (m x k, device, ld m): the projection step of ortho_vs_x (diaglib.f90:3544) the
 * way the drivers run it (one product over [x u] when u is the block right behind x) */
int32_t diaglib_b200_k_project_out(int64_t n, int32_t m, int32_t k, const double* x, int64_t ldx, const double* xu_dev, double* u,
                                   int64_t ldu);
/* the same product written to another block, Y (n x m, ldy) = U * T (measurement of in-place vs out-of-place) */
int32_t diaglib_b200_k_trmm_oop(int64_t n, const double* u, int64_t ldu, int32_t m, const double* t_dev, double* y, int64_t ldy);
/* synthetic FCI-like matrix of config C4 (same arithmetic as diaglib_b200/problems.py
 * fci_like): fills col/val/diag (device arrays) for the global rows [r0, r1) of the n-row matrix
 * from the row pointers (device, r1 - r0 + 1 entries, starting at 0).  strides: n_strides sorted
 * positive strides (host).  Columns are written in the LOCAL numbering of a row-partitioned run:
 * c in [r0, r1) -> c - r0; c in [lo_prev, r0) -> (r1 - r0) + (c - lo_prev); c in [r1, hi_next) ->
 * (r1 - r0) + (r0 - lo_prev) + (c - r1) (pass lo_prev = r0, hi_next = r1 for one rank). */
int32_t diaglib_b200_k_gen_fci(int64_t n, int64_t r0, int64_t r1, int32_t n_strides, const int64_t* strides_host,
                               double big_delta, int64_t seed, int64_t lo_prev, int64_t hi_next,
                               const int64_t* rowptr_dev, int32_t* col_dev, double* val_dev, double* diag_dev);
/* r = A x - theta x with the installed matrix on a device block (n_loc x m): per-column sum of
 * squares and max |r| over all ranks -> norms_host[0..m) and [m..2m).  An independent check of
 * a driver's returned pairs (the drivers test the recurrence residual). */
int32_t diaglib_b200_k_true_residual(int32_t n_loc, int32_t m, const double* x_dev, const double* theta_host,
                                     double* norms_host);
/* dsyev('v',uplo) replacement on a host matrix (k x k, lda): a overwritten by eigenvectors,
 * w ascending.  returns sweeps, +1000 when the one-sided solver on the Cholesky factor delivered
 * (positive definite input), < 0 if not converged.  diaglib.f90:315,406,1708 */
int32_t diaglib_b200_k_sym_eig(int32_t k, double* a_host, int32_t lda, int32_t upper, double* w_host);
/* solver selection for the reduced eigenproblems: mode 0 = one-sided block Jacobi on the Cholesky
 * factor with the two-sided solver as the fallback for matrices that are not positive definite
 * (default), 1 = two-sided only; block = columns per block of the one-sided solver (0 = automatic,
 * else 4 or 8).  Returns the previous mode. */
int32_t diaglib_b200_k_set_eig_mode(int32_t mode, int32_t block);
/* experiment switches by name (coeffs_threads, coeffs_smem, spmm_chunk, eig_block, eig_mode, chol_blocked);
 * returns the previous value, -1 for an unknown name */
int32_t diaglib_b200_k_set_tuning(const char* name, int32_t value);
/* Tile schedule of the Gram kernels (host logic, no device needed): for a block of ntp x ntq tiles of 8 x 8
 * (<= 16 x 16; sym_lower: tiles on/below the diagonal only) cover[ti + tj * ntp] receives the number of tasks
 * that compute tile (ti, tj) and load4[q] the tiles given to SM sub-partition q.  Returns the task count. */
int32_t diaglib_b200_k_gram_schedule(int32_t ntp, int32_t ntq, int32_t sym_lower, int32_t* cover, int32_t* load4);
/* device time of the small replicated kernels, milliseconds per call over `reps` back-to-back
 * launches: which = 0 chol_inv on an a x a metric; which = 1 get_coeffs(len_u = a, n_max = b, n_act = c) */
double diaglib_b200_k_time_small(int32_t which, int32_t a, int32_t b, int32_t c, int32_t reps);
/* gen_david_driver restart: 0 (default) keeps B times the restart vectors in bspace, 1 executes the
 * reference's literal `bspace = zero` (diaglib.f90:2200; also DIAGLIB_B200_REFERENCE_RESTART=1).
 * Returns the previous setting. */
int32_t diaglib_b200_k_set_reference_restart(int32_t on);
/* ortho_cd / ortho_vs_x control: 1 = speculative chains decided on the device, one host
 * synchronisation per call (default); 0 = one host decision per ortho_cd pass.  Same arithmetic
 * and same decisions either way.  Returns the previous setting. */
int32_t diaglib_b200_k_set_spec_ortho(int32_t on);
/* times `reps` back-to-back reduced eigensolves of the same host matrix on the device (CUDA events
 * on the library stream, input restored by a device copy before each solve, the copy excluded by
 * measuring it separately); returns milliseconds per solve */
double diaglib_b200_k_sym_eig_time_ms(int32_t k, const double* a_host, int32_t lda, int32_t upper, int32_t reps);
/* one factor+invert step of ortho_cd on a host metric (m x m): T = L^-T (m x m).
 * out5 = {l_norm, linv_norm, shift_used, info_first, n_shifts}; returns hard_fail.
 * diaglib.f90:3261-3316 */
int32_t diaglib_b200_k_chol_inv(int32_t m, const double* metric_host, double* t_host, double* out5);
/* get_coeffs on host matrices: a_red (len_a x len_a, eigenvectors in place) -> u_p (len_u x n_act).
 * out4 = {sweeps, cd_passes, fail, qr}.  diaglib.f90:3686-3732 */
int32_t diaglib_b200_k_get_coeffs(int32_t len_a, int32_t len_u, int32_t n_max, int32_t n_act,
                                  const double* a_red_host, double* u_p_host, int32_t* out4);

#ifdef __cplusplus
}
#endif
#endif
