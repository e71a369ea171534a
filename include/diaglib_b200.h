/*
 * diaglib_b200.h — C-ABI of the B200-native replacement of diaglib's iterative-eigensolver
 * hot path (lobpcg_driver / davidson_driver iteration body and the block kernels under it).
 *
 * Everything is passed BY REFERENCE in the argument order of the Fortran reference, so the
 * entry points can be bound from Fortran with a plain `bind(C)` interface block (see
 * diaglib_b200/fortran/diaglib_b200_shim.f90 and INTEGRATION.md).  Citations are file:line
 * into the reference tree (Molecolab-Pisa/diaglib).
 *
 * Contract differences from the CPU reference (all documented in INTEGRATION.md):
 *  - `n` is the number of LOCAL rows owned by the calling rank (== global n on one GPU);
 *    every n-long block is row-partitioned over the ranks of the communicator installed with
 *    diaglib_b200_comm_init.
 *  - the callbacks are invoked with DEVICE pointers (x, ax live in HBM) and must enqueue
 *    their work on diaglib_b200_stream().  Conforming built-in callbacks are provided
 *    (diaglib_b200_csr_matvec / diaglib_b200_diag_precnd) that act on the matrix installed
 *    with diaglib_b200_set_csr — the same "matrix in a global" pattern the reference's own
 *    callbacks use (utils.f90:4, main.f90:73,87).
 *  - `evec` / `eig` may be host or device pointers (detected); host buffers are staged
 *    through HBM once at entry and once at exit.
 *  - the reference's hard `stop`s (dsyev failure 412-415, ortho_vs_x breakdown 3568,
 *    Cholesky shift loop exhausted 3276-3284, allocation failure 3798-3801) become
 *    ok=.false. plus a non-zero diaglib_b200_last_status().
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with
 * status DIAGLIB_B200_ENODEVICE.
 */
#ifndef DIAGLIB_B200_H
#define DIAGLIB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* user callbacks — spec diaglib.f90:62-72, README.md:29-40; declared `external` 228, 1539 */
typedef void (*diaglib_matvec_t)(const int32_t* n, const int32_t* m, const double* x, double* ax);
typedef void (*diaglib_precnd_t)(const int32_t* n, const int32_t* m, const double* shift, const double* x,
                                 double* px);

enum {
  DIAGLIB_B200_OK = 0,
  DIAGLIB_B200_EDSYEV = 1,      /* reduced eigensolver did not converge   (diaglib.f90:412-415) */
  DIAGLIB_B200_EALLOC = 2,      /* device allocation failed               (diaglib.f90:3798-3801) */
  DIAGLIB_B200_ECHOL = 3,       /* Cholesky level-shift loop exhausted    (diaglib.f90:3276-3284) */
  DIAGLIB_B200_EORTHO = 4,      /* ortho_vs_x did not converge            (diaglib.f90:3568) */
  DIAGLIB_B200_ENODEVICE = 5,   /* no CUDA device / library not initialised */
  DIAGLIB_B200_EARG = 6,        /* invalid argument (n_targ > n_max, gen_eig without bvec, no matrix installed) */
  DIAGLIB_B200_ECOMM = 7        /* NCCL failure */
};

/* ---- drivers ----------------------------------------------------------------------- */

/* replaces lobpcg_driver, diaglib.f90:171-172 (argument list 221-228).  logicals are
 * 4-byte integers (gfortran logical(4)).  With gen_eig = .true. the generalized problem
 * A x = lambda B x is solved (diaglib.f90:299-302, 329-346, 357-364, 422-436, 500-526 with
 * b_ortho 3094-3183 and b_ortho_vs_x 3576-3663): bvec(n,m,x,bx) applies the metric, same
 * contract as matvec; with gen_eig = .false. bvec is never called and may be NULL. */
void diaglib_b200_lobpcg_driver(const int32_t* verbose, const int32_t* gen_eig, const int32_t* n,
                                const int32_t* n_targ, const int32_t* n_max, const int32_t* max_iter,
                                const double* tol, const double* shift, diaglib_matvec_t matvec,
                                diaglib_precnd_t precnd, diaglib_matvec_t bvec, double* eig, double* evec,
                                int32_t* ok);

/* replaces davidson_driver, diaglib.f90:1483-1484 (argument list 1532-1539) */
void diaglib_b200_davidson_driver(const int32_t* verbose, const int32_t* n, const int32_t* n_targ,
                                  const int32_t* n_max, const int32_t* max_iter, const double* tol,
                                  const int32_t* max_dav, const double* shift, diaglib_matvec_t matvec,
                                  diaglib_precnd_t precnd, double* eig, double* evec, int32_t* ok);

/* replaces gen_david_driver, diaglib.f90:1855-1856 (argument list 1907-1913): Davidson-Liu for
 * A x = lambda B x.  One deliberate deviation: after a restart the reference clears all of
 * bspace (2200), losing B times the restart vectors, and from then on converges to wrong
 * eigenvalues; this library keeps those columns. */
void diaglib_b200_gen_david_driver(const int32_t* verbose, const int32_t* n, const int32_t* n_targ,
                                   const int32_t* n_max, const int32_t* max_iter, const double* tol,
                                   const int32_t* max_dav, const double* shift, diaglib_matvec_t matvec,
                                   diaglib_precnd_t precnd, diaglib_matvec_t bvec, double* eig, double* evec,
                                   int32_t* ok);

/* lrprec(n, m, fac, xp, xm, yp, ym): preconditioner of the linear-response solver (contract of
 * lrprec_1/lrprec_2, main.f90:234-281); device pointers, work enqueued on diaglib_b200_stream() */
typedef void (*diaglib_lrprec_t)(const int32_t* n, const int32_t* m, const double* fac, const double* xp,
                                 const double* xm, double* yp, double* ym);
/* replaces caslr_eff_driver, diaglib.f90:1024-1025 (argument list 1101-1110): the linear-response
 * problem [A B; B A][Y;Z] = w [S D; -D -S][Y;Z] through the four products (A+B)x, (A-B)x, (S+D)x,
 * (S-D)x (same contract as matvec).  evec is (n2 = 2n, n_max): rows [0,n) = Y, [n,2n) = Z. */
void diaglib_b200_caslr_eff_driver(const int32_t* verbose, const int32_t* n, const int32_t* n2, const int32_t* n_targ,
                                   const int32_t* n_max, const int32_t* max_iter, const double* tol,
                                   const int32_t* max_dav, diaglib_matvec_t apbmul, diaglib_matvec_t ambmul,
                                   diaglib_matvec_t spdmul, diaglib_matvec_t smdmul, diaglib_lrprec_t lrprec,
                                   double* eig, double* evec, int32_t* ok);

/* ---- public block kernels of the reference (public list diaglib.f90:166-167) --------- */

/* replaces ortho_cd, diaglib.f90:3185 : u(n,m) in/out (host or device) */
void diaglib_b200_ortho_cd(const int32_t* n, const int32_t* m, double* u, double* growth, int32_t* ok);
/* replaces ortho_vs_x, diaglib.f90:3481 : x(n,m) in, u(n,k) in/out; ax/au are accepted for
 * signature compatibility and, as in the reference, never referenced */
void diaglib_b200_ortho_vs_x(const int32_t* n, const int32_t* m, const int32_t* k, const double* x, double* u,
                             const double* ax, double* au);
/* replaces b_ortho, diaglib.f90:3094 : u(n,m), bu(n,m) = B u, both in/out */
void diaglib_b200_b_ortho(const int32_t* n, const int32_t* m, double* u, double* bu);
/* replaces b_ortho_vs_x, diaglib.f90:3576 : x(n,m), bx(n,m) = B x in, u(n,k) in/out */
void diaglib_b200_b_ortho_vs_x(const int32_t* n, const int32_t* m, const int32_t* k, const double* x,
                               const double* bx, double* u);
/* replaces ortho (QR fallback), diaglib.f90:3052 : second argument untouched, as in the reference */
void diaglib_b200_ortho(const int32_t* n, const int32_t* m, double* u, double* w);

/* ---- built-in conforming callbacks ------------------------------------------------- */

/* CSR block matvec on the installed matrix; replaces the role of mmult, main.f90:72-90 */
void diaglib_b200_csr_matvec(const int32_t* n, const int32_t* m, const double* x, double* ax);
/* CSR block product with the metric installed by diaglib_b200_set_csr_b; a conforming bvec
 * (the reference's tests use bmult, main.f90, in the same role) */
void diaglib_b200_csr_bvec(const int32_t* n, const int32_t* m, const double* x, double* bx);
/* the four products and the preconditioner of the linear-response problem on the matrices installed
 * by diaglib_b200_set_csr_lr / set_lr_diag; roles of apbvec, ambvec, spdvec, smdvec and lrprec_2,
 * main.f90:173-232, 257-281 */
void diaglib_b200_csr_apbmul(const int32_t* n, const int32_t* m, const double* x, double* y);
void diaglib_b200_csr_ambmul(const int32_t* n, const int32_t* m, const double* x, double* y);
void diaglib_b200_csr_spdmul(const int32_t* n, const int32_t* m, const double* x, double* y);
void diaglib_b200_csr_smdmul(const int32_t* n, const int32_t* m, const double* x, double* y);
void diaglib_b200_lrprec(const int32_t* n, const int32_t* m, const double* fac, const double* xp, const double* xm,
                         double* yp, double* ym);
/* diagonal shift-and-invert preconditioner; replaces the role of mprec, main.f90:146-171 */
void diaglib_b200_diag_precnd(const int32_t* n, const int32_t* m, const double* shift, const double* x,
                              double* px);

/* ---- context ------------------------------------------------------------------------ */

/* bind the library to a CUDA device and create its stream.  Returns a DIAGLIB_B200_* code. */
int32_t diaglib_b200_init(int32_t device);
void diaglib_b200_finalize(void);
/* the drivers keep their n-long workspaces cached between calls (the reference allocates and
 * frees them per call, diaglib.f90:251-276,550-552); this returns them to the device */
void diaglib_b200_release_workspace(void);
/* cudaStream_t on which callbacks must enqueue their work */
void* diaglib_b200_stream(void);
int32_t diaglib_b200_last_status(void);
const char* diaglib_b200_last_message(void);

/* install the local rows of a CSR matrix (HOST arrays, copied to HBM).  Column indices are
 * LOCAL: 0..n_loc-1 address owned rows, n_loc..n_loc+n_halo-1 address the halo block filled
 * by the exchange plan below.  diag = the matrix diagonal of the owned rows (preconditioner). */
int32_t diaglib_b200_set_csr(int64_t n_loc, int64_t n_halo, const int64_t* rowptr, const int32_t* col,
                             const double* val, const double* diag);
/* set_csr for a matrix that already lives in HBM (device pointers; the caller keeps ownership and
 * keeps the arrays alive while the matrix is installed). */
int32_t diaglib_b200_set_csr_device(int64_t n_loc, int64_t n_halo, int64_t nnz, const int64_t* rowptr_dev,
                                    const int32_t* col_dev, const double* val_dev, const double* diag_dev);
/* optional processing order of the local rows in the built-in matvec: a permutation of
 * [0, n_loc) (host array), null = natural order.  A locality-preserving order (tiles of a
 * stencil's grid along a space-filling curve) raises the cache hit rate of the gathers; results
 * are independent of it.  Call after set_csr / set_csr_device.  The library keeps the rows that
 * reference halo columns at the end of the order so that the halo exchange overlaps the others. */
int32_t diaglib_b200_set_csr_row_order(const int32_t* order);
/* metric B of the generalized problem for the built-in bvec.  Same conventions as set_csr; if
 * it has halo columns (n_halo > 0) they use the matrix's halo numbering and exchange plan. */
int32_t diaglib_b200_set_csr_b(int64_t n_loc, int64_t n_halo, const int64_t* rowptr, const int32_t* col,
                               const double* val);
/* linear-response matrices for the built-in products: which = 0 (A+B), 1 (A-B), 2 (S+D), 3 (S-D);
 * same conventions as set_csr_b.  set_lr_diag installs diag(A) and diag(S) for the built-in lrprec. */
int32_t diaglib_b200_set_csr_lr(int32_t which, int64_t n_loc, int64_t n_halo, const int64_t* rowptr, const int32_t* col,
                                const double* val);
int32_t diaglib_b200_set_lr_diag(int64_t n_loc, const double* aa_diag, const double* sigma_diag);
/* halo exchange plan: for neighbour i, send owned rows [send_row0[i], +send_cnt[i]) and
 * receive recv_cnt[i] rows into halo rows [recv_off[i], ...). */
int32_t diaglib_b200_set_halo(int32_t n_nbr, const int32_t* peer, const int64_t* send_row0,
                              const int64_t* send_cnt, const int64_t* recv_off, const int64_t* recv_cnt);

/* multi-GPU: one process per GPU.  Rank 0 obtains an id, the host program distributes it
 * (MPI / torch.distributed / files), every rank calls comm_init. */
int32_t diaglib_b200_comm_unique_id(void* out_128_bytes);
int32_t diaglib_b200_comm_init(int32_t rank, int32_t nranks, const void* unique_id_128_bytes);
int32_t diaglib_b200_comm_rank(void);
int32_t diaglib_b200_comm_size(void);
/* How the k x k all-reduces of the drivers (diaglib.f90 has none: they replace the serial dgemm('t','n')
 * results at 403, 1691, 3256, 3543 on a row-partitioned block) travel:
 *   out4[0] ranks sharing the peer window (0: every all-reduce is an ncclAllReduce; the window needs
 *           cudaIpc between the ranks of one node, DIAGLIB_B200_PEER_REDUCE=0 disables it)
 *   out4[1] all-reduces of the last driver call that went through the window
 *   out4[2] 1 when a peer's contribution timed out (results invalid)   out4[3] completed window calls */
void diaglib_b200_peer_info(int64_t* out4);

/* ---- introspection used by the tests and bench.py ------------------------------------- */

/* per-iteration history of the last driver call (same record the reference prints when
 * verbose, diaglib.f90:459-464,1750-1755, for all n_max roots) */
int32_t diaglib_b200_history_len(void);
void diaglib_b200_history_get(int32_t* it, int32_t* n_act, double* eig, double* rms, double* max, int32_t* done);
/* seconds: [0] matvec [1] reduced eigensolve [2] orthogonalisation [3] total
 *          [4] gram [5] ritz/projection [6] residual+precnd [7] h2d+d2h staging
 * and, when profiling is on, per kernel family:
 *          [8] gram kernels [9] block_mul/trmm kernels [10] unused [11] block copies */
void diaglib_b200_timers(double* out12);
/* per-kernel-family device timing (two events per launch); off by default */
void diaglib_b200_set_profile(int32_t on);
/* counters: [0] ortho_cd passes [1] ortho_vs_x sweeps [2] QR fallbacks [3] Cholesky shifts
 *           [4] kernel launches issued by the library during the last driver call */
void diaglib_b200_stats(int64_t* out8);

#ifdef __cplusplus
}
#endif
#endif /* DIAGLIB_B200_H */
