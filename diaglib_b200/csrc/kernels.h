// Internal C++ interface of the sm_100a kernels (device pointers, explicit stream).
// Column-major everywhere, leading dimension in elements (Fortran convention of the
// reference: explicit-shape x(n,*) blocks, diaglib.f90:221-228).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace dlb {

// number of kernels launched by this library since load (bench.py reports it as gpu_launches)
extern int64_t g_launches;
extern bool g_disable_ws;
extern int g_ws_mask;
extern int g_dbg;
extern bool g_disable_tma;
extern bool g_disable_fused_gram;
extern bool g_bmul_small_tiles;
extern bool g_eig_two_sided;
extern int g_eig_coop_min_k;
extern int g_eig_mode;
extern int g_coeffs_threads;
extern int g_coeffs_smem;
extern int g_eig_block;
void set_chol_blocked(int on);   // 1 (default): blocked factor-and-invert inside chol_inv / get_coeffs; 0: column by column
// Predicate of the dense kernels (gram_tn, block_mul, block_trmm_inplace, block_mul_gram): while it
// points to a device int, every kernel these wrappers launch returns at once when that int is 0.
// The engine sets it around the steps of a speculative ortho_cd / ortho_vs_x chain, whose control
// flow (another pass? another sweep?) is decided on the device by chol_inv (OrthoCtl below).
extern const int* g_live;
extern int g_spmm_short;
extern int g_spmm_chunk;
extern int g_spmm_chunk_tiled;

// ---- peer window: reduction of the per-CTA partials fused with the all-reduce over NVLink ------
// Every rank owns a window of device memory that all ranks of the node map (cudaIpc): two parity
// halves of nranks slots of PEER_CAP doubles, plus one arrival flag per sender.  One kernel reduces
// the local partials, stores the result into its slot of EVERY rank's window (peer stores through
// NVSwitch), publishes an epoch number to every rank's flags, waits for the nranks flags of its own
// window and adds the slots in rank order - so the k x k result is bit-identical on all ranks,
// with one launch instead of gram_reduce + ncclAllReduce.  Set up by diaglib_b200_comm_init.
constexpr int PEER_MAX = 8;
constexpr int PEER_CAP = 16384;   // doubles per slot: a 128 x 128 block
struct PeerState {
  unsigned int arrive, depart;     // CTAs past the store phase / past the sum phase of the running call
  unsigned long long epoch;        // number of completed calls (same on every rank: the calls are collective)
  unsigned int error, pad;         // 1: a peer's flag did not arrive within the time-out
};
struct PeerWin {
  int nranks = 0, rank = 0;                       // nranks == 0: not available (single rank, or set-up failed)
  double* data[PEER_MAX] = {};                    // data[r]: window of rank r as mapped into this process
  unsigned long long* flags[PEER_MAX] = {};       // flags[r][s]: last epoch sender s has published to rank r
  PeerState* state = nullptr;                     // local
};
extern PeerWin g_peerwin;
extern bool g_fuse_allreduce;   // gram_tn: finish every block with the peer all-reduce (set by the engine around the call)
// in-place all-reduce of d[0, count) (count <= PEER_CAP) through the peer window: sum of the elements
// below max_from, maximum of the others.  Collective over the ranks, predicated by g_live like gram_tn.
void peer_allreduce(cudaStream_t st, double* d, int count, int max_from);

// ---- dense.cu -------------------------------------------------------------------------
// C(p x q, ldc) = A(n x p, lda)^T * B(n x q, ldb).  Replaces dgemm('t','n',p,q,n,...) at
// diaglib.f90:313,403,1691,3256,3543,3762.  If sym_lower only tiles on/below the diagonal
// are computed and the result is mirrored (callers consume one triangle: dsyev 'l' 406,
// dpotrf 'l' 3261).  `partial` is scratch of at least gram_scratch_bytes(p,q,num_sms).
size_t gram_scratch_bytes(int p, int q, int num_sms);
// tile schedule of the Gram kernels for ntp x ntq tiles of 8 x 8 (host logic only; used by the tests)
int gram_schedule_cover(int ntp, int ntq, bool sym_lower, int nwarps, int* cover, int* load4);
constexpr int GRAM_CONSUMER_WARPS = 15;   // consumer warps of the TMA / bulk-copy Gram kernels
void gram_tn(cudaStream_t st, int num_sms, int64_t n, const double* A, int64_t lda, int p, const double* B,
             int64_t ldb, int q, double* C, int ldc, bool sym_lower, double* partial);

// Y(n x q, ldy) = alpha * V(n x p, ldv) * C(p x q, ldc) + beta * Y.  Y may alias V when the
// product is row-local (q <= 128).  Replaces dgemm('n','n',n,q,p,...) at
// diaglib.f90:322,324,420,421,495,497,1717,1721,3544 and dtrmm('r','l','t','n') at 3327
// (with C = L^-T stored as a full matrix with an explicit zero triangle).
// ident_from >= 0 is a hint: rows [ident_from, p) of C are the q x q identity (Y = V1 C1 + V2 with
// V2 the last q columns of V), so the tensor pipe only visits the tiles on that diagonal.
// ident_is_tri: that block of C is upper triangular instead (Y = V1 C1 + V2 T), tiles below its diagonal are skipped.
void block_mul(cudaStream_t st, int64_t n, const double* V, int64_t ldv, int p, const double* C, int ldc, int q,
               double alpha, double beta, double* Y, int64_t ldy, bool upper_tri = false, int ident_from = -1,
               bool ident_is_tri = false);

// U <- U * T, T upper triangular m x m (ld m), in place (dtrmm at diaglib.f90:3327).
void block_trmm_inplace(cudaStream_t st, int64_t n, double* U, int64_t ldu, int m, const double* T);

// block_mul fused with the metric of its result: G (q x q, ldg) = Y^T Y after Y = alpha V C + beta Y.
// partial: scratch of gram_scratch_bytes(128,128,2*num_sms) (one q x q partial per CTA).
void block_mul_gram(cudaStream_t st, int num_sms, int64_t n, const double* V, int64_t ldv, int p, const double* C,
                    int ldc, int q, double alpha, double beta, double* Y, int64_t ldy, bool upper_tri, double* G,
                    int ldg, double* partial);

// ---- sparse.cu ------------------------------------------------------------------------
struct CsrDevice {
  int64_t n = 0;        // local rows
  int64_t nnz = 0;
  int64_t n_halo = 0;   // columns >= n index the halo block
  int max_row_nnz = 0;  // longest row (0 = unknown); selects the short-row SpMM
  // optional processing order of the rows (a permutation of [0, n)): entries [0, n_interior) are
  // rows without halo columns, the rest touch the halo.  null: natural order, n_interior = n
  // when n_halo = 0 and 0 otherwise.  tiled: the order keeps neighbouring rows of a stencil
  // together (set by the caller), so the SpMM sweeps all columns in one launch.
  const int32_t* order = nullptr;
  int64_t n_interior = 0;
  bool tiled = false;
  int64_t* rowptr = nullptr;
  int32_t* col = nullptr;
  double* val = nullptr;
  double* diag = nullptr;
};
// AX(n x m) = A * X (+ shift * X).  x_halo holds the remote rows (n_halo x m, ld n_halo).
// Built-in conforming matvec(n,m,x,ax) (contract diaglib.f90:66; toy impl main.f90:72-90).
// part: all rows, or only the rows without / with halo columns (A.order's two sections), so that
// the interior rows can run while the halo exchange is in flight.
enum { SPMM_ALL = 0, SPMM_INTERIOR = 1, SPMM_BOUNDARY = 2 };
void spmm_csr(cudaStream_t st, const CsrDevice& A, int m, const double* x, int64_t ldx, const double* x_halo,
              double* ax, int64_t ldax, double shift, int part = SPMM_ALL);
// px = x / (d + fac) where |d + fac| > 1e-5, else x  (main.f90:146-171).
void diag_precnd(cudaStream_t st, int64_t n, int m, double fac, const double* diag, const double* x, int64_t ldx,
                 double* px, int64_t ldpx);
// r(:,j) = ax(:,j) - theta[j] * x(:,j) for the columns with active[j] != 0 (others: r = ax),
// plus per-column sum of squares and max |r| of the active columns
// (dcopy 428 + daxpy 438 + dnrm2 440 + maxval 441; Davidson 1729-1731).
// norms_out[2*j] = sum r^2 (local rows), norms_out[2*j+1] = max |r|.  theta/active are device arrays.
void residual_norms(cudaStream_t st, int num_sms, int64_t n, int m, const double* ax, int64_t ldax, const double* x,
                    int64_t ldx, const double* theta, const int* active, double* r, int64_t ldr,
                    double* norms_out, double* scratch);
size_t residual_scratch_bytes(int m, int num_sms);
// y += a * x on an n x m block (daxpy 312,397)
void block_axpy(cudaStream_t st, int64_t n, int m, double a, const double* x, int64_t ldx, double* y, int64_t ldy);
// y = x on an n x m block (dcopy)
void block_copy(cudaStream_t st, int64_t n, int m, const double* x, int64_t ldx, double* y, int64_t ldy);
// gathers rows [row0,row0+cnt) of an n x m block into a dense cnt x m buffer (halo packing)
void pack_rows(cudaStream_t st, int64_t row0, int64_t cnt, int m, const double* x, int64_t ldx, double* out);

// linear-response helpers (caslr_eff_driver 1190-1193, 1333-1336; lrprec_2 of main.f90:257-281)
void lr_split(cudaStream_t st, int64_t n, int m, const double* evec, int64_t ld2, double* vp, double* vm, int64_t ldv);
void lr_merge(cudaStream_t st, int64_t n, int m, const double* ep, const double* em, int64_t lde, double* evec, int64_t ld2);
void lr_precnd(cudaStream_t st, int64_t n, int m, double fac, const double* aa, const double* sg, const double* xp,
               const double* xm, double* yp, double* ym);

// ---- small.cu (single-CTA dense kernels, replicated per rank) --------------------------
struct CholStatus {      // written by chol_inv, read back by the host control loop
  double l_norm, linv_norm, shift_used, unorm;
  int info_first;        // dpotrf info of the un-shifted attempt (0 = fine)
  int n_shifts;          // level-shift retries used (3265-3295)
  int hard_fail;         // shift loop exhausted (3276-3284)
  int pad;
};
// Device-side control of a speculative ortho_cd / ortho_vs_x chain (diaglib.f90:3246-3333,
// 3533-3568): the host enqueues the passes and sweeps it expects without reading anything back;
// each step group runs only if its cell of `live` is set, and the only kernel that sets cells is
// chol_inv, which holds the reference's decisions (macro_done 3331-3332, xu_norm < tol 3562-3566).
struct OrthoCtl {
  int live[48];          // live[g] != 0: the kernels of step group g run
  int halt;              // 0 = ran to the end of what was decided; 2 = Cholesky shift loop exhausted (3276-3284);
                         // 3 = the passes enqueued for an ortho_cd did not reach macro_done
  int done_vsx;          // 1: xu_norm = growth * eps fell below tol_ortho (ortho_vs_x finished)
  int last_phase;        // ortho_cd instance of the last factorisation that ran (0 = the initial one, s = after sweep s)
  int last_pass;         // its pass number (1-based)
  int passes, shifts;    // totals of the chain (statistics)
  int pdone[8];          // per ortho_cd instance: the pass that reached macro_done (0 = none yet)
  int deferred;          // 1: the triangular multiply of the last ortho_cd that ran was NOT applied (CholLink::defer_ok):
                         //    its T waits to be folded into the projection step of the sweep that follows
  double growth;         // of the current ortho_cd (3323)
};
struct CholLink {        // where a chol_inv call sits in the chain (ctl = null: plain call)
  OrthoCtl* ctl = nullptr;
  int self = 0;          // cell that enables this pass (gram + chol_inv)
  int trmm = 0;          // cell chol_inv sets for the triangular multiply of this pass
  int next_pass = -1;    // cell of the next pass of the same ortho_cd (-1: none enqueued)
  int next_head = -1;    // cell of the next sweep's projection step (-1: none enqueued)
  int next_first = -1;   // cell of the first pass of the ortho_cd after that sweep
  int phase = 0, pass = 1;
  int check_vsx = 0;     // this ortho_cd closes a sweep of ortho_vs_x: test xu_norm when it is done
  int defer_ok = 0;      // inside ortho_vs_x with u right behind x: when this ortho_cd is done and another sweep follows,
                         // leave its last triangular multiply to that sweep's projection step (engine.cu, project_out)
};
// metric (m x m, ldm) -> Linv_t_full (m x m, ld m): the matrix T = L^-T (upper triangular,
// explicit zeros below) such that U_ortho = U * T.  Follows dpotrf('l') 3261, the shift loop
// 3265-3295, dtrtri('l','n') 3310 and norm_est 3314-3315.
// `work` must hold 2*m*m doubles.
void chol_inv(cudaStream_t st, int m, const double* metric, int ldm, double* T, double* work, CholStatus* status_dev,
              const CholLink& link = CholLink());

struct EigStatus { int sweeps; int converged; int path; };  // path: 1 = one-sided on the Cholesky factor, 2 = two-sided
// Symmetric eigensolver replacing dsyev('v',uplo) (315,406,1708): a (k x k, lda) is
// overwritten by the eigenvectors (ascending eigenvalues in w).  Parallel cyclic Jacobi.
// `work` must hold 2*kp*kp + 4*kp doubles with kp = k rounded up to even.
size_t eig_work_doubles(int k);
void sym_eig(cudaStream_t st, int k, double* a, int lda, bool upper, double* w, double* work, EigStatus* status_dev);

// get_coeffs (3686-3732) entirely on device: a_red (len_a x len_a, eigenvectors in the
// first len_u rows/cols) -> u_p (len_u x n_act, ld len_u).  u_x is read in place from a_red.
struct CoeffStatus { int sweeps; int cd_passes; int fail; int qr; };
size_t coeffs_work_doubles(int len_u, int n_max, int n_act);
void get_coeffs(cudaStream_t st, int len_a, int len_u, int n_max, int n_act, const double* a_red, double* u_p,
                double* work, CoeffStatus* status_dev);

// cp ((m + k) x k, ldc) = [-xu (m x k, ldx); I_k]: coefficients that turn u <- u - x xu (3544) into one
// product over the adjacent blocks [x u]
void proj_coeff(cudaStream_t st, int m, int k, const double* xu, int ldx, double* cp, int ldc, const double* T = nullptr);

// reduced problem of caslr_eff_driver: c = a^T a (1303); eig(i) = sqrt(e(k-1-i)), up(:,i) = z(:,k-1-i),
// um(:,i) = sred up(:,i) / eig(i) (1314-1324)
void small_ata(cudaStream_t st, int k, const double* a, int lda, double* c, int ldc);
void lr_reduced_vectors(cudaStream_t st, int k, int n_max, const double* z, int ldz, const double* e, const double* sred,
                        int lds, double* up, int ldup, double* um, int ldum, double* eig);

}  // namespace dlb
