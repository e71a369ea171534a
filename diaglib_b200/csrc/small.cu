// Single-CTA dense kernels for the k x k reduced problems of the diaglib hot path.  They are
// launched with one CTA and replicated on every rank: identical (all-reduced) input and
// deterministic code give bit-identical results everywhere, so no broadcast is needed.
//
//   chol_inv   : dpotrf('l') + level-shift retries + dtrtri('l','n') + norm_est  (ortho_cd,
//                diaglib.f90:3261-3316)
//   sym_eig    : dsyev('v',uplo) replacement (diaglib.f90:315,406,1708), parallel cyclic Jacobi
//   get_coeffs : diaglib.f90:3686-3732 including its ortho_vs_x / ortho_cd on the small
//                coefficient blocks, all inside one kernel
#include "common.cuh"
#include "kernels.h"

#include <cooperative_groups.h>

#include <algorithm>
#include <cfloat>

namespace dlb {
namespace {

constexpr int SM_THREADS = 1024;
constexpr double EPS = DBL_EPSILON;       // epsilon(one)
constexpr double TOL_ORTHO = 2.0 * EPS;   // diaglib.f90:151
__device__ int g_chol_blocked_dev = 1;   // 0: the unblocked factor-and-invert (tuning switch chol_blocked)
constexpr size_t CHOL_SMEM_MAX = 200 * 1024;   // factors of chol_inv stay in shared memory up to m = 112

// 1/sqrt(x) for x inside the float range: FP32 hardware seed + 3 Newton steps in FP64
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y = (double)rsqrtf((float)x);
  const double hx = 0.5 * x;
  y = y * (1.5 - hx * y * y);
  y = y * (1.5 - hx * y * y);
  y = y * (1.5 - hx * y * y);
  return y;
}
// 1/x for |x| inside the float range: FP32 seed + 3 Newton steps
__device__ __forceinline__ double fast_rcp(double x) {
  double y = (double)__frcp_rn((float)x);
  y = y * (2.0 - x * y);
  y = y * (2.0 - x * y);
  y = y * (2.0 - x * y);
  return y;
}

__device__ double cta_sum(double v, double* s_red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  const int nw = (blockDim.x + 31) >> 5;
  for (int w = 0; w < nw; ++w) t += s_red[w];
  return t;
}
__device__ double cta_max(double v, double* s_red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  const int nw = (blockDim.x + 31) >> 5;
  for (int w = 0; w < nw; ++w) t = fmax(t, s_red[w]);
  return t;
}

// norm_est, diaglib.f90:3447-3479 : max |diag| + Frobenius norm of the strict lower part
__device__ double cta_norm_est(int m, const double* a, int ld, double* s_red) {
  double dmax = 0.0, od = 0.0;
  for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
    const int i = e % m, j = e / m;
    const double v = a[i + (size_t)j * ld];
    if (i == j) dmax = fmax(dmax, fabs(v));
    else if (i > j) od = fma(v, v, od);
  }
  const double d = cta_max(dmax, s_red);
  const double o = cta_sum(od, s_red);
  return d + sqrt(o);
}

// In-place lower Cholesky of the lower triangle of L (m x m, ld).  Returns LAPACK-style info
// (0 ok, j+1 = first non-positive pivot).  Uniform across the CTA.
__device__ int cta_potrf_lower(int m, double* L, int ld) {
  for (int j = 0; j < m; ++j) {
    __syncthreads();
    const double ajj = L[j + (size_t)j * ld];
    if (!(ajj > 0.0)) return j + 1;  // also catches NaN, like dpotrf's disnan test
    const double s = sqrt(ajj);
    __syncthreads();
    for (int i = j + threadIdx.x; i < m; i += blockDim.x)
      L[i + (size_t)j * ld] = (i == j) ? s : L[i + (size_t)j * ld] / s;
    __syncthreads();
    const int rem = m - j - 1;
    for (int e = threadIdx.x; e < rem * rem; e += blockDim.x) {
      const int k = j + 1 + e / rem, i = j + 1 + e % rem;
      if (i >= k) L[i + (size_t)k * ld] = fma(-L[i + (size_t)j * ld], L[k + (size_t)j * ld], L[i + (size_t)k * ld]);
    }
  }
  __syncthreads();
  return 0;
}

// Li = inverse of the lower triangular L (one thread per column, forward substitution).
// Li is written in full (zeros above the diagonal).
__device__ void cta_trtri_lower(int m, const double* L, int ldl, double* Li, int ldi) {
  for (int j = threadIdx.x; j < m; j += blockDim.x) {
    for (int i = 0; i < j; ++i) Li[i + (size_t)j * ldi] = 0.0;
    Li[j + (size_t)j * ldi] = 1.0 / L[j + (size_t)j * ldl];
    for (int i = j + 1; i < m; ++i) {
      double s = 0.0;
      for (int k = j; k < i; ++k) s = fma(L[i + (size_t)k * ldl], Li[k + (size_t)j * ldi], s);
      Li[i + (size_t)j * ldi] = -s / L[i + (size_t)i * ldl];
    }
  }
  __syncthreads();
}

// The factor-and-invert step of one ortho_cd pass (diaglib.f90:3257-3316).
//   G (m x m, ldg) : metric, lower triangle used
//   L, Li          : m x m scratch (ld m)
//   T (m x m, ld m): output L^-T, upper triangular with explicit zeros
// unorm_sq = ||U||_F^2; the reference calls dnrm2(n*m,u,1) (3268), which equals
// sqrt(trace(G)) up to rounding, and the trace is already here.
// Lower Cholesky factor AND its inverse in one elimination: the row operations that reduce
// [L | I] to [I | L^-1] are applied while the columns of L are produced, so there is no separate
// (serial) triangular inversion and no per-element index division.  L, X: m x m in SHARED memory
// (ld m), on entry L = the matrix (lower triangle used), on exit L = factor, X = L^-1 (lower, zeros
// above).  buf: 2 m doubles of shared memory.  Returns dpotrf-style info (uniform).  Two CTA
// barriers per column; threads are arranged as (row = tid mod RS, column group = tid / RS) with RS
// the power of two >= m, so that the trailing updates need only shifts.
__device__ int cta_potrf_inv_smem(int m, double* L, double* X, double* buf) {
  const int tid = threadIdx.x, nt = blockDim.x;
  int rs_log = 5;
  while ((1 << rs_log) < m) ++rs_log;
  const int RS = 1 << rs_log, CG = nt >> rs_log;   // nt >= RS is guaranteed by the callers
  const int row = tid & (RS - 1), kq = tid >> rs_log;
  double* colbuf = buf;       // column j of the factor
  double* rowbuf = buf + m;   // row j of the inverse
  for (int e = tid; e < m * m; e += nt) X[e] = 0.0;
  __syncthreads();
  for (int i = tid; i < m; i += nt) X[i + (size_t)i * m] = 1.0;
  __syncthreads();
  for (int j = 0; j < m; ++j) {
    const double piv = L[j + (size_t)j * m];
    if (!(piv > 0.0)) return j + 1;   // uniform; also catches NaN like dpotrf's disnan test
    // (measured: multiplying by a Newton-refined reciprocal square root instead of dividing, and one
    // merged update loop instead of two, changed nothing at m = 37 (34.9 vs 33.8 us) and lost at
    // m = 74 (135 vs 97 us: the merged loop diverges inside a warp) - the step is bound by its two
    // barriers and the shared-memory round trips between them, not by the arithmetic)
    const double sq = sqrt(piv);
    for (int i = j + tid; i < m; i += nt) {
      const double v = (i == j) ? sq : L[i + (size_t)j * m] / sq;
      colbuf[i] = v;
      L[i + (size_t)j * m] = v;
    }
    for (int c = tid; c <= j; c += nt) {
      const double v = X[j + (size_t)c * m] / sq;
      rowbuf[c] = v;
      X[j + (size_t)c * m] = v;
    }
    __syncthreads();
    if (row > j && row < m && kq < CG) {
      const double li = colbuf[row];
      for (int k = j + 1 + kq; k <= row; k += CG) L[row + (size_t)k * m] = fma(-li, colbuf[k], L[row + (size_t)k * m]);
      for (int c = kq; c <= j; c += CG) X[row + (size_t)c * m] = fma(-li, rowbuf[c], X[row + (size_t)c * m]);
    }
    __syncthreads();
  }
  return 0;
}

// Blocked form of the same factor-and-invert step: 8-column blocks.  The diagonal block (8 x 8) is
// factored AND inverted by one warp in registers (row r in lane r, column index compile-time, the
// pivot row travels by shuffles: no barrier inside the block); the panel below it is a product with
// that inverse (rows independent), the trailing update a rank-8 update, and L^-1 is assembled block
// diagonal by block diagonal from the small inverses (X_IJ = -D_I^-1 sum_K L_IK X_KJ).  3 CTA
// barriers per 8 columns instead of 2 per column: the unblocked form spends 0.9 us per column
// (34 us at m = 37, and ortho_cd needs it ~10 times per iteration, get_coeffs 6 more).
// L, X: m x m in SHARED memory (ld m); the strict upper triangle of L is used as scratch.
constexpr int CHB = 8;
__device__ int cta_potrf_inv_blocked(int m, double* L, double* X) {
  __shared__ int s_info;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  int rs_log = 5;
  while ((1 << rs_log) < m) ++rs_log;
  const int RS = 1 << rs_log, CG = nt >> rs_log;   // nt >= RS is guaranteed by the callers
  const int row = tid & (RS - 1), kq = tid >> rs_log;
  for (int e = tid; e < m * m; e += nt) X[e] = 0.0;
  if (tid == 0) s_info = 0;
  __syncthreads();
  for (int j0 = 0; j0 < m; j0 += CHB) {
    const int nb = min(CHB, m - j0);
    if (warp == 0) {
      // lanes 8..31 mirror lanes 0..7 (same values), so every shuffle source is a valid lane
      const int r = lane & 7;
      double d[CHB], x[CHB];
#pragma unroll
      for (int c = 0; c < CHB; ++c) {
        d[c] = (r < nb && c <= r) ? L[(j0 + r) + (size_t)(j0 + c) * m] : (r == c ? 1.0 : 0.0);   // padding rows: identity
        x[c] = (r == c) ? 1.0 : 0.0;
      }
      int info = 0;
#pragma unroll
      for (int c = 0; c < CHB; ++c) {
        const double piv = __shfl_sync(0xffffffffu, d[c], c);
        if (!(piv > 0.0) && info == 0) info = j0 + c + 1;   // uniform; also catches NaN like dpotrf's disnan test
        const double rs = rsqrt(piv);
        // column c of the factor: d[c] is 0 above the diagonal and the pivot itself on it (piv * rs = sqrt(piv))
        const double lrc = d[c] * rs;
        d[c] = lrc;
        // [L | I] -> [I | L^-1]: row c of the inverse is scaled by 1 / l_cc, rows below take -l_rc times it.
        // One formula for all rows instead of selects (the block is a single warp's instruction stream):
        // x_new = x f - g xc with (f, g) = (0, -1) on row c, (1, l_rc) below it and (1, 0) above it
        const double f = (r == c) ? 0.0 : 1.0, g = (r == c) ? -1.0 : lrc;
#pragma unroll
        for (int cc = 0; cc <= c; ++cc) {
          const double xc = __shfl_sync(0xffffffffu, x[cc] * rs, c);
          x[cc] = fma(-g, xc, x[cc] * f);
        }
        // trailing columns of the block: rows r < c2 stay 0 because l2 is masked to 0 for them
#pragma unroll
        for (int c2 = c + 1; c2 < CHB; ++c2) {
          const double l2 = __shfl_sync(0xffffffffu, lrc, c2);
          d[c2] = fma(-lrc, (r >= c2) ? l2 : 0.0, d[c2]);
        }
      }
      if (info != 0) {
        if (lane == 0) s_info = info;
      } else if (lane < nb) {
#pragma unroll
        for (int c = 0; c < CHB; ++c)
          if (c <= r) {
            L[(j0 + r) + (size_t)(j0 + c) * m] = d[c];
            X[(j0 + r) + (size_t)(j0 + c) * m] = x[c];
          }
      }
    }
    __syncthreads();
    if (s_info != 0) return s_info;
    if (j0 + nb >= m) break;
    // panel below the block: L(i, j0 + c) = sum_{c' <= c} A(i, j0 + c') D^-1(c, c'), one thread per row
    for (int i = j0 + nb + tid; i < m; i += nt) {
      double a[CHB];
#pragma unroll
      for (int c = 0; c < CHB; ++c) a[c] = c < nb ? L[i + (size_t)(j0 + c) * m] : 0.0;
#pragma unroll
      for (int c = 0; c < CHB; ++c) {
        if (c < nb) {
          double sacc = 0.0;
#pragma unroll
          for (int c1 = 0; c1 <= c; ++c1) sacc = fma(a[c1], X[(j0 + c) + (size_t)(j0 + c1) * m], sacc);
          L[i + (size_t)(j0 + c) * m] = sacc;
        }
      }
    }
    __syncthreads();
    // trailing update: L(i, k) -= sum_c L(i, j0 + c) L(k, j0 + c), k >= j0 + nb, i >= k
    if (row >= j0 + nb && row < m && kq < CG) {
      double li[CHB];
#pragma unroll
      for (int c = 0; c < CHB; ++c) li[c] = c < nb ? L[row + (size_t)(j0 + c) * m] : 0.0;
      for (int k = j0 + nb + kq; k <= row; k += CG) {
        double sacc = L[row + (size_t)k * m];
#pragma unroll
        for (int c = 0; c < CHB; ++c)
          if (c < nb) sacc = fma(-li[c], L[k + (size_t)(j0 + c) * m], sacc);
        L[row + (size_t)k * m] = sacc;
      }
    }
    __syncthreads();
  }
  // L^-1 by block diagonals: X_IJ = -D_I^-1 S_IJ, S_IJ = sum_{K = J}^{I - 1} L_IK X_KJ  (I = J + dist).
  // S_IJ is parked in the transposed (strictly upper, unused) position of L.
  const int nblk = (m + CHB - 1) / CHB;
  for (int dist = 1; dist < nblk; ++dist) {
    const int nel = (nblk - dist) * CHB * CHB;
    for (int e = tid; e < nel; e += nt) {
      const int J = e >> 6, r = (e >> 3) & 7, c = e & 7;
      const int i = (J + dist) * CHB + r, cj = J * CHB + c;
      if (i < m) {   // cj < m follows
        double sacc = 0.0;
        const int t1 = (J + dist) * CHB;
        for (int t = cj; t < t1; ++t) sacc = fma(L[i + (size_t)t * m], X[t + (size_t)cj * m], sacc);   // X(t, cj) = 0 for t < cj
        L[cj + (size_t)i * m] = sacc;
      }
    }
    __syncthreads();
    for (int e = tid; e < nel; e += nt) {
      const int J = e >> 6, r = (e >> 3) & 7, c = e & 7;
      const int i0 = (J + dist) * CHB, i = i0 + r, cj = J * CHB + c;
      if (i < m) {
        double sacc = 0.0;
        for (int t = 0; t <= r; ++t) sacc = fma(X[i + (size_t)(i0 + t) * m], L[cj + (size_t)(i0 + t) * m], sacc);
        X[i + (size_t)cj * m] = -sacc;
      }
    }
    __syncthreads();
  }
  return 0;
}

// norm_est of a lower triangular matrix in shared memory without index divisions
__device__ double cta_norm_est_lower(int m, const double* a, double* s_red) {
  double dmax = 0.0, od = 0.0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = warp; j < m; j += nw) {
    for (int i = j + lane; i < m; i += 32) {
      const double v = a[i + (size_t)j * m];
      if (i == j) dmax = fmax(dmax, fabs(v));
      else od = fma(v, v, od);
    }
  }
  const double d = cta_max(dmax, s_red);
  const double o = cta_sum(od, s_red);
  return d + sqrt(o);
}

// cta_chol_inv with the factors in shared memory (m x m each) and `buf` = 2 m doubles of shared memory
__device__ void cta_chol_inv_smem(int m, const double* G, int ldg, double* L, double* Li, double* buf, double* T,
                                  CholStatus* st, double* s_red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  double tr = 0.0;
  for (int j = warp; j < m; j += nw)
    for (int i = lane; i < m; i += 32) {
      const double v = G[i + (size_t)j * ldg];
      L[i + (size_t)j * m] = v;
      if (i == j) tr += v;
    }
  const double unorm = sqrt(fmax(cta_sum(tr, s_red), 0.0));
  int info = g_chol_blocked_dev ? cta_potrf_inv_blocked(m, L, Li) : cta_potrf_inv_smem(m, L, Li, buf);
  const int info_first = info;
  int n_shifts = 0, hard_fail = 0;
  double shift = 0.0, alpha = 100.0;
  while (info != 0) {
    if (n_shifts >= 10) { hard_fail = 1; break; }  // 3276-3284
    ++n_shifts;
    shift = fmax(EPS * alpha * unorm, TOL_ORTHO);   // 3287
    __syncthreads();
    for (int j = warp; j < m; j += nw)
      for (int i = lane; i < m; i += 32) L[i + (size_t)j * m] = G[i + (size_t)j * ldg] + (i == j ? shift : 0.0);
    __syncthreads();
    info = g_chol_blocked_dev ? cta_potrf_inv_blocked(m, L, Li) : cta_potrf_inv_smem(m, L, Li, buf);
    alpha *= 10.0;
  }
  double l_norm = 0.0, linv_norm = 0.0;
  if (!hard_fail) {
    l_norm = cta_norm_est_lower(m, L, s_red);
    linv_norm = cta_norm_est_lower(m, Li, s_red);
    for (int j = warp; j < m; j += nw)
      for (int i = lane; i < m; i += 32) T[i + (size_t)j * m] = (i <= j) ? Li[j + (size_t)i * m] : 0.0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    st->l_norm = l_norm;
    st->linv_norm = linv_norm;
    st->shift_used = shift;
    st->unorm = unorm;
    st->info_first = info_first;
    st->n_shifts = n_shifts;
    st->hard_fail = hard_fail;
    st->pad = 0;
  }
  __syncthreads();
}

__device__ void cta_chol_inv(int m, const double* G, int ldg, double* L, double* Li, double* T, CholStatus* st,
                             double* s_red) {
  double tr = 0.0;
  for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
    const int i = e % m, j = e / m;
    const double v = G[i + (size_t)j * ldg];
    L[i + (size_t)j * m] = v;
    if (i == j) tr += v;
  }
  const double unorm = sqrt(fmax(cta_sum(tr, s_red), 0.0));
  int info = cta_potrf_lower(m, L, m);
  const int info_first = info;
  int n_shifts = 0, hard_fail = 0;
  double shift = 0.0;
  if (info != 0) {
    double alpha = 100.0;
    while (info != 0) {
      if (n_shifts >= 10) { hard_fail = 1; break; }  // 3276-3284
      ++n_shifts;
      shift = fmax(EPS * alpha * unorm, TOL_ORTHO);   // 3287
      __syncthreads();
      for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
        const int i = e % m, j = e / m;
        L[i + (size_t)j * m] = G[i + (size_t)j * ldg] + (i == j ? shift : 0.0);
      }
      info = cta_potrf_lower(m, L, m);
      alpha *= 10.0;
    }
  }
  double l_norm = 0.0, linv_norm = 0.0;
  if (!hard_fail) {
    cta_trtri_lower(m, L, m, Li, m);
    l_norm = cta_norm_est(m, L, m, s_red);
    linv_norm = cta_norm_est(m, Li, m, s_red);
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
      const int i = e % m, j = e / m;
      T[i + (size_t)j * m] = (i <= j) ? Li[j + (size_t)i * m] : 0.0;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    st->l_norm = l_norm;
    st->linv_norm = linv_norm;
    st->shift_used = shift;
    st->unorm = unorm;
    st->info_first = info_first;
    st->n_shifts = n_shifts;
    st->hard_fail = hard_fail;
    st->pad = 0;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(SM_THREADS)
chol_inv_kernel(int m, const double* G, int ldg, double* T, double* work, CholStatus* st, CholLink link) {
  __shared__ double s_red[32];
  extern __shared__ __align__(16) double dyn[];
  OrthoCtl* ctl = link.ctl;
  if (ctl && ctl->live[link.self] == 0) return;   // this pass was not decided (uniform)
  if ((2 * (size_t)m * m + 2 * m) * sizeof(double) <= CHOL_SMEM_MAX && (int)blockDim.x >= m) {
    // factors on chip: factor and inverse in one elimination
    cta_chol_inv_smem(m, G, ldg, dyn, dyn + (size_t)m * m, dyn + 2 * (size_t)m * m, T, st, s_red);
  } else {
    cta_chol_inv(m, G, ldg, work, work + (size_t)m * m, T, st, s_red);
  }
  if (ctl && threadIdx.x == 0) {
    // the reference's control flow of ortho_cd / ortho_vs_x, decided here instead of on the host
    ctl->passes += 1;
    ctl->shifts += st->n_shifts;
    ctl->last_phase = link.phase;
    ctl->last_pass = link.pass;
    if (st->hard_fail) { ctl->halt = 2; return; }                 // 3276-3284: nothing after this runs
    ctl->live[link.trmm] = 1;                                      // 3327
    ctl->deferred = 0;
    const double growth = (link.pass == 1 ? 1.0 : ctl->growth) * st->linv_norm;   // 3323
    ctl->growth = growth;
    const double rcond = st->l_norm * st->linv_norm;
    if (EPS * rcond * rcond < TOL_ORTHO) {                         // macro_done, 3331-3332
      ctl->pdone[link.phase] = link.pass;
      if (link.check_vsx && growth * EPS < TOL_ORTHO) { ctl->done_vsx = 1; return; }   // 3562-3566
      // another sweep of ortho_vs_x follows (enqueued here or continued by the host): its projection step applies
      // this T together with the projection, so the block is not rewritten now
      if (link.defer_ok) { ctl->live[link.trmm] = 0; ctl->deferred = 1; }
      if (link.next_head >= 0) { ctl->live[link.next_head] = 1; ctl->live[link.next_first] = 1; }
    } else if (link.next_pass >= 0) {
      ctl->live[link.next_pass] = 1;
    } else {
      ctl->halt = 3;                                               // the host continues this ortho_cd
    }
  }
}

// ---------------------------------------------------------------------------------------
// Symmetric eigensolver (dsyev replacement), all in one CTA, A and the eigenvector matrix in
// shared memory when 2*kp*lds doubles fit, else in `work` (L2 resident): two-sided
// parallel-order cyclic Jacobi.  It keeps RELATIVE accuracy on the graded positive definite
// reduced matrices of LOBPCG/Davidson (Ritz values ~10 next to ~1e7 from the W block), which
// the parity bar (1e-10 relative) needs.  A one-sided (Hestenes) variant on G = A was tried
// in round 1 (one barrier per round instead of three): it was no faster (the FP64 pipe of one
// SM, not the barriers, bounds a round) and only absolutely accurate (7e-10 relative error on
// the small Ritz values), so it was dropped.
// The rotation angle needs one division and two reciprocal square roots per pair; they are
// computed from FP32 hardware seeds refined by Newton steps in FP64 (a Jacobi rotation only has
// to be orthogonal to working precision, c^2 + s^2 = 1, its angle may be approximate).
// Eigenvalues ascending, eigenvectors with their largest component positive.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void rr_pair(int r, int idx, int kp, int& p, int& q) {
  // round-robin tournament: round r (0..kp-2), pair idx (0..kp/2-1)
  // (0 <= r, idx < kp - 1, so one conditional subtraction replaces each modulo)
  if (idx == 0) { p = kp - 1; q = r; }
  else {
    p = r + idx;
    if (p >= kp - 1) p -= kp - 1;
    q = r - idx;
    if (q < 0) q += kp - 1;
  }
  if (p > q) { const int t = p; p = q; q = t; }
}

__device__ void eig_load(int k, int kp, int lds, const double* a, int lda, int upper, double* A, double* Z) {
  for (int e = threadIdx.x; e < kp * kp; e += blockDim.x) {
    const int i = e % kp, j = e / kp;
    double v = 0.0;
    if (i < k && j < k) {
      const int lo = i < j ? i : j, hi = i < j ? j : i;
      v = upper ? a[lo + (size_t)hi * lda] : a[hi + (size_t)lo * lda];
    }
    A[i + (size_t)j * lds] = v;
    Z[i + (size_t)j * lds] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
}

// two-sided parallel-order cyclic Jacobi; eigenvalues on the diagonal of A at exit
__device__ int jacobi_two_sided(int kp, int lds, double* A, double* Z, double* rc, double* rs, int* rp, int* s_flag,
                                int* converged) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int half = kp / 2;
  int sweeps = 0;
  *converged = 0;
  while (sweeps < 40) {
    for (int r = 0; r < kp - 1; ++r) {
      if (tid < half) {
        int p, q;
        rr_pair(r, tid, kp, p, q);
        const double app = A[p + (size_t)p * lds], aqq = A[q + (size_t)q * lds], apq = A[p + (size_t)q * lds];
        double c = 1.0, s = 0.0;
        // rotate iff |apq| > eps sqrt(|app aqq|)  (compared squared: no square root needed)
        if (apq * apq > (EPS * EPS) * fabs(app) * fabs(aqq) && fabs(apq) > 1e-150) {
          const double tau = (aqq - app) / (2.0 * apq);
          double t;
          if (fabs(tau) < 1e8) {
            const double t2 = 1.0 + tau * tau;                          // in [1, 1e16]
            const double r = t2 * fast_rsqrt(t2);                       // sqrt(1 + tau^2)
            t = (tau >= 0.0 ? 1.0 : -1.0) * fast_rcp(fabs(tau) + r);
          } else {
            t = 0.5 / tau;                                              // sqrt(1 + tau^2) == |tau| in FP64
          }
          c = fast_rsqrt(1.0 + t * t);
          s = t * c;
          *s_flag = 1;
        }
        rc[tid] = c; rs[tid] = s; rp[2 * tid] = p; rp[2 * tid + 1] = q;
      }
      __syncthreads();
      {
        const int warp = tid >> 5, lane = tid & 31, nwarp = nt >> 5;
        for (int pr = warp; pr < half; pr += nwarp) {
          const double s = rs[pr];
          if (s == 0.0) continue;
          const double c = rc[pr];
          double* Ap = A + (size_t)rp[2 * pr] * lds;
          double* Aq = A + (size_t)rp[2 * pr + 1] * lds;
          double* Zp = Z + (size_t)rp[2 * pr] * lds;
          double* Zq = Z + (size_t)rp[2 * pr + 1] * lds;
          for (int i = lane; i < kp; i += 32) {
            const double ap = Ap[i], aq = Aq[i], zp = Zp[i], zq = Zq[i];
            Ap[i] = c * ap - s * aq;
            Aq[i] = s * ap + c * aq;
            Zp[i] = c * zp - s * zq;
            Zq[i] = s * zp + c * zq;
          }
        }
      }
      __syncthreads();
      {
        const int warp = tid >> 5, lane = tid & 31, nwarp = nt >> 5;
        for (int pr = warp; pr < half; pr += nwarp) {
          const double s = rs[pr];
          if (s == 0.0) continue;
          const double c = rc[pr];
          const int p = rp[2 * pr], q = rp[2 * pr + 1];
          for (int j = lane; j < kp; j += 32) {
            double* col = A + (size_t)j * lds;
            const double ap = col[p], aq = col[q];
            double np_ = c * ap - s * aq, nq_ = s * ap + c * aq;
            if (j == q) np_ = 0.0;
            if (j == p) nq_ = 0.0;
            col[p] = np_;
            col[q] = nq_;
          }
        }
      }
      __syncthreads();
    }
    ++sweeps;
    const int any = *s_flag;
    __syncthreads();
    if (tid == 0) *s_flag = 0;
    __syncthreads();
    if (!any) { *converged = 1; break; }
  }
  return sweeps;
}

__global__ void __launch_bounds__(SM_THREADS)
sym_eig_kernel(int k, double* a, int lda, int upper, double* w, double* work, int use_smem, int gated,
               EigStatus* st) {
  extern __shared__ __align__(16) double dyn[];
  __shared__ int s_flag;
  if (gated && st->path == 1) return;   // the one-sided solver already delivered (uniform)
  const int kp = (k + 1) & ~1;
  const int lds = kp | 1;  // odd stride: row accesses of the two-sided form are bank-conflict free
  const int half = kp / 2;
  double* A = use_smem ? dyn : work;
  double* Z = A + (size_t)kp * lds;
  double* ev = Z + (size_t)kp * lds;   // kp eigenvalues
  double* rc = ev + kp;                // half cosines
  double* rs = rc + half;              // half sines
  int* rp = reinterpret_cast<int*>(rs + half);   // 2*half ints (p, q)
  int* rank = rp + 2 * half;                     // kp ints
  const int tid = threadIdx.x, nt = blockDim.x;

  if (tid == 0) s_flag = 0;
  eig_load(k, kp, lds, a, lda, upper, A, Z);
  int converged = 0;
  const int sweeps = jacobi_two_sided(kp, lds, A, Z, rc, rs, rp, &s_flag, &converged);
  for (int i = tid; i < kp; i += nt) ev[i] = A[i + (size_t)i * lds];
  __syncthreads();

  // ascending order by rank, largest-magnitude component made positive, write back into a
  for (int i = tid; i < k; i += nt) {
    const double di = ev[i];
    int rk = 0;
    for (int j = 0; j < k; ++j) {
      const double dj = ev[j];
      rk += (dj < di || (dj == di && j < i)) ? 1 : 0;
    }
    double best = 0.0, sign = 1.0;
    for (int rr = 0; rr < k; ++rr) {
      const double v = Z[rr + (size_t)i * lds];
      if (fabs(v) > best) { best = fabs(v); sign = v < 0.0 ? -1.0 : 1.0; }
    }
    w[rk] = di;
    rank[i] = sign < 0.0 ? -(rk + 1) : (rk + 1);
  }
  __syncthreads();
  for (int e = tid; e < k * k; e += nt) {
    const int rr = e % k, i = e / k;
    const int rk = rank[i];
    const double sg = rk < 0 ? -1.0 : 1.0;
    a[rr + (size_t)((rk < 0 ? -rk : rk) - 1) * lda] = sg * Z[rr + (size_t)i * lds];
  }
  if (tid == 0) { st->sweeps = sweeps; st->converged = converged; st->path = 2; }
}

// ---------------------------------------------------------------------------------------
// Multi-CTA symmetric eigensolver for reduced problems that do not fit one SM's shared memory
// (k > 118: Davidson subspaces, LOBPCG with many roots).  Same two-sided parallel-order Jacobi,
// reformulated so that every pass over the matrix is a coalesced COLUMN access:
//     A' = J^T A J  ==>  A'(:,p) = R (c a_p - s a_q),  A'(:,q) = R (s a_p + c a_q)
// where R applies the row rotations of ALL pairs of the round to a k-vector (entry i is combined
// with its partner entry).  A and Z live in global memory (L2 resident: 2 k^2 doubles), the pairs
// of a round are spread over the warps of a cooperative grid, and a round is
//     angles (needs a_pp, a_qq, a_pq) -> grid.sync -> column updates -> grid.sync.
// Replicated per rank like the single-CTA solver; deterministic (no atomics in the data path).
// ---------------------------------------------------------------------------------------
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(256)
sym_eig_coop_kernel(int k, const double* a, int lda, int upper, double* A, double* Z, double* rot_c, double* rot_s,
                    int* partner, int* flags, int max_sweeps, int gated, EigStatus* st) {
  extern __shared__ __align__(16) double stash[];  // per warp: 2*kp doubles (b_p, b_q)
  cg::grid_group grid = cg::this_grid();
  if (gated && st->path == 1) return;   // uniform over the grid: nobody reaches a grid.sync
  const int kp = (k + 1) & ~1;
  const int half = kp / 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  (void)lane; (void)warp; (void)nwarp;
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + tid, gthreads = (size_t)gridDim.x * blockDim.x;
  double* bp = stash;       // b_p, b_q of the pair this CTA is updating
  double* bq = bp + kp;

  for (size_t e = gtid; e < (size_t)kp * kp; e += gthreads) {
    const int i = (int)(e % kp), j = (int)(e / kp);
    double v = 0.0;
    if (i < k && j < k) {
      const int lo = i < j ? i : j, hi = i < j ? j : i;
      v = upper ? a[lo + (size_t)hi * lda] : a[hi + (size_t)lo * lda];
    }
    A[e] = v;
    Z[e] = (i == j) ? 1.0 : 0.0;
  }
  if (gtid == 0) { flags[0] = 0; flags[1] = 0; }
  grid.sync();

  int sweeps = 0, converged = 0;
  while (sweeps < max_sweeps) {
    for (int r = 0; r < kp - 1; ++r) {
      // ---- angles: one thread per pair
      for (size_t pr = gtid; pr < (size_t)half; pr += gthreads) {
        int p, q;
        rr_pair(r, (int)pr, kp, p, q);
        const double app = __ldcg(&A[p + (size_t)p * kp]), aqq = __ldcg(&A[q + (size_t)q * kp]),
                     apq = __ldcg(&A[p + (size_t)q * kp]);
        double c = 1.0, s = 0.0;
        if (apq * apq > (EPS * EPS) * fabs(app) * fabs(aqq) && fabs(apq) > 1e-150) {
          const double tau = (aqq - app) / (2.0 * apq);
          double t;
          if (fabs(tau) < 1e8) {
            const double t2 = 1.0 + tau * tau;
            const double rr = t2 * fast_rsqrt(t2);
            t = (tau >= 0.0 ? 1.0 : -1.0) * fast_rcp(fabs(tau) + rr);
          } else {
            t = 0.5 / tau;
          }
          c = fast_rsqrt(1.0 + t * t);
          s = t * c;
          flags[0] = 1;
        }
        // v'[p] = c v[p] - s v[q] ; v'[q] = s v[p] + c v[q]
        rot_c[p] = c; rot_s[p] = -s; partner[p] = q;
        rot_c[q] = c; rot_s[q] = s;  partner[q] = p;
      }
      grid.sync();
      // ---- column updates: one CTA per pair (all threads over the rows: the loads of a column
      //      are all in flight at once, the L2 latency is paid once per sub-phase)
      for (int pr = blockIdx.x; pr < half; pr += gridDim.x) {
        int p, q;
        rr_pair(r, pr, kp, p, q);
        const double c = __ldcg(&rot_c[q]), s = __ldcg(&rot_s[q]);
        double* Ap = A + (size_t)p * kp;
        double* Aq = A + (size_t)q * kp;
        if (s != 0.0) {
          double* Zp = Z + (size_t)p * kp;
          double* Zq = Z + (size_t)q * kp;
          for (int i = tid; i < kp; i += blockDim.x) {
            const double x = __ldcg(&Ap[i]), y = __ldcg(&Aq[i]);
            const double u = __ldcg(&Zp[i]), v = __ldcg(&Zq[i]);
            bp[i] = c * x - s * y;
            bq[i] = s * x + c * y;
            Zp[i] = c * u - s * v;
            Zq[i] = s * u + c * v;
          }
        } else {
          // not rotated itself, but its columns still receive the row rotations of the other pairs
          for (int i = tid; i < kp; i += blockDim.x) { bp[i] = __ldcg(&Ap[i]); bq[i] = __ldcg(&Aq[i]); }
        }
        __syncthreads();
        for (int i = tid; i < kp; i += blockDim.x) {
          const int pi = __ldcg(&partner[i]);
          const double ci = __ldcg(&rot_c[i]), si = __ldcg(&rot_s[i]);
          double vp = ci * bp[i] + si * bp[pi];
          double vq = ci * bq[i] + si * bq[pi];
          if (s != 0.0) {
            if (i == q) vp = 0.0;   // the rotated pivot is exactly zero
            if (i == p) vq = 0.0;
          }
          Ap[i] = vp;
          Aq[i] = vq;
        }
        __syncthreads();
      }
      grid.sync();
    }
    ++sweeps;
    const int any = __ldcg(&flags[0]);
    grid.sync();
    if (gtid == 0) flags[0] = 0;
    grid.sync();
    if (!any) { converged = 1; break; }
  }
  if (gtid == 0) { st->sweeps = sweeps; st->converged = converged; st->path = 2; }
}

// eigenvalues = diag(A), ascending order, largest component positive, eigenvectors written to a
__global__ void __launch_bounds__(1024)
eig_finish_kernel(int k, int kp, const double* A, const double* Z, double* a, int lda, double* w, int* rank,
                  const EigStatus* gate) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (gate && gate->path == 1) return;
  for (int i = tid; i < k; i += nt) {
    const double di = A[i + (size_t)i * kp];
    int rk = 0;
    for (int j = 0; j < k; ++j) {
      const double dj = A[j + (size_t)j * kp];
      rk += (dj < di || (dj == di && j < i)) ? 1 : 0;
    }
    double best = 0.0, sign = 1.0;
    for (int rr = 0; rr < k; ++rr) {
      const double v = Z[rr + (size_t)i * kp];
      if (fabs(v) > best) { best = fabs(v); sign = v < 0.0 ? -1.0 : 1.0; }
    }
    w[rk] = di;
    rank[i] = sign < 0.0 ? -(rk + 1) : (rk + 1);
  }
  __syncthreads();
  for (int e = tid; e < k * k; e += nt) {
    const int rr = e % k, i = e / k;
    const int rk = rank[i];
    a[rr + (size_t)((rk < 0 ? -rk : rk) - 1) * lda] = (rk < 0 ? -1.0 : 1.0) * Z[rr + (size_t)i * kp];
  }
}

// ---------------------------------------------------------------------------------------
// Positive definite reduced problems (every a_red = V^T A V of a positive definite A): Cholesky
// factor + ONE-SIDED block Jacobi on the factor (Veselic-Hari).  With P^T A P = L L^T (P sorts
// the diagonal in descending order) the rotations orthogonalise the COLUMNS of L:
//     L V = U S   ==>   A = (P U) S^2 (P U)^T,
// so the eigenvalues are the squared column norms (relative accuracy ~ k eps on graded matrices:
// the rotations only ever see columns, never differences of matrix entries) and the
// eigenvectors are the normalised columns themselves - no accumulated Z, no row rotations, half
// the data of the two-sided kernels.  A pair is rotated while its dot product stands above its
// own rounding floor (not merely above eps |x||y|): with the descending sort the small columns
// of L vanish in the large rows, so the floor is far below eps in the cosine and the
// eigenvectors come out accurate relative to each eigenvalue (residual |A z - w z| ~ 1e-14 |w|
// on the graded LOBPCG matrices, measured; dsyev delivers eps |A|).
//
// Parallel form: the kp columns are split into nblk = 2 G blocks of b columns; a sweep is a
// round-robin tournament of the blocks (nblk - 1 rounds, CTA g owns block pair g of the round)
// plus one round for the pairs inside a block.  In a round a CTA stages its 2 b columns in
// shared memory and runs b inner rounds of b disjoint column pairs (one warp per pair, 3 dot
// products + 1 rotation), so a grid barrier is paid once per b inner rounds and the factor
// crosses L2 once per block round.  L lives in global memory (L2 resident).  The Cholesky
// factorisation runs in the same cooperative kernel: right-looking panels, CTA 0 factors a
// panel in shared memory, all CTAs update the trailing matrix (k <= 158: the whole matrix is
// one panel and never leaves CTA 0's shared memory).
// Not positive definite (pivot <= 0) or not converged: status.converged = 0 and `a` is left
// untouched; the two-sided kernels above then run as the fallback (gate argument).
// ---------------------------------------------------------------------------------------
struct OsjCtl {
  unsigned bar;                      // monotonic grid-barrier counter
  int fail;                          // Cholesky pivot <= 0 / NaN
  int pad[2];
  unsigned long long sweep_max[60];  // per sweep: largest rotated |x.y| / sum|x_i y_i| (double bits)
};
constexpr int OSJ_NB = 16;           // Cholesky panel width when the matrix does not fit one CTA

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// grid-wide barrier of a cooperative launch (all CTAs co-resident): one atomic per CTA on a
// monotonic counter.  `target` is the calling CTA's private running total.
__device__ __forceinline__ void osj_grid_bar(unsigned* ctr, unsigned& target) {
  __syncthreads();
  if (gridDim.x == 1) return;
  if (threadIdx.x == 0) {
    target += gridDim.x;
    // release: the CTA's writes (ordered before this thread by the barrier above) become visible
    // before the increment; acquire: nothing after the spin is satisfied from stale data
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(ctr) : "memory");
    while (ld_acquire_u32(ctr) < target) { }
  }
  __syncthreads();
}

// 1/sqrt(x) for x inside the float range: FP32 hardware seed (rel. error ~1e-7) + 2 Newton steps
// in FP64 (1.5e-14, then rounding level)
__device__ __forceinline__ double fast_rsqrt2(double x) {
  double y = (double)rsqrtf((float)x);
  const double hx = 0.5 * x;
  y = y * (1.5 - hx * y * y);
  y = y * (1.5 - hx * y * y);
  return y;
}

__global__ void __launch_bounds__(256, 1)
sym_eig_osj_kernel(int k, int b, int nblk, double* a, int lda, int upper, double* w, double* L, int ldl, double* gn,
                   int* perm, OsjCtl* ctl, int max_sweeps, int nb_chol, EigStatus* st) {
  extern __shared__ __align__(16) double S[];   // Jacobi: 2 b columns of ldl doubles; Cholesky: one panel
  __shared__ double s_max[8];
  const int kp = nblk * b;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
  const int gwarp = blockIdx.x * nwarp + warp, twarps = gridDim.x * nwarp;
  const size_t gtid = (size_t)blockIdx.x * nt + tid, gthreads = (size_t)gridDim.x * nt;
  unsigned bar_target = 0;
  const double tol = EPS * sqrt((double)k);

  // ---- 0a: permutation that sorts the diagonal in descending order (ties: lower index first)
  for (size_t i = gtid; i < (size_t)k; i += gthreads) {
    const double di = a[i + i * (size_t)lda];
    int rk = 0;
    for (int j = 0; j < k; ++j) {
      const double dj = a[j + (size_t)j * lda];
      rk += (dj > di || (dj == di && j < (int)i)) ? 1 : 0;
    }
    if (!(di > 0.0)) ctl->fail = 1;   // also NaN: the ranks below would collide
    else perm[rk] = (int)i;
  }
  osj_grid_bar(&ctl->bar, bar_target);
  if (*(volatile int*)&ctl->fail) { if (gtid == 0) { st->sweeps = 0; st->converged = 0; st->path = 0; } return; }
  // ---- 0b: L = lower triangle of P^T A P, zero elsewhere (padding columns included)
  for (size_t e = gtid; e < (size_t)ldl * kp; e += gthreads) {
    const int i = (int)(e % ldl), j = (int)(e / ldl);
    double v = 0.0;
    if (i < k && j < k && i >= j) {
      const int oi = __ldcg(&perm[i]), oj = __ldcg(&perm[j]);
      const int lo = oi < oj ? oi : oj, hi = oi < oj ? oj : oi;
      v = upper ? a[lo + (size_t)hi * lda] : a[hi + (size_t)lo * lda];
    }
    L[e] = v;
  }
  osj_grid_bar(&ctl->bar, bar_target);
  // ---- 0c: right-looking panel Cholesky
  for (int J = 0; J < k; J += nb_chol) {
    const int nb = min(nb_chol, k - J), rows = k - J;
    if (blockIdx.x == 0) {
      // panel (rows x nb, column c at S + c * rows) in shared memory, one warp per column
      for (int c = warp; c < nb; c += nwarp)
        for (int i = lane; i < rows; i += 32) S[c * rows + i] = __ldcg(&L[(J + i) + (size_t)(J + c) * ldl]);
      __syncthreads();
      bool bad = false;
      for (int c = 0; c < nb; ++c) {
        const double piv = S[c * rows + c];
        if (!(piv > 0.0)) { bad = true; break; }   // uniform: every thread reads the same value
        const double sq = sqrt(piv);
        __syncthreads();
        for (int i = c + tid; i < rows; i += nt) S[c * rows + i] = (i == c) ? sq : S[c * rows + i] / sq;
        __syncthreads();
        for (int c2 = c + 1 + warp; c2 < nb; c2 += nwarp) {
          const double l = S[c * rows + c2];
          for (int i = c2 + lane; i < rows; i += 32) S[c2 * rows + i] = fma(-S[c * rows + i], l, S[c2 * rows + i]);
        }
        __syncthreads();
      }
      if (bad) { if (tid == 0) ctl->fail = 1; }
      else {
        for (int c = warp; c < nb; c += nwarp)
          for (int i = lane; i < rows; i += 32) L[(J + i) + (size_t)(J + c) * ldl] = (i >= c) ? S[c * rows + i] : 0.0;
      }
    }
    osj_grid_bar(&ctl->bar, bar_target);
    if (*(volatile int*)&ctl->fail) { if (gtid == 0) { st->sweeps = 0; st->converged = 0; st->path = 0; } return; }
    if (J + nb < k) {
      // trailing update: column j (one warp) -= panel(:, 0:nb) * panel(j, 0:nb)^T, rows >= j
      for (int j = J + nb + gwarp; j < k; j += twarps) {
        double lj[OSJ_NB];
        const double mine = (lane < nb) ? __ldcg(&L[j + (size_t)(J + lane) * ldl]) : 0.0;
#pragma unroll
        for (int c = 0; c < OSJ_NB; ++c) lj[c] = __shfl_sync(0xffffffffu, mine, c);
        for (int i = (j & ~31) + lane; i < k; i += 32) {
          if (i < j) continue;
          double acc = __ldcg(&L[i + (size_t)j * ldl]);
#pragma unroll
          for (int c = 0; c < OSJ_NB; ++c)
            if (c < nb) acc = fma(-__ldcg(&L[i + (size_t)(J + c) * ldl]), lj[c], acc);
          L[i + (size_t)j * ldl] = acc;
        }
      }
      osj_grid_bar(&ctl->bar, bar_target);
    }
  }

  // ---- 1: one-sided block Jacobi sweeps on the columns of L
  int sweeps = 0, converged = 0;
  const int g = blockIdx.x;
  while (sweeps < max_sweeps) {
    double my_max = 0.0;
    for (int R = 0; R < nblk; ++R) {
      int P, Q;
      const bool cross = R < nblk - 1;
      if (cross) rr_pair(R, g, nblk, P, Q);
      else { P = 2 * g; Q = 2 * g + 1; }
      // stage the 2 b columns (warp w: column w of P and column w of Q)
      for (int c = warp; c < 2 * b; c += nwarp) {
        const double* src = L + (size_t)((c < b ? P * b + c : Q * b + (c - b))) * ldl;
        double* dst = S + (size_t)c * ldl;
        for (int i = lane; i < ldl; i += 32) dst[i] = __ldcg(&src[i]);
      }
      __syncthreads();
      const int n_inner = cross ? b : b - 1;
      bool dirty = false;
      for (int t = 0; t < n_inner; ++t) {
        for (int pr = warp; pr < b; pr += nwarp) {
          int cp, cq;
          if (cross) { cp = pr; cq = b + ((pr + t) & (b - 1)); }
          else {
            const int hb = b / 2, off = pr < hb ? 0 : b;
            rr_pair(t, pr % hb, b, cp, cq);
            cp += off; cq += off;
          }
          double* x = S + (size_t)cp * ldl;
          double* y = S + (size_t)cq * ldl;
          double app = 0.0, aqq = 0.0, apq = 0.0, sab = 0.0;
          for (int i = lane; i < ldl; i += 32) {
            const double xv = x[i], yv = y[i], xy = xv * yv;
            app = fma(xv, xv, app); aqq = fma(yv, yv, aqq); apq += xy; sab += fabs(xy);
          }
          app = warp_sum(app); aqq = warp_sum(aqq); apq = warp_sum(apq); sab = warp_sum(sab);
          // rotate while the dot product stands above its own rounding floor, sqrt(k) eps sum|x_i y_i|
          // (<= sqrt(k) eps |x||y|, the usual cosine test, with equality for ungraded columns): on
          // graded factors this orthogonalises a small column against a large one far below eps in
          // the cosine, which is what makes the eigenvectors accurate relative to each eigenvalue
          if (fabs(apq) > tol * sab) {
            // Jacobi angle for the pair: with d = (aqq - app)/2, h = hypot(d, apq), u = |d| + h:
            // cos = sqrt(u / 2h), sin = sign(d) apq / sqrt(2 h u)   (cos^2 + sin^2 = 1 identically)
            const double big = fmax(app, aqq);
            const int ex = (__double2hiint(big) >> 20) & 0x7ff;
            const double sc = __hiloint2double((2046 - ex) << 20, 0);   // big * sc in [1, 2)
            // largest rotated ratio of the sweep: only compared with 1e-10, single precision is plenty
            // (scaled so that neither term leaves the float range; an underflow reads as "large")
            my_max = fmax(my_max, (double)__fdividef((float)(fabs(apq) * sc), fmaxf((float)(sab * sc), 1e-37f)));
            const double bq = apq * sc, d = 0.5 * (aqq - app) * sc;
            const double h2 = fma(d, d, bq * bq);
            double cs, sn;
            if (h2 > 1e-30 && ex > 0 && ex < 2046) {
              const double rh = fast_rsqrt2(h2);
              const double u = fabs(d) + h2 * rh;
              const double w = 0.5 * u * rh;            // in [1/2, 1]
              const double rw = fast_rsqrt2(w);
              cs = w * rw;
              sn = (d >= 0.0 ? 0.5 : -0.5) * bq * rh * rw;
            } else {
              const double zeta = (aqq - app) / (2.0 * apq);
              const double tt = fabs(zeta) < 1e8 ? (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta)) : 0.5 / zeta;
              cs = 1.0 / sqrt(1.0 + tt * tt);
              sn = tt * cs;
            }
            for (int i = lane; i < ldl; i += 32) {
              const double xv = x[i], yv = y[i];
              x[i] = cs * xv - sn * yv;
              y[i] = sn * xv + cs * yv;
            }
            dirty = true;
          }
        }
        __syncthreads();
      }
      // write the block pair back if any of its columns changed
      if (__syncthreads_or(dirty ? 1 : 0)) {
        for (int c = warp; c < 2 * b; c += nwarp) {
          double* dst = L + (size_t)((c < b ? P * b + c : Q * b + (c - b))) * ldl;
          const double* src = S + (size_t)c * ldl;
          for (int i = lane; i < ldl; i += 32) dst[i] = src[i];
        }
      }
      if (R == nblk - 1) {
        // publish this CTA's largest relative off-diagonal of the sweep before the last barrier
        if (lane == 0) s_max[warp] = my_max;
        __syncthreads();
        if (tid == 0) {
          double m = 0.0;
          for (int q = 0; q < nwarp; ++q) m = fmax(m, s_max[q]);
          atomicMax(&ctl->sweep_max[sweeps], (unsigned long long)__double_as_longlong(m));
        }
      }
      osj_grid_bar(&ctl->bar, bar_target);
    }
    const double mx = __longlong_as_double((long long)__ldcg(&ctl->sweep_max[sweeps]));
    ++sweeps;
    // mx = largest |x.y| / sum|x_i y_i| that was rotated in this sweep (0: a clean sweep).  Quadratic
    // convergence: a sweep whose rotations were all below 1e-10 leaves ~1e-20, under every floor
    if (mx <= 1e-10) { converged = 1; break; }
  }
  if (!converged) { if (gtid == 0) { st->sweeps = sweeps; st->converged = 0; st->path = 0; } return; }

  // ---- 2: eigenvalues = squared column norms, ascending; eigenvectors = normalised columns,
  //         rows back in the caller's order, largest component positive
  for (int j = gwarp; j < k; j += twarps) {
    double s = 0.0;
    for (int i = lane; i < k; i += 32) { const double v = __ldcg(&L[i + (size_t)j * ldl]); s = fma(v, v, s); }
    s = warp_sum(s);
    if (lane == 0) gn[j] = s;
  }
  osj_grid_bar(&ctl->bar, bar_target);
  for (int j = gwarp; j < k; j += twarps) {
    const double gj = __ldcg(&gn[j]);
    int rk = 0;
    for (int i = lane; i < k; i += 32) {
      const double gi = __ldcg(&gn[i]);
      rk += (gi < gj || (gi == gj && i < j)) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rk += __shfl_xor_sync(0xffffffffu, rk, o);
    double best = -1.0, bval = 0.0;
    int bidx = 0x7fffffff;
    for (int i = lane; i < k; i += 32) {
      const double v = __ldcg(&L[i + (size_t)j * ldl]);
      const int oi = __ldcg(&perm[i]);
      if (fabs(v) > best || (fabs(v) == best && oi < bidx)) { best = fabs(v); bval = v; bidx = oi; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o), ov = __shfl_xor_sync(0xffffffffu, bval, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ob > best || (ob == best && oi < bidx)) { best = ob; bval = ov; bidx = oi; }
    }
    const double nrm = sqrt(gj) * (bval < 0.0 ? -1.0 : 1.0);
    for (int i = lane; i < k; i += 32) a[__ldcg(&perm[i]) + (size_t)rk * lda] = __ldcg(&L[i + (size_t)j * ldl]) / nrm;
    if (lane == 0) w[rk] = gj;
  }
  if (gtid == 0) { st->sweeps = sweeps; st->converged = 1; st->path = 1; }
}

// ---------------------------------------------------------------------------------------
// get_coeffs (diaglib.f90:3686-3732) in one CTA: u_p = u_x(:,active) - e_i, then
// ortho_vs_x(len_u, n_max, n_act, u_x, u_p) (3481-3574) with ortho_cd (3185-3341) inside.
// ---------------------------------------------------------------------------------------
// dot product with four independent accumulators: a single FMA chain over ~111 rows is bound by the
// dependent-issue latency of the FP64 pipe (ncu source view: these loops were a third of get_coeffs)
__device__ __forceinline__ double dot4(int n, const double* __restrict__ a, const double* __restrict__ b) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int r = 0;
  for (; r + 4 <= n; r += 4) {
    s0 = fma(a[r], b[r], s0);
    s1 = fma(a[r + 1], b[r + 1], s1);
    s2 = fma(a[r + 2], b[r + 2], s2);
    s3 = fma(a[r + 3], b[r + 3], s3);
  }
  for (; r < n; ++r) s0 = fma(a[r], b[r], s0);
  return (s0 + s1) + (s2 + s3);
}
// G(m x m) = U^T U for a small n x m block
__device__ void cta_gram(int n, int m, const double* U, int ldu, double* G) {
  for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
    const int i = e % m, j = e / m;
    if (i < j) continue;  // lower triangle + mirror
    const double* ui = U + (size_t)i * ldu;
    const double* uj = U + (size_t)j * ldu;
    const double s = dot4(n, ui, uj);
    G[i + (size_t)j * m] = s;
    G[j + (size_t)i * m] = s;
  }
  __syncthreads();
}

// returns ok (uniform); growth accumulated as in 3323
// (buf != null: G, L, Li, T live in shared memory and buf is the 2 m scratch of cta_chol_inv_smem)
__device__ bool cta_ortho_cd(int n, int m, double* U, int ldu, double* G, double* L, double* Li, double* T,
                             double* tmp, CholStatus* st, double* s_red, double& growth, int* passes, double* buf) {
  growth = 1.0;
  for (int it = 1;; ++it) {
    if (it > 10) return false;  // 3248-3254
    if (threadIdx.x == 0) ++(*passes);
    cta_gram(n, m, U, ldu, G);
    if (buf) cta_chol_inv_smem(m, G, m, L, Li, buf, T, st, s_red);
    else cta_chol_inv(m, G, m, L, Li, T, st, s_red);
    if (st->hard_fail) return false;
    const double l_norm = st->l_norm, linv_norm = st->linv_norm;
    const double rcond = l_norm * linv_norm;
    growth *= linv_norm;
    // U <- U T  (T upper triangular): out of place into tmp, then copy back
    for (int e = threadIdx.x; e < n * m; e += blockDim.x) {
      const int r = e % n, j = e / n;
      double s0 = 0.0, s1 = 0.0;
      int k = 0;
      for (; k + 1 <= j; k += 2) {
        s0 = fma(U[r + (size_t)k * ldu], T[k + (size_t)j * m], s0);
        s1 = fma(U[r + (size_t)(k + 1) * ldu], T[(k + 1) + (size_t)j * m], s1);
      }
      if (k <= j) s0 = fma(U[r + (size_t)k * ldu], T[k + (size_t)j * m], s0);
      tmp[e] = s0 + s1;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < n * m; e += blockDim.x) U[(e % n) + (size_t)(e / n) * ldu] = tmp[e];
    __syncthreads();
    if (EPS * rcond * rcond < TOL_ORTHO) return true;  // 3331-3332
  }
}

// modified Gram-Schmidt with one re-orthogonalisation: device stand-in for the Householder
// fallback `ortho` (3052-3092) on the small coefficient blocks.  Cold path.
__device__ void cta_mgs2(int n, int m, double* U, int ldu, double* s_red) {
  for (int j = 0; j < m; ++j) {
    double* uj = U + (size_t)j * ldu;
    for (int pass = 0; pass < 2; ++pass)
      for (int i = 0; i < j; ++i) {
        const double* ui = U + (size_t)i * ldu;
        double s = 0.0;
        for (int r = threadIdx.x; r < n; r += blockDim.x) s = fma(ui[r], uj[r], s);
        const double d = cta_sum(s, s_red);
        for (int r = threadIdx.x; r < n; r += blockDim.x) uj[r] = fma(-d, ui[r], uj[r]);
        __syncthreads();
      }
    double s = 0.0;
    for (int r = threadIdx.x; r < n; r += blockDim.x) s = fma(uj[r], uj[r], s);
    const double nrm = sqrt(cta_sum(s, s_red));
    for (int r = threadIdx.x; r < n; r += blockDim.x) uj[r] = uj[r] / nrm;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(SM_THREADS)
get_coeffs_kernel(int len_a, int len_u, int n_max, int n_act, const double* a_red, double* u_p_out, double* work,
                  int in_smem, CoeffStatus* cst) {
  __shared__ double s_red[32];
  __shared__ CholStatus s_chol;
  __shared__ int s_passes;
  extern __shared__ __align__(16) double dyn[];
  const int off_x = n_max - n_act;
  const double* u_x = a_red;  // len_u x n_max, ld ldx = len_a (eigenvectors left there by sym_eig)
  // working set: in shared memory when it fits (the factorisations are chains of dependent steps:
  // from global memory every step pays an L2 round trip, 0.98 ms per call at 111 x 37 against
  // ~0.1 ms on chip), else in `work`
  double* base = in_smem ? dyn : work;
  double* G = base;                                 // n_act^2
  double* L = G + (size_t)n_act * n_act;            // n_act^2
  double* Li = L + (size_t)n_act * n_act;           // n_act^2
  double* T = Li + (size_t)n_act * n_act;           // n_act^2
  double* xu = T + (size_t)n_act * n_act;           // n_max x n_act
  double* tmp = xu + (size_t)n_max * n_act;         // len_u x n_act
  double* u_p = in_smem ? tmp + (size_t)len_u * n_act : u_p_out;   // len_u x n_act (copied out at the end)
  double* buf = in_smem ? u_p + (size_t)len_u * n_act : nullptr;   // 2 n_act
  int ldx = len_a;
  if (in_smem == 2) {   // the eigenvectors too: every overlap / projection below walks them once per sweep
    double* ux_s = buf + 2 * (size_t)n_act;   // len_u x n_max
    for (int e = threadIdx.x; e < len_u * n_max; e += blockDim.x) {
      const int r = e % len_u, j = e / len_u;
      ux_s[e] = u_x[r + (size_t)j * len_a];
    }
    u_x = ux_s;
    ldx = len_u;
  }
  if (threadIdx.x == 0) s_passes = 0;
  // u_p = u_x(:,ind_x:n_max) with 1 removed from the x coefficient (3716-3722)
  for (int e = threadIdx.x; e < len_u * n_act; e += blockDim.x) {
    const int r = e % len_u, j = e / len_u;
    double v = a_red[r + (size_t)(off_x + j) * len_a];
    if (r == off_x + j) v -= 1.0;
    u_p[r + (size_t)j * len_u] = v;
  }
  __syncthreads();
  // ortho_vs_x(len_u, n_max, n_act, u_x, u_p)
  double growth = 1.0;
  int sweeps = 0, fail = 0, qr = 0;
  bool ok = cta_ortho_cd(len_u, n_act, u_p, len_u, G, L, Li, T, tmp, &s_chol, s_red, growth, &s_passes, buf);
  if (!ok) { cta_mgs2(len_u, n_act, u_p, len_u, s_red); ++qr; }
  bool done = false;
  while (!done) {
    ++sweeps;
    // xu = u_x^T u_p ; u_p -= u_x xu   (3543-3544)
    for (int e = threadIdx.x; e < n_max * n_act; e += blockDim.x) {
      const int i = e % n_max, j = e / n_max;
      xu[e] = dot4(len_u, u_x + (size_t)i * ldx, u_p + (size_t)j * len_u);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < len_u * n_act; e += blockDim.x) {
      const int r = e % len_u, j = e / len_u;
      double s0 = u_p[r + (size_t)j * len_u], s1 = 0.0;
      int i = 0;
      for (; i + 2 <= n_max; i += 2) {
        s0 = fma(-u_x[r + (size_t)i * ldx], xu[i + (size_t)j * n_max], s0);
        s1 = fma(-u_x[r + (size_t)(i + 1) * ldx], xu[(i + 1) + (size_t)j * n_max], s1);
      }
      if (i < n_max) s0 = fma(-u_x[r + (size_t)i * ldx], xu[i + (size_t)j * n_max], s0);
      tmp[e] = s0 + s1;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < len_u * n_act; e += blockDim.x) u_p[e] = tmp[e];
    __syncthreads();
    ok = cta_ortho_cd(len_u, n_act, u_p, len_u, G, L, Li, T, tmp, &s_chol, s_red, growth, &s_passes, buf);
    double xu_norm;
    if (!ok) {
      cta_mgs2(len_u, n_act, u_p, len_u, s_red);
      ++qr;
      double s = 0.0;
      for (int e = threadIdx.x; e < n_max * n_act; e += blockDim.x) {
        const int i = e % n_max, j = e / n_max;
        const double d = dot4(len_u, u_x + (size_t)i * ldx, u_p + (size_t)j * len_u);
        s = fma(d, d, s);
      }
      xu_norm = sqrt(cta_sum(s, s_red));
    } else {
      xu_norm = growth * EPS;  // 3562
    }
    done = xu_norm < TOL_ORTHO;
    if (sweeps > 10) { fail = 1; break; }  // 3568 (unconditional in the reference)
  }
  __syncthreads();
  if (in_smem)
    for (int e = threadIdx.x; e < len_u * n_act; e += blockDim.x) u_p_out[e] = u_p[e];
  if (threadIdx.x == 0) { cst->sweeps = sweeps; cst->cd_passes = s_passes; cst->fail = fail; cst->qr = qr; }
}

}  // namespace

void chol_inv(cudaStream_t st, int m, const double* metric, int ldm, double* T, double* work, CholStatus* status_dev,
              const CholLink& link) {
  static bool attr_set = false;
  if (!attr_set) {
    DLB_CUDA_CHECK(cudaFuncSetAttribute(chol_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CHOL_SMEM_MAX));
    attr_set = true;
  }
  const size_t need = (2 * (size_t)m * m + 2 * m) * sizeof(double);
  const size_t smem = need <= CHOL_SMEM_MAX ? need : 0;
  const int threads = m <= 32 ? 256 : (m <= 64 ? 512 : SM_THREADS);
  chol_inv_kernel<<<1, threads, smem, st>>>(m, metric, ldm, T, work, status_dev, link);
  ++g_launches;
  DLB_CUDA_CHECK(cudaGetLastError());
}

namespace {
struct OsjPlan { int b, G, nblk, kp, ldl, nb_chol, threads; size_t smem; bool ok; };
}
int g_eig_mode = 0;        // DIAGLIB_B200_EIG_MODE: 0 = one-sided on the Cholesky factor, two-sided as fallback; 1 = two-sided only
int g_eig_block = 0;       // DIAGLIB_B200_EIG_BLOCK: columns per block of the one-sided solver (0 = automatic)
static OsjPlan osj_plan(int k, int num_sms) {
  OsjPlan p{};
  p.b = g_eig_block > 0 ? g_eig_block : (k < 64 ? 4 : 8);
  if (p.b != 2 && p.b != 4 && p.b != 8) p.b = 8;
  p.G = (k + 2 * p.b - 1) / (2 * p.b);
  p.nblk = 2 * p.G;
  p.kp = p.nblk * p.b;
  p.ldl = (k + 31) & ~31;
  const size_t jac = (size_t)2 * p.b * p.ldl;
  const size_t whole = (size_t)k * k;
  const size_t cap = (size_t)200 * 1024 / sizeof(double);
  p.nb_chol = whole <= cap ? k : OSJ_NB;                       // whole matrix in CTA 0's shared memory when it fits
  const size_t chol = p.nb_chol == k ? whole : (size_t)OSJ_NB * k;
  p.smem = std::max(jac, chol) * sizeof(double);
  p.threads = 32 * p.b;
  p.ok = p.G <= num_sms && p.smem <= (size_t)200 * 1024 && k >= 1;
  return p;
}

size_t eig_work_doubles(int k) {
  const int kp = (k + 1) & ~1, lds = kp | 1;
  // single-CTA layout: A, Z (kp x lds) + eigenvalues + rotation tables;
  // cooperative layout: A, Z (kp x kp) + rot_c, rot_s (kp) + partner, rank (kp ints) + flags
  const size_t two_sided = 2 * (size_t)kp * lds + 6 * (size_t)kp + 16;
  // one-sided layout: L (ldl x kp8) + column norms + permutation + control block
  const int ldl = (k + 31) & ~31, kp8 = ((k + 15) / 16) * 16 + 16;
  const size_t one_sided = (size_t)ldl * kp8 + 2 * (size_t)kp8 + sizeof(OsjCtl) / sizeof(double) + 16;
  return std::max(two_sided, one_sided);
}

bool g_eig_two_sided = false;  // unused (kept for ABI of the A/B switch)
int g_eig_coop_min_k = 119;    // DIAGLIB_B200_EIG_COOP_MIN_K: reduced problems at least this large use the multi-CTA solver

void sym_eig(cudaStream_t st, int k, double* a, int lda, bool upper, double* w, double* work, EigStatus* status_dev) {
  static bool attr_set = false;
  static int num_sms = 0, coop_ok = 0;
  if (!attr_set) {
    DLB_CUDA_CHECK(cudaFuncSetAttribute(sym_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    DLB_CUDA_CHECK(cudaFuncSetAttribute(sym_eig_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DLB_CUDA_CHECK(cudaFuncSetAttribute(sym_eig_osj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int dev = 0;
    DLB_CUDA_CHECK(cudaGetDevice(&dev));
    DLB_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    DLB_CUDA_CHECK(cudaDeviceGetAttribute(&coop_ok, cudaDevAttrCooperativeLaunch, dev));
    attr_set = true;
  }
  const int kp = (k + 1) & ~1;
  int gated = 0;
  if (g_eig_mode == 0 && coop_ok) {
    const OsjPlan pl = osj_plan(k, num_sms);
    if (pl.ok) {
      // one-sided block Jacobi on the Cholesky factor; leaves status.converged = 0 (and `a` untouched)
      // when the matrix is not positive definite, in which case the gated two-sided solver below runs
      double* L = work;
      double* gn = L + (size_t)pl.ldl * pl.kp;
      int* perm = reinterpret_cast<int*>(gn + pl.kp);
      OsjCtl* ctl = reinterpret_cast<OsjCtl*>(gn + 2 * (size_t)pl.kp);
      DLB_CUDA_CHECK(cudaMemsetAsync(ctl, 0, sizeof(OsjCtl), st));
      int ki = k, bi = pl.b, nblk = pl.nblk, ldai = lda, up = upper ? 1 : 0, ldl = pl.ldl, max_sweeps = 40, nbc = pl.nb_chol;
      void* args[] = {&ki, &bi, &nblk, &a, &ldai, &up, &w, &L, &ldl, &gn, &perm, &ctl, &max_sweeps, &nbc, &status_dev};
      DLB_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)sym_eig_osj_kernel, dim3(pl.G), dim3(pl.threads), args, pl.smem, st));
      ++g_launches;
      gated = 1;
    }
  }
  const size_t need = (2 * (size_t)kp * (kp | 1) + 6 * (size_t)kp + 16) * sizeof(double);
  const bool fits = need <= 224 * 1024;
  if (coop_ok && (k >= g_eig_coop_min_k || !fits)) {
    // multi-CTA solver: one CTA per pair and round, stash = 2 kp doubles
    const int threads = kp <= 128 ? 128 : 256;
    const int warps = threads / 32;
    const size_t smem = (size_t)2 * kp * sizeof(double);
    if (smem <= 200 * 1024) {
      const int half = kp / 2;
      int grid = std::max(1, std::min(num_sms, half));
      double* A = work;
      double* Z = A + (size_t)kp * kp;
      double* rot_c = Z + (size_t)kp * kp;
      double* rot_s = rot_c + kp;
      int* partner = reinterpret_cast<int*>(rot_s + kp);
      int* rank = partner + kp;
      int* flags = rank + kp;
      int max_sweeps = 40, ki = k, ldai = lda, up = upper ? 1 : 0;
      const double* ain = a;
      void* args[] = {&ki, &ain, &ldai, &up, &A, &Z, &rot_c, &rot_s, &partner, &flags, &max_sweeps, &gated, &status_dev};
      DLB_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)sym_eig_coop_kernel, dim3(grid), dim3(warps * 32), args, smem, st));
      ++g_launches;
      eig_finish_kernel<<<1, 1024, 0, st>>>(k, kp, A, Z, a, lda, w, rank, gated ? status_dev : nullptr);
      ++g_launches;
      DLB_CUDA_CHECK(cudaGetLastError());
      return;
    }
  }
  const int use_smem = fits ? 1 : 0;
  sym_eig_kernel<<<1, SM_THREADS, use_smem ? need : 0, st>>>(k, a, lda, upper ? 1 : 0, w, work, use_smem, gated, status_dev);
  ++g_launches;
  DLB_CUDA_CHECK(cudaGetLastError());
}

size_t coeffs_work_doubles(int len_u, int n_max, int n_act) {
  return 4 * (size_t)n_act * n_act + (size_t)n_max * n_act + (size_t)len_u * n_act + 8;
}

void set_chol_blocked(int on) {
  DLB_CUDA_CHECK(cudaMemcpyToSymbol(g_chol_blocked_dev, &on, sizeof(int)));
}
int g_coeffs_threads = 0;   // 0: automatic (512 / 1024 threads); experiment switch (diaglib_b200_k_set_tuning "coeffs_threads")
int g_coeffs_smem = 1;      // 0: working set of get_coeffs in global memory (round-1 behaviour)
void get_coeffs(cudaStream_t st, int len_a, int len_u, int n_max, int n_act, const double* a_red, double* u_p,
                double* work, CoeffStatus* status_dev) {
  static bool attr_set = false;
  if (!attr_set) {
    DLB_CUDA_CHECK(cudaFuncSetAttribute(get_coeffs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  // 512 threads up to 64 active columns (the CTA barriers of the factorisations are cheaper and every loop still has
  // a thread per element: 511 vs 545 us at 111 x 37 x 37), 1024 beyond (C5: 20.6 vs 22.1 ms at 399 x 133 x 133)
  const int threads = g_coeffs_threads > 0 ? g_coeffs_threads : (n_act <= 64 ? 512 : SM_THREADS);
  // G, L, Li, T, xu, tmp, u_p, buf in shared memory when they fit (and the block is wide enough for
  // the in-shared-memory factorisation: threads >= the power of two above n_act)
  const size_t need = (4 * (size_t)n_act * n_act + (size_t)n_max * n_act + 2 * (size_t)len_u * n_act + 2 * (size_t)n_act) * sizeof(double);
  int rs = 32;
  while (rs < n_act) rs <<= 1;
  int in_smem = (need <= 200 * 1024 && threads >= rs && g_coeffs_smem) ? 1 : 0;
  const size_t need_x = need + (size_t)len_u * n_max * sizeof(double);   // + the eigenvectors u_x
  if (in_smem && need_x <= 200 * 1024) in_smem = 2;
  get_coeffs_kernel<<<1, threads, in_smem == 2 ? need_x : (in_smem ? need : 0), st>>>(len_a, len_u, n_max, n_act, a_red, u_p, work,
                                                                                     in_smem, status_dev);
  ++g_launches;
  DLB_CUDA_CHECK(cudaGetLastError());
}


// ---- reduced problem of caslr_eff_driver (diaglib.f90:1293-1324) ---------------------------
namespace {
// c(i,j) = sum_k a(k,i) a(k,j), i,j < k_  (dgemm('t','n') 1303); one thread per element
__global__ void small_ata_kernel(int k_, const double* __restrict__ a, int lda, double* __restrict__ c, int ldc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= k_ || j >= k_) return;
  double s = 0.0;
  for (int k = 0; k < k_; ++k) s = fma(a[k + (size_t)i * lda], a[k + (size_t)j * lda], s);
  c[i + (size_t)j * ldc] = s;
}
// eig(i) = sqrt(e(k-1-i)), up(:,i) = z(:,k-1-i)  (1314-1317)
__global__ void lr_pick_kernel(int k_, int n_max, const double* __restrict__ z, int ldz, const double* __restrict__ e,
                               double* __restrict__ up, int ldup, double* __restrict__ eig) {
  const int i = blockIdx.y, r = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_max) return;
  if (r == 0) eig[i] = sqrt(e[k_ - 1 - i]);
  if (r < k_) up[r + (size_t)i * ldup] = z[r + (size_t)(k_ - 1 - i) * ldz];
}
// um(:,i) = (s_red * up(:,i)) / eig(i)  (1321-1324)
__global__ void lr_um_kernel(int k_, int n_max, const double* __restrict__ sred, int lds, const double* __restrict__ up,
                             int ldup, const double* __restrict__ e, double* __restrict__ um, int ldum) {
  const int i = blockIdx.y, r = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_max || r >= k_) return;
  double s = 0.0;
  for (int k = 0; k < k_; ++k) s = fma(sred[r + (size_t)k * lds], up[k + (size_t)i * ldup], s);
  um[r + (size_t)i * ldum] = s / sqrt(e[k_ - 1 - i]);
}
}  // namespace

namespace {
// cp ((m + k) x k, ldc) = [-xu (m x k, ldx); I_k], or with T (k x k upper triangular, ld k) [-xu T; T]: the
// coefficients of u <- u T - x (xu T), the projection step applied to a block whose last triangular multiply
// was deferred (xu was taken with the block before that multiply)
__global__ void proj_coeff_kernel(int m, int k, const double* __restrict__ xu, int ldx, const double* __restrict__ T,
                                  double* __restrict__ cp, int ldc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= m + k || j >= k) return;
  double v;
  if (!T) {
    v = i < m ? -xu[i + (size_t)j * ldx] : (i - m == j ? 1.0 : 0.0);
  } else if (i >= m) {
    v = (i - m <= j) ? T[(i - m) + (size_t)j * k] : 0.0;
  } else {
    double s0 = 0.0, s1 = 0.0;
    int l = 0;
    for (; l + 1 <= j; l += 2) {
      s0 = fma(xu[i + (size_t)l * ldx], T[l + (size_t)j * k], s0);
      s1 = fma(xu[i + (size_t)(l + 1) * ldx], T[(l + 1) + (size_t)j * k], s1);
    }
    if (l <= j) s0 = fma(xu[i + (size_t)l * ldx], T[l + (size_t)j * k], s0);
    v = -(s0 + s1);
  }
  cp[i + (size_t)j * ldc] = v;
}
}  // namespace
void proj_coeff(cudaStream_t st, int m, int k, const double* xu, int ldx, double* cp, int ldc, const double* T) {
  if (k <= 0) return;
  proj_coeff_kernel<<<dim3((m + k + 127) / 128, k), 128, 0, st>>>(m, k, xu, ldx, T, cp, ldc);
  ++g_launches;
}

void small_ata(cudaStream_t st, int k, const double* a, int lda, double* c, int ldc) {
  if (k <= 0) return;
  small_ata_kernel<<<dim3((k + 127) / 128, k), 128, 0, st>>>(k, a, lda, c, ldc);
  ++g_launches;
}
void lr_reduced_vectors(cudaStream_t st, int k, int n_max, const double* z, int ldz, const double* e, const double* sred,
                        int lds, double* up, int ldup, double* um, int ldum, double* eig) {
  if (k <= 0) return;
  lr_pick_kernel<<<dim3((k + 127) / 128, n_max), 128, 0, st>>>(k, n_max, z, ldz, e, up, ldup, eig);
  lr_um_kernel<<<dim3((k + 127) / 128, n_max), 128, 0, st>>>(k, n_max, sred, lds, up, ldup, e, um, ldum);
  g_launches += 2;
}

}  // namespace dlb
