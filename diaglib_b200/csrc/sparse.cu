// HBM-bound streaming kernels of the diaglib hot path: CSR block matvec (SpMM), diagonal
// shift-and-invert preconditioner, fused residual + norms, block axpy/copy, halo packing.
#include "common.cuh"
#include "kernels.h"

#include <algorithm>

namespace dlb {
int g_spmm_short = 1;
int g_spmm_chunk = 24;
int g_spmm_chunk_tiled = 0;   // DIAGLIB_B200_SPMM_CHUNK_TILED=1: column chunks also with a caller-given row order
namespace {

// ---------------------------------------------------------------------------------------
// SpMM, one thread per row, JB columns of the block accumulated in registers per sweep over
// the row.  Consecutive threads own consecutive rows, so for banded/stencil matrices every
// gather x[col + j*ld] is a shifted coalesced access and each row sum is formed in CSR
// order with FMAs (the same order as the CPU oracle: results are bit-identical).
// Columns >= n address the halo block (rows owned by neighbouring ranks).
// ---------------------------------------------------------------------------------------
// Row sums for columns [j_begin, m) of one row, CSR order, JB columns per sweep over the row.
template <int JB>
__device__ __forceinline__ void spmm_row_generic(int64_t row, int64_t b, int64_t e, int j_begin, int64_t n, int64_t n_halo,
                                                 const int32_t* __restrict__ col, const double* __restrict__ val, int m,
                                                 const double* __restrict__ x, int64_t ldx, const double* __restrict__ xh,
                                                 double* __restrict__ ax, int64_t ldax, double shift) {
  for (int j0 = j_begin; j0 < m; j0 += JB) {
    double acc[JB];
#pragma unroll
    for (int jj = 0; jj < JB; ++jj) acc[jj] = 0.0;
    if (j0 + JB <= m) {
      for (int64_t k = b; k < e; ++k) {
        const int64_t c = col[k];
        const double v = val[k];
        const double* xp;
        int64_t ld;
        if (c < n) { xp = x + c + (int64_t)j0 * ldx; ld = ldx; }
        else { xp = xh + (c - n) + (int64_t)j0 * n_halo; ld = n_halo; }
#pragma unroll
        for (int jj = 0; jj < JB; ++jj) acc[jj] = fma(v, xp[(int64_t)jj * ld], acc[jj]);
      }
#pragma unroll
      for (int jj = 0; jj < JB; ++jj) {
        double s = acc[jj];
        if (shift != 0.0) s = fma(shift, x[row + (int64_t)(j0 + jj) * ldx], s);
        ax[row + (int64_t)(j0 + jj) * ldax] = s;
      }
    } else {
      const int jn = m - j0;
      for (int64_t k = b; k < e; ++k) {
        const int64_t c = col[k];
        const double v = val[k];
        const double* xp;
        int64_t ld;
        if (c < n) { xp = x + c + (int64_t)j0 * ldx; ld = ldx; }
        else { xp = xh + (c - n) + (int64_t)j0 * n_halo; ld = n_halo; }
#pragma unroll
        for (int jj = 0; jj < JB; ++jj)
          if (jj < jn) acc[jj] = fma(v, xp[(int64_t)jj * ld], acc[jj]);
      }
#pragma unroll
      for (int jj = 0; jj < JB; ++jj)
        if (jj < jn) {
          double s = acc[jj];
          if (shift != 0.0) s = fma(shift, x[row + (int64_t)(j0 + jj) * ldx], s);
          ax[row + (int64_t)(j0 + jj) * ldax] = s;
        }
    }
  }
}

// `order` (optional) lists the rows in processing order; the launch covers entries
// [first, first + count) of it (natural order when null).
template <int JB>
__global__ void __launch_bounds__(256)
spmm_csr_kernel(int64_t n, int64_t n_halo, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                const double* __restrict__ val, int m, const double* __restrict__ x, int64_t ldx,
                const double* __restrict__ xh, double* __restrict__ ax, int64_t ldax, double shift,
                const int32_t* __restrict__ order, int64_t first, int64_t count) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= count) return;
  const int64_t row = order ? (int64_t)order[first + gid] : first + gid;
  spmm_row_generic<JB>(row, rowptr[row], rowptr[row + 1], 0, n, n_halo, col, val, m, x, ldx, xh, ax, ldax, shift);
}

// Short-row variant (every row has at most KMAX entries: stencils).  The column indices of
// the row are read once into registers, so the gathers of all KMAX entries are independent
// of any other load and the scheduler keeps many of them in flight; the generic kernel has
// a dependent col -> x chain per entry and is latency-bound (ncu: no memory level above 60 %).
// Same CSR-order FMA chain per row sum -> bit-identical results.  Warps that touch halo
// columns (c >= n) take the generic path.
// unconditional read-only loads: written as asm so that the compiler cannot sink them under
// the row-length predicate (it turns `on ? fma(v, x[..], acc) : acc` into a branch per load)
__device__ __forceinline__ double ld_nc_f64(const double* p) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int32_t ld_nc_s32(const int32_t* p) {
  int32_t v;
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// The m mod JB remainder goes through the generic row loop.  Three ways to give it the register-resident
// scheme were measured and removed again (profiles/spmm_variants_r02.json, n = 2^24, m = 37, tiled
// order): a second inlined block 4.52 ms, the same at 3 CTAs per SM 2.69 ms, a not-inlined tail
// function 2.85 ms, against 2.57 ms with the generic tail - the extra code changes ptxas' schedule
// of the main loop (fewer gathers in flight under the 64-register cap).  Two uniform-loop forms fared
// no better (profiles/spmm_tail_variants_r02.json): a last pass shifted back to columns [m - 8, m)
// 2.84 ms, a last pass with the column index clamped to m - 1 (surplus gathers are L1 hits) 3.53 ms -
// and both already lose at m = 32, where no remainder exists: any change to the pass loop costs
// loads in flight.
template <int JB, int KMAX>
__global__ void __launch_bounds__(256, 4)
spmm_csr_short_kernel(int64_t n, int64_t n_halo, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                      const double* __restrict__ val, int m, const double* __restrict__ x, int64_t ldx,
                      const double* __restrict__ xh, double* __restrict__ ax, int64_t ldax, double shift,
                      const int32_t* __restrict__ order, int64_t first, int64_t count) {
  const int64_t gid0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = gid0 < count;
  const int64_t gid = valid ? gid0 : count - 1;
  const int64_t row = order ? (int64_t)order[first + gid] : first + gid;
  const int64_t b = rowptr[row];
  const int len = (int)(rowptr[row + 1] - b);
  // entries k >= len repeat the last real entry with a zero coefficient:
  // fma(0, x, acc) == acc exactly for finite x, so the row sum is the CSR-order FMA chain
  int32_t c[KMAX];
  bool local = true;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int kk = len > 0 ? (k < len ? k : len - 1) : 0;
    c[k] = len > 0 ? ld_nc_s32(col + b + kk) : (int32_t)row;
    local = local && c[k] < n;
  }
  if (!__all_sync(0xffffffffu, local)) {   // a halo column in this warp: generic path
    if (valid) spmm_row_generic<JB>(row, b, b + len, 0, n, n_halo, col, val, m, x, ldx, xh, ax, ldax, shift);
    return;
  }
  // kept inline on purpose: wrapped in a device function the same loop schedules its loads
  // much later (measured 3.08 vs 2.27 ms at m = 32, n = 2^24)
  int j0 = 0;
  for (; j0 + JB <= m; j0 += JB) {
    double acc[JB];
#pragma unroll
    for (int jj = 0; jj < JB; ++jj) acc[jj] = 0.0;
    const double* xb = x + (int64_t)j0 * ldx;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int kk = len > 0 ? (k < len ? k : len - 1) : 0;
      double v = len > 0 ? ld_nc_f64(val + b + kk) : 0.0;
      v = k < len ? v : 0.0;
      const double* xp = xb + c[k];
#pragma unroll
      for (int jj = 0; jj < JB; ++jj) acc[jj] = fma(v, ld_nc_f64(xp + (int64_t)jj * ldx), acc[jj]);
    }
    if (valid) {
#pragma unroll
      for (int jj = 0; jj < JB; ++jj) {
        double s = acc[jj];
        if (shift != 0.0) s = fma(shift, xb[row + (int64_t)jj * ldx], s);
        ax[row + (int64_t)(j0 + jj) * ldax] = s;
      }
    }
  }
  if (valid && j0 < m)
    spmm_row_generic<JB>(row, b, b + len, j0, n, n_halo, col, val, m, x, ldx, xh, ax, ldax, shift);
}

__global__ void __launch_bounds__(256)
diag_precnd_kernel(int64_t n, int m, double fac, const double* __restrict__ diag, const double* __restrict__ x,
                   int64_t ldx, double* __restrict__ px, int64_t ldpx) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  const double d = diag[row] + fac;
  const bool use = fabs(d) > 1.0e-5;
  for (int j = 0; j < m; ++j) {
    const double v = x[row + (int64_t)j * ldx];
    px[row + (int64_t)j * ldpx] = use ? v / d : v;
  }
}

// one CTA per (row slice, column): grid = (slices, m)
constexpr int RN_THREADS = 256;
__global__ void __launch_bounds__(RN_THREADS)
residual_kernel(int64_t n, int m, const double* ax, int64_t ldax, const double* __restrict__ x, int64_t ldx,
                const double* __restrict__ theta, const int* __restrict__ active, double* r, int64_t ldr,
                double* __restrict__ scratch) {
  const int j = blockIdx.y;
  const int nsl = gridDim.x;
  const int64_t per = (n + nsl - 1) / nsl;
  const int64_t r0 = (int64_t)blockIdx.x * per, r1 = r0 + per < n ? r0 + per : n;
  const bool act = active[j] != 0;
  const double th = theta[j];
  const double* axj = ax + (int64_t)j * ldax;
  const double* xj = x + (int64_t)j * ldx;
  double* rj = r + (int64_t)j * ldr;
  double ss = 0.0, mx = 0.0;
  if (act) {
    for (int64_t i = r0 + threadIdx.x; i < r1; i += RN_THREADS) {
      const double v = fma(-th, xj[i], axj[i]);
      rj[i] = v;
      ss = fma(v, v, ss);
      mx = fmax(mx, fabs(v));
    }
  } else if (axj != rj) {
    for (int64_t i = r0 + threadIdx.x; i < r1; i += RN_THREADS) rj[i] = axj[i];
  }
  __shared__ double s_ss[RN_THREADS / 32], s_mx[RN_THREADS / 32];
  ss = warp_sum(ss);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) { s_ss[threadIdx.x >> 5] = ss; s_mx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < RN_THREADS / 32; ++w) { a += s_ss[w]; b = fmax(b, s_mx[w]); }
    scratch[(size_t)blockIdx.x * 2 * m + j] = a;
    scratch[(size_t)blockIdx.x * 2 * m + m + j] = b;
  }
}
__global__ void residual_final_kernel(int nsl, int m, const double* __restrict__ scratch, double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  double a = 0.0, b = 0.0;
  for (int s = 0; s < nsl; ++s) {
    a += scratch[(size_t)s * 2 * m + j];
    b = fmax(b, scratch[(size_t)s * 2 * m + m + j]);
  }
  out[j] = a;
  out[m + j] = b;
}

__global__ void __launch_bounds__(256)
axpy_kernel(int64_t n, int m, double a, const double* __restrict__ x, int64_t ldx, double* __restrict__ y,
            int64_t ldy) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  for (int j = 0; j < m; ++j) y[row + (int64_t)j * ldy] = fma(a, x[row + (int64_t)j * ldx], y[row + (int64_t)j * ldy]);
}

// linear-response helpers (caslr_eff_driver): split evec = [Y; Z] into Y+Z / Y-Z, merge back,
// and the diagonal preconditioner of main.f90:257-281
__global__ void __launch_bounds__(256)
lr_split_kernel(int64_t n, int m, const double* __restrict__ evec, int64_t ld2, double* __restrict__ vp,
                double* __restrict__ vm, int64_t ldv) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  for (int j = 0; j < m; ++j) {
    const double y = evec[row + (int64_t)j * ld2], z = evec[n + row + (int64_t)j * ld2];
    vp[row + (int64_t)j * ldv] = y + z;
    vm[row + (int64_t)j * ldv] = y - z;
  }
}
__global__ void __launch_bounds__(256)
lr_merge_kernel(int64_t n, int m, const double* __restrict__ ep, const double* __restrict__ em, int64_t lde,
                double* __restrict__ evec, int64_t ld2) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  for (int j = 0; j < m; ++j) {
    const double a = ep[row + (int64_t)j * lde], b = em[row + (int64_t)j * lde];
    evec[row + (int64_t)j * ld2] = a + b;
    evec[n + row + (int64_t)j * ld2] = a - b;
  }
}
__global__ void __launch_bounds__(256)
lr_precnd_kernel(int64_t n, int m, double fac, const double* __restrict__ aa, const double* __restrict__ sg,
                 const double* __restrict__ xp, const double* __restrict__ xm, double* __restrict__ yp,
                 double* __restrict__ ym) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  const double a = aa[row], s = sg[row];
  double denom = fac * fac * a * a - s * s;
  denom = 1.0 / denom;
  for (int j = 0; j < m; ++j) {
    const int64_t o = row + (int64_t)j * n;
    const double p = xp[o], q = xm[o];
    yp[o] = denom * (fac * a * p + s * q);
    ym[o] = denom * (fac * a * q + s * p);
  }
}

}  // namespace

void lr_split(cudaStream_t st, int64_t n, int m, const double* evec, int64_t ld2, double* vp, double* vm, int64_t ldv) {
  if (n <= 0 || m <= 0) return;
  lr_split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, m, evec, ld2, vp, vm, ldv);
  ++g_launches;
  DLB_CUDA_CHECK(cudaGetLastError());
}
void lr_merge(cudaStream_t st, int64_t n, int m, const double* ep, const double* em, int64_t lde, double* evec, int64_t ld2) {
  if (n <= 0 || m <= 0) return;
  lr_merge_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, m, ep, em, lde, evec, ld2);
  ++g_launches;
  DLB_CUDA_CHECK(cudaGetLastError());
}
void lr_precnd(cudaStream_t st, int64_t n, int m, double fac, const double* aa, const double* sg, const double* xp,
               const double* xm, double* yp, double* ym) {
  if (n <= 0 || m <= 0) return;
  lr_precnd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, m, fac, aa, sg, xp, xm, yp, ym);
  ++g_launches;
  DLB_CUDA_CHECK(cudaGetLastError());
}

void spmm_csr(cudaStream_t st, const CsrDevice& A, int m, const double* x, int64_t ldx, const double* x_halo,
              double* ax, int64_t ldax, double shift, int part) {
  if (A.n <= 0 || m <= 0) return;
  // rows of this launch: all of them, or the interior / boundary section of A.order
  int64_t first = 0, count = A.n;
  if (part == SPMM_INTERIOR) count = A.n_interior;
  else if (part == SPMM_BOUNDARY) { first = A.n_interior; count = A.n - A.n_interior; }
  if (count <= 0) return;
  const unsigned grid = (unsigned)((count + 255) / 256);
  if (A.max_row_nnz > 0 && A.max_row_nnz <= 7 && g_spmm_short > 0) {
    // column chunks, one launch each: with the gathers pipelined the kernel is DRAM-bound on its
    // ACTUAL traffic, and beyond ~24 columns the far-neighbour reuse window (2 planes x m columns,
    // read + written) falls out of L2 and x is fetched from DRAM more than once (ncu at m = 37:
    // 15.9 GB moved for 11.5 GB algorithmic).  A chunk re-reads the matrix but keeps x in L2.
    // With a locality-preserving row order (A.tiled) the window is a tile neighbourhood and one
    // launch sweeps all columns.
    int jc = m;
    if (g_spmm_chunk > 0 && m > g_spmm_chunk && (!A.tiled || g_spmm_chunk_tiled)) {
      const int npass = (m + g_spmm_chunk - 1) / g_spmm_chunk;
      jc = (((m + npass - 1) / npass) + 7) / 8 * 8;
    }
    for (int j0 = 0; j0 < m; j0 += jc) {
      const int mc = std::min(jc, m - j0);
      // halo block columns are n_halo apart
      const double* xc = x + (int64_t)j0 * ldx;
      const double* hc = x_halo ? x_halo + (int64_t)j0 * A.n_halo : nullptr;
      double* axc = ax + (int64_t)j0 * ldax;
      spmm_csr_short_kernel<8, 7><<<grid, 256, 0, st>>>(A.n, A.n_halo, A.rowptr, A.col, A.val, mc, xc, ldx, hc, axc, ldax, shift,
                                                        A.order, first, count);
      ++g_launches;
    }
  } else {
    spmm_csr_kernel<8><<<grid, 256, 0, st>>>(A.n, A.n_halo, A.rowptr, A.col, A.val, m, x, ldx, x_halo, ax, ldax, shift,
                                             A.order, first, count);
    ++g_launches;
  }
  DLB_CUDA_CHECK(cudaGetLastError());
}

void diag_precnd(cudaStream_t st, int64_t n, int m, double fac, const double* diag, const double* x, int64_t ldx,
                 double* px, int64_t ldpx) {
  if (n <= 0 || m <= 0) return;
  diag_precnd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, m, fac, diag, x, ldx, px, ldpx);
  ++g_launches;
  DLB_CUDA_CHECK(cudaGetLastError());
}

static int residual_slices(int m, int num_sms) { return std::max(1, (num_sms * 8 + m - 1) / m); }
size_t residual_scratch_bytes(int m, int num_sms) {
  return (size_t)residual_slices(std::max(1, m), num_sms) * 2 * std::max(1, m) * sizeof(double);
}
void residual_norms(cudaStream_t st, int num_sms, int64_t n, int m, const double* ax, int64_t ldax, const double* x,
                    int64_t ldx, const double* theta, const int* active, double* r, int64_t ldr,
                    double* norms_out, double* scratch) {
  if (m <= 0) return;
  int nsl = residual_slices(m, num_sms);
  nsl = (int)std::max<int64_t>(1, std::min<int64_t>(nsl, (n + RN_THREADS - 1) / RN_THREADS));
  residual_kernel<<<dim3(nsl, m), RN_THREADS, 0, st>>>(n, m, ax, ldax, x, ldx, theta, active, r, ldr, scratch);
  ++g_launches;
  residual_final_kernel<<<(m + 63) / 64, 64, 0, st>>>(nsl, m, scratch, norms_out);
  ++g_launches;
  DLB_CUDA_CHECK(cudaGetLastError());
}

void block_axpy(cudaStream_t st, int64_t n, int m, double a, const double* x, int64_t ldx, double* y, int64_t ldy) {
  if (n <= 0 || m <= 0) return;
  axpy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, m, a, x, ldx, y, ldy);
  ++g_launches;
  DLB_CUDA_CHECK(cudaGetLastError());
}

void block_copy(cudaStream_t st, int64_t n, int m, const double* x, int64_t ldx, double* y, int64_t ldy) {
  if (n <= 0 || m <= 0 || x == y) return;
  if (ldx == n && ldy == n)
    DLB_CUDA_CHECK(cudaMemcpyAsync(y, x, sizeof(double) * (size_t)n * m, cudaMemcpyDeviceToDevice, st));
  else
    DLB_CUDA_CHECK(cudaMemcpy2DAsync(y, sizeof(double) * ldy, x, sizeof(double) * ldx, sizeof(double) * n, m,
                                     cudaMemcpyDeviceToDevice, st));
}

void pack_rows(cudaStream_t st, int64_t row0, int64_t cnt, int m, const double* x, int64_t ldx, double* out) {
  if (cnt <= 0 || m <= 0) return;
  DLB_CUDA_CHECK(cudaMemcpy2DAsync(out, sizeof(double) * cnt, x + row0, sizeof(double) * ldx, sizeof(double) * cnt, m,
                                   cudaMemcpyDeviceToDevice, st));
}

}  // namespace dlb
