// Tall-skinny FP64 GEMMs of the diaglib hot path on the sm_100a FP64 tensor pipe (DMMA).
//
//   gram_tn   : C(p x q)  = A(n x p)^T B(n x q)      reduction over the long dimension n
//   block_mul : Y(n x q)  = alpha V(n x p) C(p x q) + beta Y
//
// Both stream the n-long operands exactly once from HBM through a cp.async (LDGSTS)
// multi-stage shared-memory ring whose column stride is padded (== 4 mod 16 doubles) so
// that every DMMA m8n8k4 fragment load is bank-conflict free.
#include "common.cuh"
#include "kernels.h"

#include <cuda.h>  // CUtensorMap (types only; the encoder is resolved through the runtime)

#include <algorithm>
#include <map>
#include <vector>

namespace dlb {
const int* g_live = nullptr;   // see kernels.h
PeerWin g_peerwin;
bool g_fuse_allreduce = false;

bool g_disable_ws = false;
int g_dbg = 0;       // DIAGLIB_B200_DBG (race hunting): 1 / 2 = an empty kernel before / after each ws block multiply, 4 = proxy fence before the first bulk copy, 8 = __threadfence after the last store
int g_ws_mask = 0;   // DIAGLIB_B200_WS_MASK (A/B testing): 1 = no warp-specialised Gram kernels, 2 = none for the block multiply, 4 = no bulk-copy Gram
bool g_disable_tma = false;
bool g_disable_fused_gram = false;  // DIAGLIB_B200_NO_FUSED_GRAM=1
bool g_bmul_small_tiles = false;   // DIAGLIB_B200_BMUL_RT256=0 selects 128-row tiles with two CTAs per SM (see launch_blockmul: not reproducible)

// =====================================================================================
// gram_tn
// =====================================================================================
namespace {

constexpr int GR_THREADS = 512;
constexpr int GR_WARPS = GR_THREADS / 32;
constexpr int GR_STAGES = 3;
constexpr int GR_MAXB = 128;       // max p / q handled by one launch
// Rows of n per pipeline stage (KT) are chosen per launch: narrow blocks get long stages so
// that enough bytes are in flight per SM; the shared-memory column stride is KT + 4 doubles
// (== 4 mod 16), which makes every DMMA fragment load bank-conflict free.

// A warp task is (part of) a 2 x 4 frame of 8x8 output tiles; bit r*4+c of tmask says whether
// tile (ti0+r, tj0+c) is computed.  The host picks the task shape (2x4, 2x2, 1x2 or 1x1 tiles)
// so that all 16 warps have work even for narrow blocks, and balances the tasks over the
// four SM sub-partitions (warp id mod 4), which matters for the triangular (sym_lower) case.
struct GramTask {
  uint8_t ti0, tj0;
  uint8_t nc0, nc1;  // number of tiles computed in frame row 0 / row 1 (a prefix tj0 .. tj0+nc-1)
};

// All k-steps of one stage for one task, specialised on the tile counts of the two frame rows
// so that the inner loop is branch free (the counts are warp-uniform but only known at run time;
// the dispatch happens once per stage, outside the k loop).
template <int NC0, int NC1>
__device__ __forceinline__ void gram_task(double (&acc)[2][4][2], const double* pa, const double* pb, int KT, int S) {
  constexpr int NB = NC0 > NC1 ? NC0 : NC1;
  for (int k8 = 0; k8 < KT; k8 += 16) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int k = k8 + ks * 4;
      double a0 = 0.0, a1 = 0.0, b[NB > 0 ? NB : 1];
      if (NC0 > 0) a0 = pa[k];
      if (NC1 > 0) a1 = pa[8 * S + k];
#pragma unroll
      for (int c = 0; c < NB; ++c) b[c] = pb[c * 8 * S + k];
#pragma unroll
      for (int c = 0; c < NC0; ++c) dmma884(acc[0][c][0], acc[0][c][1], a0, b[c]);
#pragma unroll
      for (int c = 0; c < NC1; ++c) dmma884(acc[1][c][0], acc[1][c][1], a1, b[c]);
    }
  }
}

__device__ __forceinline__ void gram_task_dispatch(int nc0, int nc1, double (&acc)[2][4][2], const double* pa,
                                                   const double* pb, int KT, int S) {
#define DLB_CASE(A, B) case (A) * 5 + (B): gram_task<A, B>(acc, pa, pb, KT, S); break;
  switch (nc0 * 5 + nc1) {
    DLB_CASE(0, 1) DLB_CASE(0, 2) DLB_CASE(0, 3) DLB_CASE(0, 4)
    DLB_CASE(1, 0) DLB_CASE(1, 1) DLB_CASE(1, 2) DLB_CASE(1, 3) DLB_CASE(1, 4)
    DLB_CASE(2, 0) DLB_CASE(2, 1) DLB_CASE(2, 2) DLB_CASE(2, 3) DLB_CASE(2, 4)
    DLB_CASE(3, 0) DLB_CASE(3, 1) DLB_CASE(3, 2) DLB_CASE(3, 3) DLB_CASE(3, 4)
    DLB_CASE(4, 0) DLB_CASE(4, 1) DLB_CASE(4, 2) DLB_CASE(4, 3) DLB_CASE(4, 4)
    default: break;
  }
#undef DLB_CASE
}
struct GramSched {
  GramTask t[GR_WARPS][2];
};

template <bool ALIGN16, int NT = GR_THREADS>
__device__ __forceinline__ void gram_load_stage(double* s, const double* __restrict__ M, int64_t ld, int ncols,
                                                int64_t k0, int64_t n, int tid, int KT) {
  const int S = KT + 4;
  if (ALIGN16) {
    const int half = KT >> 1;
    const int total = ncols * half;
    for (int id = tid; id < total; id += NT) {
      const int col = id / half, part = id - col * half;
      const int64_t row = k0 + part * 2;
      int64_t rem = (n - row) * 8;
      const int bytes = rem >= 16 ? 16 : (rem > 0 ? (int)rem : 0);
      const double* src = M + (int64_t)col * ld + (bytes > 0 ? row : 0);
      cp_async16(s + col * S + part * 2, src, bytes);
    }
  } else {
    const int total = ncols * KT;
    for (int id = tid; id < total; id += NT) {
      const int col = id / KT, part = id - col * KT;
      const int64_t row = k0 + part;
      const int bytes = row < n ? 8 : 0;
      const double* src = M + (int64_t)col * ld + (bytes > 0 ? row : 0);
      cp_async8(s + col * S + part, src, bytes);
    }
  }
}

template <bool ALIGN16>
__global__ void __launch_bounds__(GR_THREADS, 1)
gram_kernel(int64_t n, const double* __restrict__ A, int64_t lda, int p, const double* __restrict__ B, int64_t ldb,
            int q, int same, int KT, const __grid_constant__ GramSched sched, double* __restrict__ partial, int PB,
            int QB, const int* __restrict__ live) {
  if (live && *live == 0) return;   // predicated step of a speculative ortho chain (engine.cu)
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = KT + 4;
  const int stage_doubles = (PB + (same ? 0 : QB)) * S;

  const GramTask t0 = sched.t[warp][0], t1 = sched.t[warp][1];
  double acc[2][2][4][2];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[s][r][c][0] = acc[s][r][c][1] = 0.0;

  const int64_t nchunks = (n + KT - 1) / KT;
  const int64_t first = blockIdx.x, stride = gridDim.x;
  const int64_t my_chunks = first < nchunks ? (nchunks - first + stride - 1) / stride : 0;

  auto issue = [&](int64_t local_idx) {
    if (local_idx < my_chunks) {
      double* s = smem + (local_idx % GR_STAGES) * stage_doubles;
      const int64_t k0 = (first + local_idx * stride) * KT;
      gram_load_stage<ALIGN16>(s, A, lda, p, k0, n, tid, KT);
      if (!same) gram_load_stage<ALIGN16>(s + PB * S, B, ldb, q, k0, n, tid, KT);
    }
    cp_async_commit();
  };

#pragma unroll
  for (int s = 0; s < GR_STAGES - 1; ++s) issue(s);

  const int frag_off = (lane >> 2) * S + (lane & 3);
  for (int64_t it = 0; it < my_chunks; ++it) {
    cp_async_wait<GR_STAGES - 2>();
    __syncthreads();
    issue(it + GR_STAGES - 1);
    const double* sA = smem + (it % GR_STAGES) * stage_doubles;
    const double* sB = same ? sA : sA + PB * S;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const GramTask t = s == 0 ? t0 : t1;
      if ((t.nc0 | t.nc1) == 0) continue;
      gram_task_dispatch(t.nc0, t.nc1, acc[s], sA + (t.ti0 * 8) * S + frag_off, sB + (t.tj0 * 8) * S + frag_off, KT, S);
    }
  }
  cp_async_wait<0>();

  // per-CTA partial result, PB x QB column-major
  double* out = partial + (size_t)blockIdx.x * PB * QB;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const GramTask t = s == 0 ? t0 : t1;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < (r == 0 ? t.nc0 : t.nc1)) {
          const int i = (t.ti0 + r) * 8 + (lane >> 2);
          const int j = (t.tj0 + c) * 8 + (lane & 3) * 2;
          out[i + (size_t)j * PB] = acc[s][r][c][0];
          out[i + (size_t)(j + 1) * PB] = acc[s][r][c][1];
        }
  }
}

// Warp-specialised variant of gram_kernel: a 17th warp is the producer and feeds the ring with
// 1-D bulk async copies (TMA engine; one KT*8-byte column segment per copy, completion counted
// on full[stage]); the 16 consumer warps never meet at a CTA-wide barrier.
constexpr int GRW_THREADS = GR_THREADS;   // 15 consumer warps + 1 producer warp
constexpr int GRW_CONS = GR_WARPS - 1;

__global__ void __launch_bounds__(GRW_THREADS, 1)
gram_ws_kernel(int64_t n, const double* __restrict__ A, int64_t lda, int p, const double* __restrict__ B, int64_t ldb,
               int q, int same, int KT, const __grid_constant__ GramSched sched, double* __restrict__ partial, int PB,
               int QB, const int* __restrict__ live) {
  if (live && *live == 0) return;   // predicated step of a speculative ortho chain (engine.cu)
  extern __shared__ __align__(16) double smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + GR_STAGES;
  double* ring = smem + 2 * GR_STAGES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = KT + 4;
  const int stage_doubles = (PB + (same ? 0 : QB)) * S;
  const int64_t nchunks = (n + KT - 1) / KT;
  const int64_t first = blockIdx.x, stride = gridDim.x;
  const int64_t my_chunks = first < nchunks ? (nchunks - first + stride - 1) / stride : 0;

  if (tid == 0) {
    for (int s = 0; s < GR_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], GRW_CONS); }
    mbar_fence_init();
  }
  // columns p..PB-1 (q..QB-1) are never loaded: keep them finite
  for (int id = tid; id < GR_STAGES * stage_doubles; id += GRW_THREADS) ring[id] = 0.0;
  fence_proxy_async();
  __syncthreads();

  if (warp == GRW_CONS) {
    // ---------------- producer ----------------
    int s = 0;
    uint32_t ph = 0;
    const int ncopy = p + (same ? 0 : q);
    for (int64_t it = 0; it < my_chunks; ++it) {
      const int64_t k0 = (first + it * stride) * KT;
      const int64_t rem = n - k0;
      const uint32_t rows = rem < KT ? (uint32_t)rem : (uint32_t)KT;
      double* st = ring + (size_t)s * stage_doubles;
      mbar_wait(&empty[s], ph ^ 1);
      if (rows < (uint32_t)KT) {
        // last chunk of the block: rows beyond n must contribute zero
        for (int c = lane; c < ncopy; c += 32) {
          double* d = st + (c < p ? c : PB + (c - p)) * S;
          for (int r = rows; r < KT; ++r) d[r] = 0.0;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)ncopy * rows * 8u);
      __syncwarp();
      for (int c = lane; c < ncopy; c += 32) {
        if (c < p) bulk_g2s(st + c * S, A + (int64_t)c * lda + k0, rows * 8u, &full[s]);
        else bulk_g2s(st + (PB + (c - p)) * S, B + (int64_t)(c - p) * ldb + k0, rows * 8u, &full[s]);
      }
      if (++s == GR_STAGES) { s = 0; ph ^= 1; }
    }
    return;
  }

  // ---------------- consumers ----------------
  const GramTask t0 = sched.t[warp][0], t1 = sched.t[warp][1];
  double acc[2][2][4][2];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[s][r][c][0] = acc[s][r][c][1] = 0.0;
  const int frag_off = (lane >> 2) * S + (lane & 3);
  int s = 0;
  uint32_t ph = 0;
  for (int64_t it = 0; it < my_chunks; ++it) {
    mbar_wait(&full[s], ph);
    const double* sA = ring + (size_t)s * stage_doubles;
    const double* sB = same ? sA : sA + PB * S;
    if ((t0.nc0 | t0.nc1) != 0)
      gram_task_dispatch(t0.nc0, t0.nc1, acc[0], sA + (t0.ti0 * 8) * S + frag_off, sB + (t0.tj0 * 8) * S + frag_off, KT, S);
    if ((t1.nc0 | t1.nc1) != 0)
      gram_task_dispatch(t1.nc0, t1.nc1, acc[1], sA + (t1.ti0 * 8) * S + frag_off, sB + (t1.tj0 * 8) * S + frag_off, KT, S);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    if (++s == GR_STAGES) { s = 0; ph ^= 1; }
  }
  double* out = partial + (size_t)blockIdx.x * PB * QB;
#pragma unroll
  for (int sl = 0; sl < 2; ++sl) {
    const GramTask t = sl == 0 ? t0 : t1;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < (r == 0 ? t.nc0 : t.nc1)) {
          const int i = (t.ti0 + r) * 8 + (lane >> 2);
          const int j = (t.tj0 + c) * 8 + (lane & 3) * 2;
          out[i + (size_t)j * PB] = acc[sl][r][c][0];
          out[i + (size_t)(j + 1) * PB] = acc[sl][r][c][1];
        }
  }
}

// Warp-specialised variant with cp.async producers: two producer warps issue the 16-byte
// LDGSTS copies of a stage and signal full[stage] through cp.async.mbarrier.arrive; fourteen
// consumer warps run the DMMA tasks.  Used for wide blocks, where the column segments of a
// stage are too short (256-512 B) for the bulk-copy engine to reach HBM bandwidth.
constexpr int GRC_PROD = 2;
constexpr int GRC_CONS = GR_WARPS - GRC_PROD;

template <bool ALIGN16>
__global__ void __launch_bounds__(GR_THREADS, 1)
gram_wsc_kernel(int64_t n, const double* __restrict__ A, int64_t lda, int p, const double* __restrict__ B, int64_t ldb,
                int q, int same, int KT, const __grid_constant__ GramSched sched, double* __restrict__ partial, int PB,
                int QB, const int* __restrict__ live) {
  if (live && *live == 0) return;   // predicated step of a speculative ortho chain (engine.cu)
  extern __shared__ __align__(16) double smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + GR_STAGES;
  double* ring = smem + 2 * GR_STAGES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = KT + 4;
  const int stage_doubles = (PB + (same ? 0 : QB)) * S;
  const int64_t nchunks = (n + KT - 1) / KT;
  const int64_t first = blockIdx.x, stride = gridDim.x;
  const int64_t my_chunks = first < nchunks ? (nchunks - first + stride - 1) / stride : 0;

  if (tid == 0) {
    for (int s = 0; s < GR_STAGES; ++s) { mbar_init(&full[s], GRC_PROD * 32); mbar_init(&empty[s], GRC_CONS); }
    mbar_fence_init();
  }
  for (int id = tid; id < GR_STAGES * stage_doubles; id += GR_THREADS) ring[id] = 0.0;
  __syncthreads();

  if (warp >= GRC_CONS) {
    // ---------------- producers ----------------
    const int ptid = tid - GRC_CONS * 32;
    int s = 0;
    uint32_t ph = 0;
    for (int64_t it = 0; it < my_chunks; ++it) {
      const int64_t k0 = (first + it * stride) * KT;
      double* st = ring + (size_t)s * stage_doubles;
      mbar_wait(&empty[s], ph ^ 1);
      gram_load_stage<ALIGN16, GRC_PROD * 32>(st, A, lda, p, k0, n, ptid, KT);
      if (!same) gram_load_stage<ALIGN16, GRC_PROD * 32>(st + PB * S, B, ldb, q, k0, n, ptid, KT);
      cp_async_mbar_arrive_noinc(&full[s]);
      if (++s == GR_STAGES) { s = 0; ph ^= 1; }
    }
    cp_async_wait<0>();
    return;
  }

  // ---------------- consumers ----------------
  const GramTask t0 = sched.t[warp][0], t1 = sched.t[warp][1];
  double acc[2][2][4][2];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[s][r][c][0] = acc[s][r][c][1] = 0.0;
  const int frag_off = (lane >> 2) * S + (lane & 3);
  int s = 0;
  uint32_t ph = 0;
  for (int64_t it = 0; it < my_chunks; ++it) {
    mbar_wait(&full[s], ph);
    const double* sA = ring + (size_t)s * stage_doubles;
    const double* sB = same ? sA : sA + PB * S;
    if ((t0.nc0 | t0.nc1) != 0)
      gram_task_dispatch(t0.nc0, t0.nc1, acc[0], sA + (t0.ti0 * 8) * S + frag_off, sB + (t0.tj0 * 8) * S + frag_off, KT, S);
    if ((t1.nc0 | t1.nc1) != 0)
      gram_task_dispatch(t1.nc0, t1.nc1, acc[1], sA + (t1.ti0 * 8) * S + frag_off, sB + (t1.tj0 * 8) * S + frag_off, KT, S);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    if (++s == GR_STAGES) { s = 0; ph ^= 1; }
  }
  double* out = partial + (size_t)blockIdx.x * PB * QB;
#pragma unroll
  for (int sl = 0; sl < 2; ++sl) {
    const GramTask t = sl == 0 ? t0 : t1;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < (r == 0 ? t.nc0 : t.nc1)) {
          const int i = (t.ti0 + r) * 8 + (lane >> 2);
          const int j = (t.tj0 + c) * 8 + (lane & 3) * 2;
          out[i + (size_t)j * PB] = acc[sl][r][c][0];
          out[i + (size_t)(j + 1) * PB] = acc[sl][r][c][1];
        }
  }
}

// ---------------------------------------------------------------------------------------
// TMA-tiled variant (cp.async.bulk.tensor, SASS UTMALDG).  One elected producer lane issues
// a 2-D box load per 16 rows of n per operand: box = {16 rows (128 B, the swizzle span), all
// columns of the block}, SWIZZLE_128B, out-of-bounds rows/columns zero-filled by the hardware
// (so the tail of n and the padding columns need no code).  Shared-memory image of a box:
// column c occupies the 128-byte line c; its 16-byte chunk j (rows 2j, 2j+1) sits at chunk
// j ^ (c & 7).  DMMA k-step s of a box uses rows {2s, 2s+1, 8+2s, 9+2s}: with that row choice
// the 16 lanes of a half-warp hit 16 distinct 8-byte banks, i.e. fragment loads are conflict
// free without any padding.  (Any row permutation is legal: both operands use the same one.)
// ---------------------------------------------------------------------------------------
constexpr int GT_BOX_ROWS = 16;
constexpr int GT_LINE = 16;  // doubles per 128-byte line

template <int NC0, int NC1>
__device__ __forceinline__ void gram_task_tma(double (&acc)[2][4][2], const double* pa, const double* pb, int nbox,
                                              int boxA, int boxB, const int (&koff)[4]) {
  constexpr int NB = NC0 > NC1 ? NC0 : NC1;
  for (int bx = 0; bx < nbox; ++bx) {
    const double* a_ = pa + bx * boxA;
    const double* b_ = pb + bx * boxB;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      double a0 = 0.0, a1 = 0.0, b[NB > 0 ? NB : 1];
      if (NC0 > 0) a0 = a_[koff[s]];
      if (NC1 > 0) a1 = a_[8 * GT_LINE + koff[s]];
#pragma unroll
      for (int c = 0; c < NB; ++c) b[c] = b_[c * 8 * GT_LINE + koff[s]];
#pragma unroll
      for (int c = 0; c < NC0; ++c) dmma884(acc[0][c][0], acc[0][c][1], a0, b[c]);
#pragma unroll
      for (int c = 0; c < NC1; ++c) dmma884(acc[1][c][0], acc[1][c][1], a1, b[c]);
    }
  }
}
__device__ __forceinline__ void gram_task_tma_dispatch(int nc0, int nc1, double (&acc)[2][4][2], const double* pa,
                                                       const double* pb, int nbox, int boxA, int boxB,
                                                       const int (&koff)[4]) {
#define DLB_CASE(A, B) case (A) * 5 + (B): gram_task_tma<A, B>(acc, pa, pb, nbox, boxA, boxB, koff); break;
  switch (nc0 * 5 + nc1) {
    DLB_CASE(0, 1) DLB_CASE(0, 2) DLB_CASE(0, 3) DLB_CASE(0, 4)
    DLB_CASE(1, 0) DLB_CASE(1, 1) DLB_CASE(1, 2) DLB_CASE(1, 3) DLB_CASE(1, 4)
    DLB_CASE(2, 0) DLB_CASE(2, 1) DLB_CASE(2, 2) DLB_CASE(2, 3) DLB_CASE(2, 4)
    DLB_CASE(3, 0) DLB_CASE(3, 1) DLB_CASE(3, 2) DLB_CASE(3, 3) DLB_CASE(3, 4)
    DLB_CASE(4, 0) DLB_CASE(4, 1) DLB_CASE(4, 2) DLB_CASE(4, 3) DLB_CASE(4, 4)
    default: break;
  }
#undef DLB_CASE
}

__global__ void __launch_bounds__(GR_THREADS, 1)
gram_tma_kernel(int64_t n, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int same,
                int KT, const __grid_constant__ GramSched sched, double* __restrict__ partial, int PB, int QB, const int* __restrict__ live) {
  if (live && *live == 0) return;   // predicated step of a speculative ortho chain (engine.cu)
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // [ring: GR_STAGES x (nbox x (PB + QB) lines of 128 B)] [barriers]
  const int nbox = KT / GT_BOX_ROWS;
  const int boxA = PB * GT_LINE, boxB = QB * GT_LINE;                 // doubles per box
  const int stage_doubles = nbox * (boxA + (same ? 0 : boxB));
  // SWIZZLE_128B needs 1024-byte aligned boxes: align the ring explicitly (1 KB of slack is allocated)
  double* ring = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)GR_STAGES * stage_doubles);
  uint64_t* empty = full + GR_STAGES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t nchunks = (n + KT - 1) / KT;
  const int64_t first = blockIdx.x, stride = gridDim.x;
  const int64_t my_chunks = first < nchunks ? (nchunks - first + stride - 1) / stride : 0;

  if (tid == 0) {
    for (int s = 0; s < GR_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], GRW_CONS); }
    mbar_fence_init();
    tma_prefetch_desc(&tmA);
    if (!same) tma_prefetch_desc(&tmB);
  }
  __syncthreads();

  if (warp == GRW_CONS) {
    // ---------------- producer: one elected lane ----------------
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t bytes = (uint32_t)stage_doubles * 8u;
      for (int64_t it = 0; it < my_chunks; ++it) {
        const int64_t k0 = (first + it * stride) * KT;
        double* st = ring + (size_t)s * stage_doubles;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], bytes);
        for (int bx = 0; bx < nbox; ++bx) {
          tma_load_2d(st + bx * boxA, &tmA, (int)(k0 + bx * GT_BOX_ROWS), 0, &full[s]);
          if (!same) tma_load_2d(st + nbox * boxA + bx * boxB, &tmB, (int)(k0 + bx * GT_BOX_ROWS), 0, &full[s]);
        }
        if (++s == GR_STAGES) { s = 0; ph ^= 1; }
      }
    }
    return;
  }

  // ---------------- consumers ----------------
  const GramTask t0 = sched.t[warp][0], t1 = sched.t[warp][1];
  double acc[2][2][4][2];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[s][r][c][0] = acc[s][r][c][1] = 0.0;
  // per-lane fragment offsets (doubles) inside a box: column line i = lane>>2 of the 8-column
  // tile, rows {2s, 2s+1, 8+2s, 9+2s}[lane&3], 16-byte chunks swizzled by the line index
  const int i8 = lane >> 2, kk = lane & 3;
  int koff[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) koff[s] = i8 * GT_LINE + ((((s + 4 * (kk >> 1)) ^ i8) << 1) | (kk & 1));
  int s = 0;
  uint32_t ph = 0;
  for (int64_t it = 0; it < my_chunks; ++it) {
    mbar_wait(&full[s], ph);
    const double* sA = ring + (size_t)s * stage_doubles;
    const double* sB = same ? sA : sA + nbox * boxA;
    const int bB = same ? boxA : boxB;
    if ((t0.nc0 | t0.nc1) != 0)
      gram_task_tma_dispatch(t0.nc0, t0.nc1, acc[0], sA + t0.ti0 * 8 * GT_LINE, sB + t0.tj0 * 8 * GT_LINE, nbox, boxA, bB, koff);
    if ((t1.nc0 | t1.nc1) != 0)
      gram_task_tma_dispatch(t1.nc0, t1.nc1, acc[1], sA + t1.ti0 * 8 * GT_LINE, sB + t1.tj0 * 8 * GT_LINE, nbox, boxA, bB, koff);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    if (++s == GR_STAGES) { s = 0; ph ^= 1; }
  }
  double* out = partial + (size_t)blockIdx.x * PB * QB;
#pragma unroll
  for (int sl = 0; sl < 2; ++sl) {
    const GramTask t = sl == 0 ? t0 : t1;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < (r == 0 ? t.nc0 : t.nc1)) {
          const int i = (t.ti0 + r) * 8 + (lane >> 2);
          const int j = (t.tj0 + c) * 8 + (lane & 3) * 2;
          out[i + (size_t)j * PB] = acc[sl][r][c][0];
          out[i + (size_t)(j + 1) * PB] = acc[sl][r][c][1];
        }
  }
}

// host side: tensor map for an n x ncols column-major block (ld), box = {16 rows, box_cols}
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encoder() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
    else
      cudaGetLastError();
  }
  return fn;
}
bool make_tmap(CUtensorMap* tm, const double* base, int64_t n, int ncols, int64_t ld, int box_cols) {
  PFN_encodeTiled enc = get_encoder();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)ncols};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
  cuuint32_t box[2] = {(cuuint32_t)GT_BOX_ROWS, (cuuint32_t)box_cols};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// deterministic (fixed-order) sum of the per-CTA partials; mirrors the lower triangle if sym.
// A CTA of GRED_SL warps produces 32 consecutive elements: warp s adds the partials c = s, s + GRED_SL,
// ... of its 32 elements (coalesced 256-byte reads, all loads of a thread independent), then the
// GRED_SL slice sums are added in a fixed order through shared memory.  The replicated tail of every
// Gram call: 22 us with one thread per element walking all ~148-296 partials, ~5 us in this form.
constexpr int GRED_SL = 8;
__global__ void __launch_bounds__(GRED_SL * 32)
gram_reduce_kernel(const double* __restrict__ partial, int ncta, int PB, int QB, int p, int q,
                   int sym, double* __restrict__ C, int ldc, double* __restrict__ Ct,
                   const int* __restrict__ live) {
  if (live && *live == 0) return;   // predicated step of a speculative ortho chain (engine.cu)
  __shared__ double red[GRED_SL][32];
  const int lane = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int idx = blockIdx.x * 32 + lane;
  const bool valid = idx < p * q;
  int i = 0, j = 0;
  double s0 = 0.0, s1 = 0.0;
  if (valid) {
    i = idx % p;
    j = idx / p;
    int si = i, sj = j;
    if (sym && j > i) { si = j; sj = i; }
    const size_t step = (size_t)PB * QB;
    const double* src = partial + si + (size_t)sj * PB + (size_t)sl * step;
    int c = sl;
#pragma unroll 4
    for (; c + GRED_SL < ncta; c += 2 * GRED_SL) {
      s0 += src[0];
      s1 += src[(size_t)GRED_SL * step];
      src += (size_t)2 * GRED_SL * step;
    }
    if (c < ncta) s0 += src[0];
  }
  red[sl][lane] = s0 + s1;
  __syncthreads();
  if (sl == 0 && valid) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < GRED_SL; ++w) v += red[w][lane];
    C[i + (size_t)j * ldc] = v;
    if (Ct) Ct[j + (size_t)i * ldc] = v;  // mirror of an off-diagonal block of a symmetric product
  }
}

// The same reduction finished by the all-reduce over the peer windows (kernels.h): phase 1 stores the
// locally reduced element into this rank's slot of every rank's window and the last CTA to get
// there publishes the epoch; phase 2 waits for the epochs of all senders in the own window and adds
// the slots in rank order.  All CTAs of the grid are co-resident (<= 512 CTAs of 256 threads).
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__global__ void __launch_bounds__(GRED_SL * 32)
gram_reduce_peer_kernel(const double* partial, int ncta, int PB, int QB, int p, int q, int sym,
                        double* C, int ldc, double* Ct, const int* __restrict__ live,
                        PeerWin w, int max_from) {
  if (live && *live == 0) return;   // the same on every rank: the decisions are taken on all-reduced data
  __shared__ double red[GRED_SL][32];
  const int lane = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int idx = blockIdx.x * 32 + lane;
  const bool valid = idx < p * q;
  const unsigned long long e = *reinterpret_cast<volatile unsigned long long*>(&w.state->epoch) + 1;
  const size_t half = (size_t)(e & 1) * w.nranks * PEER_CAP;
  int i = 0, j = 0;
  double s0 = 0.0, s1 = 0.0;
  if (valid) {
    i = idx % p;
    j = idx / p;
    int si = i, sj = j;
    if (sym && j > i) { si = j; sj = i; }
    const size_t step = (size_t)PB * QB;
    const double* src = partial + si + (size_t)sj * PB + (size_t)sl * step;
    int c = sl;
#pragma unroll 4
    for (; c + GRED_SL < ncta; c += 2 * GRED_SL) {
      s0 += src[0];
      s1 += src[(size_t)GRED_SL * step];
      src += (size_t)2 * GRED_SL * step;
    }
    if (c < ncta) s0 += src[0];
  }
  red[sl][lane] = s0 + s1;
  __syncthreads();
  // ---- phase 1: warp r stores the 32 values into this rank's slot of rank r's window
  if (valid) {
    double v = 0.0;
#pragma unroll
    for (int x = 0; x < GRED_SL; ++x) v += red[x][lane];
    for (int r = sl; r < w.nranks; r += GRED_SL) w.data[r][half + (size_t)w.rank * PEER_CAP + idx] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(&w.state->arrive, 1u);
    if (t == gridDim.x - 1) {   // every CTA's stores are ordered before its arrival: publish the epoch
      w.state->arrive = 0;
      __threadfence_system();
      for (int r = 0; r < w.nranks; ++r) st_release_sys_u64(&w.flags[r][w.rank], e);
    }
    // ---- phase 2: wait for every sender's epoch in the own window (bounded: a lost peer must not hang the GPU)
    const unsigned long long t0 = global_timer_ns();
    for (int r = 0; r < w.nranks; ++r) {
      const unsigned long long* f = &w.flags[w.rank][r];
      while (ld_acquire_sys_u64(f) < e) {
        if (global_timer_ns() - t0 > 20000000000ull) { w.state->error = 1; break; }
      }
    }
  }
  __syncthreads();
  if (sl == 0 && valid) {
    const double* slot = w.data[w.rank] + half + idx;
    double v = __ldcg(slot);
    if (idx < max_from) {
      for (int r = 1; r < w.nranks; ++r) v += __ldcg(slot + (size_t)r * PEER_CAP);
    } else {
      for (int r = 1; r < w.nranks; ++r) v = fmax(v, __ldcg(slot + (size_t)r * PEER_CAP));
    }
    C[i + (size_t)j * ldc] = v;
    if (Ct) Ct[j + (size_t)i * ldc] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(&w.state->depart, 1u);
    if (t == gridDim.x - 1) {   // nobody reads the epoch any more: the call is complete
      w.state->depart = 0;
      w.state->epoch = e;
    }
  }
}

// reduction of the partials of one block: plain, or finished by the peer all-reduce when the engine
// asked for it (g_fuse_allreduce) and the window exists
static void launch_gram_reduce(cudaStream_t st, const double* partial, int ncta, int PB, int QB, int pb, int qb, int sym,
                               double* C, int ldc, double* Ct) {
  const int tot = pb * qb;
  if (g_fuse_allreduce && g_peerwin.nranks > 1 && tot > PEER_CAP) {   // the engine would skip its own all-reduce
    std::fprintf(stderr, "diaglib_b200: a %d x %d Gram block exceeds a peer-window slot\n", pb, qb);
    std::abort();
  }
  if (g_fuse_allreduce && g_peerwin.nranks > 1)
    gram_reduce_peer_kernel<<<(tot + 31) / 32, GRED_SL * 32, 0, st>>>(partial, ncta, PB, QB, pb, qb, sym, C, ldc, Ct, g_live,
                                                                       g_peerwin, 1 << 30);
  else
    gram_reduce_kernel<<<(tot + 31) / 32, GRED_SL * 32, 0, st>>>(partial, ncta, PB, QB, pb, qb, sym, C, ldc, Ct, g_live);
  ++g_launches;
}

// Build a balanced task schedule for a p x q block (tile counts ntp x ntq).
GramSched make_sched_uncached(int ntp, int ntq, bool sym_lower, int nwarps);
// the schedule depends on the tile counts only: built once per shape (the refinement below costs ~0.1 ms of
// host time, a solve launches ~1500 Gram kernels)
GramSched make_sched(int ntp, int ntq, bool sym_lower, int nwarps) {
  static std::map<uint32_t, GramSched> cache;
  const uint32_t key = (uint32_t)ntp | ((uint32_t)ntq << 8) | ((uint32_t)nwarps << 16) | (sym_lower ? 1u << 24 : 0u);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  const GramSched s = make_sched_uncached(ntp, ntq, sym_lower, nwarps);
  cache.emplace(key, s);
  return s;
}
GramSched make_sched_uncached(int ntp, int ntq, bool sym_lower, int nwarps) {
  struct T { int ti0, tj0, nc0, nc1, cnt; };
  const int shapes[4][2] = {{2, 4}, {2, 2}, {1, 2}, {1, 1}};
  std::vector<T> tasks;
  for (int sh = 0; sh < 4; ++sh) {
    const int tr = shapes[sh][0], tc = shapes[sh][1];
    tasks.clear();
    for (int ti0 = 0; ti0 < ntp; ti0 += tr)
      for (int tj0 = 0; tj0 < ntq; tj0 += tc) {
        int nc[2] = {0, 0};
        for (int r = 0; r < tr; ++r)
          for (int c = 0; c < tc; ++c) {
            const int ti = ti0 + r, tj = tj0 + c;
            if (ti < ntp && tj < ntq && (!sym_lower || tj <= ti)) nc[r] = c + 1;  // valid tiles form a prefix
          }
        if (nc[0] + nc[1]) tasks.push_back({ti0, tj0, nc[0], nc[1], nc[0] + nc[1]});
      }
    // the coarsest shape that keeps every warp busy; finer shapes only if they still fit 32 slots
    if ((int)tasks.size() >= nwarps || sh == 3) break;
    // peek: would the next finer shape overflow the 2 slots per warp?
    const int ntr = shapes[sh + 1][0], ntc = shapes[sh + 1][1];
    int nxt = 0;
    for (int ti0 = 0; ti0 < ntp; ti0 += ntr)
      for (int tj0 = 0; tj0 < ntq; tj0 += ntc) {
        bool any = false;
        for (int r = 0; r < ntr && !any; ++r)
          for (int c = 0; c < ntc && !any; ++c)
            any = (ti0 + r < ntp && tj0 + c < ntq && (!sym_lower || tj0 + c <= ti0 + r));
        nxt += any;
      }
    if (nxt > 2 * nwarps) break;
  }
  // LPT assignment: tasks by decreasing size onto the least-loaded SM sub-partition (the FP64 tensor pipe is
  // per sub-partition), then onto its least-loaded warp with a free slot.  Returns false on slot overflow.
  auto assign = [&](std::vector<T> ts, GramSched* out, int* load_q) -> bool {
    std::stable_sort(ts.begin(), ts.end(), [](const T& a, const T& b) { return a.cnt > b.cnt; });
    int load_w[GR_WARPS] = {0};
    int nslot[GR_WARPS] = {0};
    for (int qd = 0; qd < 4; ++qd) load_q[qd] = 0;
    for (const T& t : ts) {
      int bq = -1;
      for (int qd = 0; qd < 4; ++qd) {
        bool has_free = false;
        for (int w = qd; w < nwarps; w += 4) has_free |= nslot[w] < 2;
        if (has_free && (bq < 0 || load_q[qd] < load_q[bq])) bq = qd;
      }
      if (bq < 0) return false;
      int bw = -1;
      for (int w = bq; w < nwarps; w += 4)
        if (nslot[w] < 2 && (bw < 0 || load_w[w] < load_w[bw])) bw = w;
      if (out) out->t[bw][nslot[bw]] = GramTask{(uint8_t)t.ti0, (uint8_t)t.tj0, (uint8_t)t.nc0, (uint8_t)t.nc1};
      ++nslot[bw];
      load_w[bw] += t.cnt;
      load_q[bq] += t.cnt;
    }
    return true;
  };
  auto cost = [](const int* lq) {   // (largest sub-partition load, then the spread)
    int mx = 0, sq = 0;
    for (int qd = 0; qd < 4; ++qd) { mx = std::max(mx, lq[qd]); sq += lq[qd] * lq[qd]; }
    return std::make_pair(mx, sq);
  };
  // Refinement: while slots are left, split one task (a two-row task into its rows, a one-row task into two
  // column halves) if that lowers the largest sub-partition load.  The 74 x 37 overlap of ortho_vs_x (10 x 5
  // tiles as 2 x 2 tasks) goes from loads 14/12/12/12 to 13/13/12/12, the 74 x 74 first-iteration Gram from
  // 15/12/14/14 to 14/13/14/14: the kernel's DMMA pipe share follows the largest load.
  int lq[4];
  if (!assign(tasks, nullptr, lq)) { std::fprintf(stderr, "diaglib_b200: gram schedule overflow\n"); std::abort(); }
  while ((int)tasks.size() < 2 * nwarps) {
    bool improved = false;
    std::vector<size_t> order(tasks.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return tasks[a].cnt < tasks[b].cnt; });
    for (size_t oi = 0; oi < order.size() && !improved; ++oi) {
      const T t = tasks[order[oi]];
      T a, b;
      if (t.nc0 > 0 && t.nc1 > 0) {
        a = T{t.ti0, t.tj0, t.nc0, 0, t.nc0};
        b = T{t.ti0 + 1, t.tj0, t.nc1, 0, t.nc1};
      } else if (t.nc1 == 0 && t.nc0 >= 2) {
        const int h = t.nc0 / 2;
        a = T{t.ti0, t.tj0, t.nc0 - h, 0, t.nc0 - h};
        b = T{t.ti0, t.tj0 + (t.nc0 - h), h, 0, h};
      } else {
        continue;
      }
      std::vector<T> cand = tasks;
      cand[order[oi]] = a;
      cand.push_back(b);
      int lq2[4];
      if (assign(cand, nullptr, lq2) && cost(lq2) < cost(lq)) {
        tasks.swap(cand);
        for (int qd = 0; qd < 4; ++qd) lq[qd] = lq2[qd];
        improved = true;
      }
    }
    if (!improved) break;
  }
  GramSched s{};
  assign(tasks, &s, lq);
  return s;
}

// stage length: as long as three stages fit in shared memory, capped at 128 rows (cp.async
// kernels) or 256 rows (bulk-copy producer: longer column segments per copy)
int pick_kt(int cols, int cap = 128) {
  const int budget = 200 * 1024 / (GR_STAGES * 8);  // doubles per stage
  const int cand[] = {256, 192, 128, 64, 32, 16};
  for (int kt : cand)
    if (kt <= cap && cols * (kt + 4) <= budget) return kt;
  return 16;
}

}  // namespace

// host-side view of a schedule for the tests: cover[ti + tj * ntp] = how many tasks compute tile (ti, tj),
// load4[q] = tiles assigned to SM sub-partition q.  Returns the number of tasks.
int gram_schedule_cover(int ntp, int ntq, bool sym_lower, int nwarps, int* cover, int* load4) {
  const GramSched s = make_sched(ntp, ntq, sym_lower, nwarps);
  for (int i = 0; i < ntp * ntq; ++i) cover[i] = 0;
  for (int q = 0; q < 4; ++q) load4[q] = 0;
  int ntasks = 0;
  for (int w = 0; w < nwarps; ++w)
    for (int sl = 0; sl < 2; ++sl) {
      const GramTask& t = s.t[w][sl];
      if (t.nc0 + t.nc1 == 0) continue;
      ++ntasks;
      for (int c = 0; c < t.nc0; ++c) { ++cover[t.ti0 + (t.tj0 + c) * ntp]; ++load4[w & 3]; }
      for (int c = 0; c < t.nc1; ++c) { ++cover[(t.ti0 + 1) + (t.tj0 + c) * ntp]; ++load4[w & 3]; }
    }
  return ntasks;
}

size_t gram_scratch_bytes(int p, int q, int num_sms) {
  const int pb = std::min(p, GR_MAXB), qb = std::min(q, GR_MAXB);
  const int PB = (pb + 7) / 8 * 8, QB = (qb + 7) / 8 * 8;
  return (size_t)num_sms * PB * QB * sizeof(double);
}

void gram_tn(cudaStream_t st, int num_sms, int64_t n, const double* A, int64_t lda, int p, const double* B,
             int64_t ldb, int q, double* C, int ldc, bool sym_lower, double* partial) {
  if (p <= 0 || q <= 0) return;
  static bool attr_set = false;
  if (!attr_set) {
    DLB_CUDA_CHECK(cudaFuncSetAttribute(gram_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    DLB_CUDA_CHECK(cudaFuncSetAttribute(gram_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const bool al16 = aligned16(A) && aligned16(B) && (lda % 2 == 0) && (ldb % 2 == 0);
  for (int p0 = 0; p0 < p; p0 += GR_MAXB)
    for (int q0 = 0; q0 < q; q0 += GR_MAXB) {
      const bool diag_blk = sym_lower && p0 == q0;
      if (sym_lower && q0 > p0) continue;  // mirrored below
      const int pb = std::min(GR_MAXB, p - p0), qb = std::min(GR_MAXB, q - q0);
      const int ntp = (pb + 7) / 8, ntq = (qb + 7) / 8;
      const int PB = ntp * 8, QB = ntq * 8;
      const double* Ab = A + (int64_t)p0 * lda;
      const double* Bb = B + (int64_t)q0 * ldb;
      const int same = (Ab == Bb && lda == ldb && pb == qb) ? 1 : 0;
      const int cols = PB + (same ? 0 : QB);
      int KT = pick_kt(cols);
      const int ncoarse = ((ntp + 1) / 2) * ((ntq + 3) / 4);
      // bulk-copy producer: only when a column segment of a stage is >= 1 KB (narrow blocks);
      // cp.async producers otherwise; the barrier-synchronised kernel is the fallback
      const bool no_ws_gram = g_disable_ws || (g_ws_mask & 1);
      const bool use_bulk = al16 && (n % 2 == 0) && !no_ws_gram && !(g_ws_mask & 4) && KT >= 128 && ncoarse <= 2 * GRW_CONS;
      if (use_bulk) KT = pick_kt(cols, 256);
      const bool use_wsc = !use_bulk && !no_ws_gram && ncoarse <= 2 * GRC_CONS;
      const GramSched sched = make_sched(ntp, ntq, diag_blk, use_bulk ? GRW_CONS : (use_wsc ? GRC_CONS : GR_WARPS));
      const int64_t nchunks = (n + KT - 1) / KT;
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(num_sms, nchunks));
      const size_t smem = (size_t)GR_STAGES * cols * (KT + 4) * sizeof(double);
      const size_t smem_ws = smem + 2 * GR_STAGES * sizeof(double);
      static bool ws_attr = false;
      if (!ws_attr) {
        DLB_CUDA_CHECK(cudaFuncSetAttribute(gram_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        DLB_CUDA_CHECK(cudaFuncSetAttribute(gram_wsc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        DLB_CUDA_CHECK(cudaFuncSetAttribute(gram_wsc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ws_attr = true;
      }
      bool launched = false;
      // (narrow blocks whose column segments reach 1 KB per stage are served better by the 1-D
      //  bulk-copy producer below: 3.9 vs 3.2 TB/s on the 37-column metric of ortho_cd)
      if (al16 && !use_bulk && !no_ws_gram && !g_disable_tma && ncoarse <= 2 * GRW_CONS && n < (int64_t)1 << 31) {
        // TMA-tiled kernel: stage length from the unpadded box footprint
        int kt = 128;
        while (kt > 16 && (size_t)GR_STAGES * cols * kt * 8 > 200 * 1024) kt >>= 1;
        CUtensorMap tmA, tmB;
        if (make_tmap(&tmA, Ab, n, pb, lda, PB) && (same || make_tmap(&tmB, Bb, n, qb, ldb, QB))) {
          static bool tma_attr = false;
          if (!tma_attr) {
            DLB_CUDA_CHECK(cudaFuncSetAttribute(gram_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            tma_attr = true;
          }
          if (same) tmB = tmA;
          const GramSched sch = make_sched(ntp, ntq, diag_blk, GRW_CONS);
          const int64_t nch = (n + kt - 1) / kt;
          const int g = (int)std::max<int64_t>(1, std::min<int64_t>(num_sms, nch));
          const size_t sm = (size_t)GR_STAGES * cols * kt * 8 + 2 * GR_STAGES * sizeof(uint64_t) + 1024;
          gram_tma_kernel<<<g, GR_THREADS, sm, st>>>(n, tmA, tmB, same, kt, sch, partial, PB, QB, g_live);
          ++g_launches;
          double* Cblk = C + p0 + (size_t)q0 * ldc;
          double* Cmir = (sym_lower && !diag_blk) ? C + q0 + (size_t)p0 * ldc : nullptr;
          launch_gram_reduce(st, partial, g, PB, QB, pb, qb, diag_blk ? 1 : 0, Cblk, ldc, Cmir);
          launched = true;
        }
      }
      if (launched) continue;
      if (use_bulk)
        gram_ws_kernel<<<grid, GRW_THREADS, smem_ws, st>>>(n, Ab, lda, pb, Bb, ldb, qb, same, KT, sched, partial, PB, QB, g_live);
      else if (use_wsc && al16)
        gram_wsc_kernel<true><<<grid, GR_THREADS, smem_ws, st>>>(n, Ab, lda, pb, Bb, ldb, qb, same, KT, sched, partial, PB, QB, g_live);
      else if (use_wsc)
        gram_wsc_kernel<false><<<grid, GR_THREADS, smem_ws, st>>>(n, Ab, lda, pb, Bb, ldb, qb, same, KT, sched, partial, PB, QB, g_live);
      else if (al16)
        gram_kernel<true><<<grid, GR_THREADS, smem, st>>>(n, Ab, lda, pb, Bb, ldb, qb, same, KT, sched, partial, PB, QB, g_live);
      else
        gram_kernel<false><<<grid, GR_THREADS, smem, st>>>(n, Ab, lda, pb, Bb, ldb, qb, same, KT, sched, partial, PB, QB, g_live);
      ++g_launches;
      double* Cblk = C + p0 + (size_t)q0 * ldc;
      double* Cmir = (sym_lower && !diag_blk) ? C + q0 + (size_t)p0 * ldc : nullptr;
      launch_gram_reduce(st, partial, grid, PB, QB, pb, qb, diag_blk ? 1 : 0, Cblk, ldc, Cmir);
    }
  DLB_CUDA_CHECK(cudaGetLastError());
}

void peer_allreduce(cudaStream_t st, double* d, int count, int max_from) {
  if (count <= 0 || g_peerwin.nranks <= 1) return;
  // the buffer itself is the one "partial": p = count, q = 1
  gram_reduce_peer_kernel<<<(count + 31) / 32, GRED_SL * 32, 0, st>>>(d, 1, count, 1, count, 1, 0, d, count, nullptr, g_live,
                                                                     g_peerwin, max_from);
  ++g_launches;
  DLB_CUDA_CHECK(cudaGetLastError());
}

// =====================================================================================
// block_mul
// =====================================================================================
namespace {

constexpr int BM_THREADS = 256;
constexpr int BM_RT = 128;            // rows per CTA tile (8 warps x 16 rows)
constexpr int BM_SV = BM_RT + 4;      // padded stride of a V column in smem
constexpr int BM_KC = 16;             // columns of V (rows of C) per stage
constexpr int BM_SC = BM_KC + 4;      // padded stride of a C column chunk
constexpr int BM_STAGES = 4;

template <int NQT, bool ALIGN16>
__global__ void __launch_bounds__(BM_THREADS)
blockmul_kernel(int64_t n, const double* __restrict__ V, int64_t ldv, int p, const double* __restrict__ C, int ldc,
                int q, double alpha, double beta, double* Y, int64_t ldy, const int* __restrict__ live) {
  if (live && *live == 0) return;   // predicated step of a speculative ortho chain (engine.cu)
  extern __shared__ __align__(16) double smem[];
  constexpr int QB = NQT * 8;
  constexpr int STAGE = BM_KC * BM_SV + QB * BM_SC;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * BM_RT;
  const int nk = (p + BM_KC - 1) / BM_KC;

  double acc[2][NQT][2];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < NQT; ++c) acc[r][c][0] = acc[r][c][1] = 0.0;

  auto issue = [&](int kc) {
    if (kc < nk) {
      double* sV = smem + (kc % BM_STAGES) * STAGE;
      double* sC = sV + BM_KC * BM_SV;
      const int k0 = kc * BM_KC;
      if (ALIGN16) {
        for (int id = tid; id < BM_KC * (BM_RT / 2); id += BM_THREADS) {
          const int col = id / (BM_RT / 2), part = id % (BM_RT / 2);
          const int64_t row = row0 + part * 2;
          int64_t rem = (k0 + col < p) ? (n - row) * 8 : 0;
          const int bytes = rem >= 16 ? 16 : (rem > 0 ? (int)rem : 0);
          const double* src = bytes > 0 ? V + (int64_t)(k0 + col) * ldv + row : V;
          cp_async16(sV + col * BM_SV + part * 2, src, bytes);
        }
      } else {
        for (int id = tid; id < BM_KC * BM_RT; id += BM_THREADS) {
          const int col = id / BM_RT, part = id % BM_RT;
          const int64_t row = row0 + part;
          const int bytes = (k0 + col < p && row < n) ? 8 : 0;
          const double* src = bytes > 0 ? V + (int64_t)(k0 + col) * ldv + row : V;
          cp_async8(sV + col * BM_SV + part, src, bytes);
        }
      }
      for (int id = tid; id < QB * BM_KC; id += BM_THREADS) {
        const int j = id / BM_KC, kk = id % BM_KC;
        const int bytes = (j < q && k0 + kk < p) ? 8 : 0;
        const double* src = bytes > 0 ? C + (size_t)j * ldc + k0 + kk : C;
        cp_async8(sC + j * BM_SC + kk, src, bytes);
      }
    }
    cp_async_commit();
  };

#pragma unroll
  for (int s = 0; s < BM_STAGES - 1; ++s) issue(s);

  const int a_off = (lane & 3) * BM_SV + warp * 16 + (lane >> 2);
  const int b_off = (lane >> 2) * BM_SC + (lane & 3);
  for (int kc = 0; kc < nk; ++kc) {
    cp_async_wait<BM_STAGES - 2>();
    __syncthreads();
    issue(kc + BM_STAGES - 1);
    const double* sV = smem + (kc % BM_STAGES) * STAGE;
    const double* sC = sV + BM_KC * BM_SV;
#pragma unroll
    for (int k4 = 0; k4 < BM_KC / 4; ++k4) {
      double a[2], b[NQT];
#pragma unroll
      for (int r = 0; r < 2; ++r) a[r] = sV[k4 * 4 * BM_SV + a_off + r * 8];
#pragma unroll
      for (int c = 0; c < NQT; ++c) b[c] = sC[c * 8 * BM_SC + k4 * 4 + b_off];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < NQT; ++c) dmma884(acc[r][c][0], acc[r][c][1], a[r], b[c]);
    }
  }
  cp_async_wait<0>();

#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int64_t row = row0 + warp * 16 + r * 8 + (lane >> 2);
    if (row >= n) continue;
#pragma unroll
    for (int c = 0; c < NQT; ++c) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = c * 8 + (lane & 3) * 2 + e;
        if (col < q) {
          double* dst = Y + row + (int64_t)col * ldy;
          double v = alpha * acc[r][c][e];
          if (beta != 0.0) v += beta * (*dst);
          *dst = v;
        }
      }
    }
  }
}

// Persistent variant: C (p x q) stays resident in shared memory for the CTA's lifetime, the
// CTA walks over row tiles (stride gridDim.x) and the cp.async ring runs continuously across
// (tile, k-chunk) pairs, so there is no pipeline fill/drain bubble per row tile.
template <int NQT, bool ALIGN16>
__global__ void __launch_bounds__(BM_THREADS)
blockmul_persistent_kernel(int64_t n, const double* __restrict__ V, int64_t ldv, int p, const double* __restrict__ C,
                           int ldc, int q, double alpha, double beta, double* Y, int64_t ldy, int PS, int tri, const int* __restrict__ live) {
  if (live && *live == 0) return;   // predicated step of a speculative ortho chain (engine.cu)
  extern __shared__ __align__(16) double smem[];
  constexpr int QB = NQT * 8;
  constexpr int STAGE = BM_KC * BM_SV;
  double* sC = smem;                       // [QB][PS], zero padded
  double* ring = smem + (size_t)QB * PS;   // BM_STAGES x STAGE
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nk = (p + BM_KC - 1) / BM_KC;
  const int64_t ntiles = (n + BM_RT - 1) / BM_RT;
  const int64_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t my_chunks = my_tiles * nk;

  for (int id = tid; id < QB * PS; id += BM_THREADS) {
    const int j = id / PS, k = id - j * PS;
    sC[id] = (j < q && k < p) ? C[(size_t)j * ldc + k] : 0.0;
  }

  auto issue = [&](int64_t c) {
    if (c < my_chunks) {
      const int64_t ti = c / nk;
      const int kc = (int)(c - ti * nk);
      const int64_t row0 = (blockIdx.x + ti * gridDim.x) * BM_RT;
      double* sV = ring + (c % BM_STAGES) * STAGE;
      const int k0 = kc * BM_KC;
      if (ALIGN16) {
        for (int id = tid; id < BM_KC * (BM_RT / 2); id += BM_THREADS) {
          const int col = id / (BM_RT / 2), part = id % (BM_RT / 2);
          const int64_t row = row0 + part * 2;
          int64_t rem = (k0 + col < p) ? (n - row) * 8 : 0;
          const int bytes = rem >= 16 ? 16 : (rem > 0 ? (int)rem : 0);
          const double* src = bytes > 0 ? V + (int64_t)(k0 + col) * ldv + row : V;
          cp_async16(sV + col * BM_SV + part * 2, src, bytes);
        }
      } else {
        for (int id = tid; id < BM_KC * BM_RT; id += BM_THREADS) {
          const int col = id / BM_RT, part = id % BM_RT;
          const int64_t row = row0 + part;
          const int bytes = (k0 + col < p && row < n) ? 8 : 0;
          const double* src = bytes > 0 ? V + (int64_t)(k0 + col) * ldv + row : V;
          cp_async8(sV + col * BM_SV + part, src, bytes);
        }
      }
    }
    cp_async_commit();
  };

#pragma unroll
  for (int s = 0; s < BM_STAGES - 1; ++s) issue(s);

  double acc[2][NQT][2];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < NQT; ++c) acc[r][c][0] = acc[r][c][1] = 0.0;

  const int a_off = (lane & 3) * BM_SV + warp * 16 + (lane >> 2);
  const int b_off = (lane >> 2) * PS + (lane & 3);
  int kc = 0;
  int64_t ti = 0;
  for (int64_t c = 0; c < my_chunks; ++c) {
    cp_async_wait<BM_STAGES - 2>();
    __syncthreads();
    issue(c + BM_STAGES - 1);
    const double* sV = ring + (c % BM_STAGES) * STAGE;
    const double* sCk = sC + kc * BM_KC + b_off;
#pragma unroll
    for (int k4 = 0; k4 < BM_KC / 4; ++k4) {
      double a[2], b[NQT];
#pragma unroll
      for (int r = 0; r < 2; ++r) a[r] = sV[k4 * 4 * BM_SV + a_off + r * 8];
#pragma unroll
      for (int cc = 0; cc < NQT; ++cc) b[cc] = sCk[cc * 8 * PS + k4 * 4];
      if (!tri) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int cc = 0; cc < NQT; ++cc) dmma884(acc[r][cc][0], acc[r][cc][1], a[r], b[cc]);
      } else {
        // C upper triangular: rows k >= 8*(cc+1) of tile column cc are zero, skip them
        const int kbase = kc * BM_KC + k4 * 4;
#pragma unroll
        for (int cc = 0; cc < NQT; ++cc)
          if (kbase < (cc + 1) * 8) {
#pragma unroll
            for (int r = 0; r < 2; ++r) dmma884(acc[r][cc][0], acc[r][cc][1], a[r], b[cc]);
          }
      }
    }
    if (++kc == nk) {
      const int64_t row0 = (blockIdx.x + ti * gridDim.x) * BM_RT;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int64_t row = row0 + warp * 16 + r * 8 + (lane >> 2);
#pragma unroll
        for (int cc = 0; cc < NQT; ++cc) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = cc * 8 + (lane & 3) * 2 + e;
            if (row < n && col < q) {
              double* dst = Y + row + (int64_t)col * ldy;
              double v = alpha * acc[r][cc][e];
              if (beta != 0.0) v += beta * (*dst);
              *dst = v;
            }
            acc[r][cc][e] = 0.0;
          }
        }
      }
      kc = 0;
      ++ti;
    }
  }
  cp_async_wait<0>();
}

// Warp-specialised variant: one producer warp feeds the shared-memory ring with 1-D bulk async
// copies (TMA engine, one 1 KB column segment per copy, completion counted on an mbarrier) and
// eight consumer warps run the DMMA loop.  There is no CTA-wide barrier in the main loop:
// consumers wait on full[stage], release the stage on empty[stage], and drift freely, which
// keeps the FP64 tensor pipe busy while other warps wait for data.
constexpr int BMW_STAGES = 4;

// NCONS consumer warps (16 rows each) + 1 producer warp; the row tile is NCONS*16 rows, so a
// bulk copy moves NCONS*128 bytes: 2 KB with 16 consumers, which the copy engine needs to get
// past ~4.4 TB/s on the HBM-bound shapes (trmm, 74x37).
// GRAM = true additionally accumulates G = Y'^T Y' of the stored block (lower tiles), the metric
// the next ortho_cd pass needs (diaglib.f90:3256), so that U is not read again: the finished
// 16 x 40 accumulator tile of a warp is re-used as DMMA operands through warp shuffles (the
// C-fragment of lane (row g, columns 2t,2t+1) becomes the A/B fragment of lane (column, row)).
// one 16-column chunk (index KC, compile time) of Y += V C for an upper triangular C: only the tiles
// with columns >= the chunk's rows are multiplied, decided at compile time
template <int NQT, int KC>
__device__ __forceinline__ void bm_chunk_tri(double (&acc)[2][NQT][2], const double* sV, const double* sCb, int a_off, int SV,
                                             int PS) {
  const double* sCk = sCb + KC * BM_KC;
#pragma unroll
  for (int k4 = 0; k4 < BM_KC / 4; ++k4) {
    const int kbase = KC * BM_KC + k4 * 4;
    if (kbase >= NQT * 8) continue;
    double a[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) a[r] = sV[k4 * 4 * SV + a_off + r * 8];
#pragma unroll
    for (int cc = 0; cc < NQT; ++cc) {
      if (kbase < (cc + 1) * 8) {
        const double b = sCk[cc * 8 * PS + k4 * 4];
#pragma unroll
        for (int r = 0; r < 2; ++r) dmma884(acc[r][cc][0], acc[r][cc][1], a[r], b);
      }
    }
  }
}

template <int NQT, int NCONS, bool GRAM>
__global__ void __launch_bounds__((NCONS + 1) * 32)
blockmul_ws_kernel(int64_t n, const double* __restrict__ V, int64_t ldv, int p, const double* __restrict__ C, int ldc,
                   int q, double alpha, double beta, double* Y, int64_t ldy, int PS, int tri, double* gpartial, const int* __restrict__ live,
                   int dbg) {
  if (live && *live == 0) return;   // predicated step of a speculative ortho chain (engine.cu)
  extern __shared__ __align__(16) double smem[];
  constexpr int QB = NQT * 8;
  constexpr int BMW_CONS = NCONS;
  constexpr int BMW_THREADS = (NCONS + 1) * 32;
  constexpr int RT = NCONS * 16;
  constexpr int SV = RT + 4;
  constexpr int STAGE = BM_KC * SV;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);   // BMW_STAGES
  uint64_t* empty = full + BMW_STAGES;                   // BMW_STAGES
  double* sC = smem + 2 * BMW_STAGES;                    // [QB][PS], zero padded
  double* ring = sC + (size_t)QB * PS;                   // BMW_STAGES x STAGE
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nk = (p + BM_KC - 1) / BM_KC;
  const int64_t ntiles = (n + RT - 1) / RT;
  const int64_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
    for (int s = 0; s < BMW_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], BMW_CONS); }
    mbar_fence_init();
  }
  for (int id = tid; id < QB * PS; id += BMW_THREADS) {
    const int j = id / PS, k = id - j * PS;
    sC[id] = (j < q && k < p) ? C[(size_t)j * ldc + k] : 0.0;
  }
  for (int id = tid; id < BMW_STAGES * STAGE; id += BMW_THREADS) ring[id] = 0.0;  // stale data must stay finite
  fence_proxy_async();
  __syncthreads();

  if (warp == BMW_CONS) {
    // ---------------- producer ----------------
    int s = 0;
    uint32_t ph = 0;
    if (dbg & 4) asm volatile("fence.proxy.async.global;\n" ::: "memory");
    for (int64_t ti = 0; ti < my_tiles; ++ti) {
      const int64_t row0 = (blockIdx.x + ti * gridDim.x) * RT;
      const int64_t rem = n - row0;
      const uint32_t rows = rem < RT ? (uint32_t)rem : (uint32_t)RT;
      for (int kc = 0; kc < nk; ++kc) {
        const int k0 = kc * BM_KC;
        const int ncols = (p - k0) < BM_KC ? (p - k0) : BM_KC;
        mbar_wait(&empty[s], ph ^ 1);
        if (lane == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)ncols * rows * 8u);
        __syncwarp();
        if (lane < ncols)
          bulk_g2s(ring + (size_t)s * STAGE + lane * SV, V + (int64_t)(k0 + lane) * ldv + row0, rows * 8u, &full[s]);
        if (++s == BMW_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ---------------- consumers ----------------
    double acc[2][NQT][2];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < NQT; ++c) acc[r][c][0] = acc[r][c][1] = 0.0;
    const int a_off = (lane & 3) * SV + warp * 16 + (lane >> 2);
    const int b_off = (lane >> 2) * PS + (lane & 3);
    constexpr int NGT = GRAM ? NQT * (NQT + 1) / 2 : 1;   // lower tiles of the QB x QB metric
    double gacc[NGT][2];
#pragma unroll
    for (int g = 0; g < NGT; ++g) gacc[g][0] = gacc[g][1] = 0.0;
    int s = 0;
    uint32_t ph = 0;
    for (int64_t ti = 0; ti < my_tiles; ++ti) {
      // one k-chunk of the tile: wait for its stage, multiply, release the stage
      auto stage_wait = [&]() -> const double* {
        mbar_wait(&full[s], ph);
        return ring + (size_t)s * STAGE;
      };
      auto stage_release = [&]() {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == BMW_STAGES) { s = 0; ph ^= 1; }
      };
      if (tri == 1 && NQT <= 10) {
        // C upper triangular (the dtrmm of ortho_cd, p = q <= 8 NQT): at most NQT / 2 chunks, written out
        // with the chunk index as a compile-time constant so that the tiles below the diagonal are not
        // even issued.  A predicated-off DMMA still occupies its slot of the FP64 pipe (ncu on the
        // round-1 form: pipe 71 % busy with 56 % of the issued DMMAs predicated off, math-pipe-throttle
        // the top stall), which made the triangular multiply pipe-bound although it needs half the
        // flops: 2.23 -> 1.68 ms at 37 columns, n = 2^24 (0.68 -> 0.91 of the HBM copy peak).
#define DLB_TRI_CHUNK(KC)                                                                   \
  if constexpr (KC * BM_KC < NQT * 8) {                                                     \
    if (KC < nk) {                                                                          \
      const double* sV = stage_wait();                                                      \
      bm_chunk_tri<NQT, KC>(acc, sV, sC + b_off, a_off, SV, PS);                            \
      stage_release();                                                                      \
    }                                                                                       \
  }
        DLB_TRI_CHUNK(0) DLB_TRI_CHUNK(1) DLB_TRI_CHUNK(2) DLB_TRI_CHUNK(3)
        DLB_TRI_CHUNK(4) DLB_TRI_CHUNK(5) DLB_TRI_CHUNK(6) DLB_TRI_CHUNK(7)
#undef DLB_TRI_CHUNK
      } else {
        for (int kc = 0; kc < nk; ++kc) {
          const double* sV = stage_wait();
          if (tri >= 2 && kc * BM_KC >= tri - 2) {
            // rows >= tri - 2 of C are an identity block (Y = V1 C1 + V2) and this chunk lies inside it:
            // its columns of V are ADDED to the accumulators they belong to, no tensor instruction
            const int shift_k = (tri - 2) - kc * BM_KC;   // chunk column of output column `col` = col + shift_k
#pragma unroll
            for (int cc = 0; cc < NQT; ++cc)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int col = cc * 8 + (lane & 3) * 2 + e;
                const int kk = col + shift_k;
                if (kk >= 0 && kk < BM_KC && col < q) {
#pragma unroll
                  for (int r = 0; r < 2; ++r) acc[r][cc][e] += sV[kk * SV + warp * 16 + r * 8 + (lane >> 2)];
                }
              }
          } else if (NQT == 5 && tri <= -2 && (kc + 1) * BM_KC > -tri - 2) {
            // rows >= -tri - 2 of C are an upper triangular block (Y = V1 C1 + V2 T: the deferred triangular multiply of
            // ortho_cd folded into the projection step of ortho_vs_x): a k-step whose four rows lie inside the block
            // only needs the tiles from the one that holds its first row's diagonal on.  The skipped tiles are not
            // issued (a branch on a warp-uniform value, not a predicate: a predicated-off DMMA keeps its pipe slot).
            const int r0 = -tri - 2;
            const double* sCk = sC + kc * BM_KC + b_off;
#pragma unroll
            for (int k4 = 0; k4 < BM_KC / 4; ++k4) {
              const int rb = kc * BM_KC + k4 * 4 - r0;   // first row of this k-step relative to the block
              const int ccmin = rb <= 0 ? 0 : (rb >> 3);
              if (ccmin >= NQT) continue;                 // (rows beyond the block's columns: padding, all zero)
              double a[2];
#pragma unroll
              for (int r = 0; r < 2; ++r) a[r] = sV[k4 * 4 * SV + a_off + r * 8];
#define DLB_TILE(CC)                                                      \
  {                                                                       \
    const double bv = sCk[(CC) * 8 * PS + k4 * 4];                        \
    dmma884(acc[0][CC][0], acc[0][CC][1], a[0], bv);                      \
    dmma884(acc[1][CC][0], acc[1][CC][1], a[1], bv);                      \
  }
              switch (ccmin) {
                case 0: DLB_TILE(0)   // fall through
                case 1: DLB_TILE(1)   // fall through
                case 2: DLB_TILE(2)   // fall through
                case 3: DLB_TILE(3)   // fall through
                default: DLB_TILE(4)
              }
#undef DLB_TILE
            }
          } else {
            const double* sCk = sC + kc * BM_KC + b_off;
#pragma unroll
            for (int k4 = 0; k4 < BM_KC / 4; ++k4) {
              double a[2], b[NQT];
#pragma unroll
              for (int r = 0; r < 2; ++r) a[r] = sV[k4 * 4 * SV + a_off + r * 8];
#pragma unroll
              for (int cc = 0; cc < NQT; ++cc) b[cc] = sCk[cc * 8 * PS + k4 * 4];
              if (NQT <= 10 || tri != 1) {
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                  for (int cc = 0; cc < NQT; ++cc) dmma884(acc[r][cc][0], acc[r][cc][1], a[r], b[cc]);
              } else {
                // 128-column triangular blocks (C5): predicated tiles.  Written out per chunk like the
                // narrower widths, this instantiation makes ptxas demote the 64 accumulators to local
                // memory (2.4 KB stack frame, ortho phase of C5 2x slower), so it keeps the loop.
                const int kbase = kc * BM_KC + k4 * 4;
#pragma unroll
                for (int cc = 0; cc < NQT; ++cc)
                  if (kbase < (cc + 1) * 8) {
#pragma unroll
                    for (int r = 0; r < 2; ++r) dmma884(acc[r][cc][0], acc[r][cc][1], a[r], b[cc]);
                  }
              }
            }
          }
          stage_release();
        }
      }
      const int64_t row0 = (blockIdx.x + ti * gridDim.x) * RT;
      // The multiply by alpha stays although alpha = 1 in every hot call and the DMUL queues for the FP64 pipe behind
      // the other warps' DMMAs (15 % of the projection step's stall samples, ~1 % of its time).  Storing the
      // accumulator registers themselves - which the next tile's DMMAs overwrite right away - made whole
      // solves irreproducible from run to run (tools/determinism_check.py: 2-7 of 8 repetitions differed,
      // one did not converge); with the product in a register of its own every repetition is bit-identical.
#define DLB_STORE_TILE(VEXPR)                                                        \
  _Pragma("unroll") for (int r = 0; r < 2; ++r) {                                    \
    const int64_t row = row0 + warp * 16 + r * 8 + (lane >> 2);                      \
    _Pragma("unroll") for (int cc = 0; cc < NQT; ++cc) {                             \
      _Pragma("unroll") for (int e = 0; e < 2; ++e) {                                \
        const int col = cc * 8 + (lane & 3) * 2 + e;                                 \
        double v = 0.0;                                                              \
        if (row < n && col < q) {                                                    \
          double* dst = Y + row + (int64_t)col * ldy;                                \
          VEXPR;                                                                     \
          if (beta != 0.0) v += beta * (*dst);                                       \
          *dst = v;                                                                  \
        }                                                                            \
        acc[r][cc][e] = GRAM ? v : 0.0;                                              \
      }                                                                              \
    }                                                                                \
  }
#ifndef DLB_DBG
#define DLB_DBG 0
#endif
#if DLB_DBG == 0
      DLB_STORE_TILE(v = alpha * acc[r][cc][e])
#elif DLB_DBG == 1
      DLB_STORE_TILE(v = acc[r][cc][e])
#elif DLB_DBG == 2
      DLB_STORE_TILE(asm volatile("mov.b64 %0, %1;" : "=d"(v) : "d"(acc[r][cc][e])))
#elif DLB_DBG == 3
      if (tri == 1) { DLB_STORE_TILE(v = acc[r][cc][e]) } else { DLB_STORE_TILE(v = alpha * acc[r][cc][e]) }
#elif DLB_DBG == 4
      if (tri == 0) { DLB_STORE_TILE(v = acc[r][cc][e]) } else { DLB_STORE_TILE(v = alpha * acc[r][cc][e]) }
#elif DLB_DBG == 5
      if (tri >= 2 || tri <= -2) { DLB_STORE_TILE(v = acc[r][cc][e]) } else { DLB_STORE_TILE(v = alpha * acc[r][cc][e]) }
#elif DLB_DBG == 6
      __nanosleep(1000);
      DLB_STORE_TILE(v = acc[r][cc][e])
#endif
#undef DLB_STORE_TILE
      if (GRAM) {
        // G += Y'^T Y' over this warp's 16 rows: 4 k-steps of 4 rows
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const int r = ks >> 1;
          const int src = ((4 * (ks & 1) + (lane & 3)) << 2) + (lane >> 3);   // lane holding (row, column pair)
          double f[NQT];
#pragma unroll
          for (int cc = 0; cc < NQT; ++cc) {
            const double x0 = __shfl_sync(0xffffffffu, acc[r][cc][0], src);
            const double x1 = __shfl_sync(0xffffffffu, acc[r][cc][1], src);
            f[cc] = ((lane >> 2) & 1) ? x1 : x0;
          }
          int g = 0;
#pragma unroll
          for (int ti2 = 0; ti2 < NQT; ++ti2)
#pragma unroll
            for (int tj2 = 0; tj2 <= ti2; ++tj2) {
              dmma884(gacc[g][0], gacc[g][1], f[ti2], f[tj2]);
              ++g;
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int cc = 0; cc < NQT; ++cc) acc[r][cc][0] = acc[r][cc][1] = 0.0;
      }
    }
    if (dbg & 8) __threadfence();
    if (GRAM) {
      // reduce the 8 warps' accumulators in a fixed order through the (drained) ring, one partial per CTA
      asm volatile("bar.sync 1, %0;\n" ::"r"(BMW_CONS * 32));
      double* red = ring;   // [warp][tile][lane][2]
#pragma unroll
      for (int g = 0; g < NGT; ++g) {
        red[(((size_t)warp * NGT + g) * 32 + lane) * 2 + 0] = gacc[g][0];
        red[(((size_t)warp * NGT + g) * 32 + lane) * 2 + 1] = gacc[g][1];
      }
      asm volatile("bar.sync 1, %0;\n" ::"r"(BMW_CONS * 32));
      double* out = gpartial + (size_t)blockIdx.x * QB * QB;
      for (int id = tid; id < NGT * 64; id += BMW_CONS * 32) {
        const int g = id >> 6, l = (id >> 1) & 31, e = id & 1;
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < BMW_CONS; ++w) sum += red[(((size_t)w * NGT + g) * 32 + l) * 2 + e];
        // tile g -> (ti2, tj2) of the lower triangle
        int ti2 = 0, acc_g = 0;
        while (acc_g + ti2 + 1 <= g) { acc_g += ti2 + 1; ++ti2; }
        const int tj2 = g - acc_g;
        const int i = ti2 * 8 + (l >> 2), j = tj2 * 8 + (l & 3) * 2 + e;
        out[i + (size_t)j * QB] = sum;
      }
    }
  }
}

__global__ void dbg_noop_kernel() {}
template <int NQT>
void launch_blockmul(cudaStream_t st, int64_t n, const double* V, int64_t ldv, int p, const double* C, int ldc,
                     int q, double alpha, double beta, double* Y, int64_t ldy, int mode) {
  const bool tri = mode == 1;   // the barrier-pipelined fallbacks know the triangular case only (an identity block is just data)
  constexpr int QB = NQT * 8;
  constexpr size_t smem = (size_t)BM_STAGES * (BM_KC * BM_SV + QB * BM_SC) * sizeof(double);
  static bool attr_set = false;
  static int num_sms = 0, occ_a = 0, occ_u = 0;
  if (!attr_set) {
    DLB_CUDA_CHECK(cudaFuncSetAttribute(blockmul_kernel<NQT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DLB_CUDA_CHECK(cudaFuncSetAttribute(blockmul_kernel<NQT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DLB_CUDA_CHECK(cudaFuncSetAttribute(blockmul_persistent_kernel<NQT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    DLB_CUDA_CHECK(cudaFuncSetAttribute(blockmul_persistent_kernel<NQT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int dev = 0;
    DLB_CUDA_CHECK(cudaGetDevice(&dev));
    DLB_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    attr_set = true;
  }
  const bool al16 = aligned16(V) && (ldv % 2 == 0);
  const int64_t ntiles = (n + BM_RT - 1) / BM_RT;
  // persistent path: C resident in shared memory
  const int p16 = (p + 15) / 16 * 16;
  const int PS = p16 + 4;
  const size_t smem_p = ((size_t)QB * PS + (size_t)BM_STAGES * BM_KC * BM_SV) * sizeof(double);
  if (al16 && (n % 2 == 0) && !g_disable_ws && !(g_ws_mask & 2)) {
    const size_t sc_bytes = (2 * BMW_STAGES + (size_t)QB * PS) * sizeof(double);
    const size_t smem16 = sc_bytes + (size_t)BMW_STAGES * BM_KC * (16 * 16 + 4) * sizeof(double);
    const size_t smem8 = sc_bytes + (size_t)BMW_STAGES * BM_KC * (8 * 16 + 4) * sizeof(double);
    static bool ws_attr = false;
    if (!ws_attr) {
      DLB_CUDA_CHECK(cudaFuncSetAttribute(blockmul_ws_kernel<NQT, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      if (NQT == 5)
        DLB_CUDA_CHECK(cudaFuncSetAttribute(blockmul_ws_kernel<5, 16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      ws_attr = true;
    }
    if (NQT == 5 && smem16 <= 220 * 1024 && n >= 256 * 2 && !g_bmul_small_tiles) {
      const int64_t nt16 = (n + 255) / 256;
      const unsigned grid = (unsigned)std::min<int64_t>(nt16, (int64_t)num_sms);
      if (g_dbg & 1) dbg_noop_kernel<<<1, 32, 0, st>>>();
      blockmul_ws_kernel<5, 16, false><<<grid, 17 * 32, smem16, st>>>(n, V, ldv, p, C, ldc, q, alpha, beta, Y, ldy, PS, mode, nullptr, g_live, g_dbg);
      ++g_launches;
      return;
    }
    if (smem8 <= 200 * 1024 && ntiles >= 2) {
      // two CTAs per SM while C (p x q) is small enough to be resident twice, one beyond (Davidson:
      // p = ldu up to ~400 with q <= 40)
      // ONE CTA per SM.  Two co-resident CTAs of this kernel (smem8 <= 110 KB; the round-1 default, ~3 % faster on the
      // 111 -> 37 product) made whole solves irreproducible from run to run - rarely with the shipped epilogue
      // (one non-converged solve in ~40 at n = 2^24), in 9 of 10 repetitions with an epilogue that stores the
      // accumulator registers directly - while every variant with one CTA per SM (this grid, or the 16-consumer
      // kernel above) reproduced all repetitions bit for bit, as do the barrier-pipelined fallbacks
      // (profiles/determinism_bisect_r02.log).  The kernel in isolation and ortho_vs_x as a sequence are reproducible
      // in either form; the interaction was not identified.  DIAGLIB_B200_DBG=32 restores two CTAs per SM.
      const unsigned grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)num_sms * ((smem8 <= 110 * 1024 && (g_dbg & 32)) ? 2 : 1));
      if (g_dbg & 1) dbg_noop_kernel<<<1, 32, 0, st>>>();
      // the request is padded to more than half an SM's shared memory so that the block scheduler cannot place two
      // of these CTAs on one SM even when it has SMs to spare (227 KB attribute set above)
      const size_t smem_req = (g_dbg & 32) ? smem8 : std::max(smem8, (size_t)116 * 1024);
      blockmul_ws_kernel<NQT, 8, false><<<grid, 9 * 32, smem_req, st>>>(n, V, ldv, p, C, ldc, q, alpha, beta, Y, ldy, PS, mode, nullptr, g_live, g_dbg);
      if (g_dbg & 2) dbg_noop_kernel<<<1, 32, 0, st>>>();
      ++g_launches;
      return;
    }
  }
  if (smem_p <= 110 * 1024 && ntiles >= 2) {
    int occ = 0;
    if (al16)
      DLB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, blockmul_persistent_kernel<NQT, true>, BM_THREADS, smem_p));
    else
      DLB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, blockmul_persistent_kernel<NQT, false>, BM_THREADS, smem_p));
    (void)occ_a; (void)occ_u;
    const unsigned grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)num_sms * std::max(1, occ));
    if (al16)
      blockmul_persistent_kernel<NQT, true><<<grid, BM_THREADS, smem_p, st>>>(n, V, ldv, p, C, ldc, q, alpha, beta, Y, ldy, PS, tri ? 1 : 0, g_live);
    else
      blockmul_persistent_kernel<NQT, false><<<grid, BM_THREADS, smem_p, st>>>(n, V, ldv, p, C, ldc, q, alpha, beta, Y, ldy, PS, tri ? 1 : 0, g_live);
    ++g_launches;
    return;
  }
  const unsigned grid = (unsigned)ntiles;
  if (al16)
    blockmul_kernel<NQT, true><<<grid, BM_THREADS, smem, st>>>(n, V, ldv, p, C, ldc, q, alpha, beta, Y, ldy, g_live);
  else
    blockmul_kernel<NQT, false><<<grid, BM_THREADS, smem, st>>>(n, V, ldv, p, C, ldc, q, alpha, beta, Y, ldy, g_live);
  ++g_launches;
}

}  // namespace

void block_mul(cudaStream_t st, int64_t n, const double* V, int64_t ldv, int p, const double* C, int ldc, int q,
               double alpha, double beta, double* Y, int64_t ldy, bool upper_tri, int ident_from, bool ident_is_tri) {
  if (n <= 0 || q <= 0) return;
  if (q > 128) {
    // column blocks of Y are produced one launch at a time: Y must not overlap the columns of V
    const double* v0 = V;
    const double* v1 = V + (int64_t)(p - 1) * ldv + n;
    const double* y0 = Y;
    const double* y1 = Y + (int64_t)(q - 1) * ldy + n;
    if (y0 < v1 && v0 < y1) {
      std::fprintf(stderr, "diaglib_b200: block_mul called in place with q = %d > 128 columns\n", q);
      std::abort();
    }
  }
  for (int q0 = 0; q0 < q; q0 += 128) {
    const int qb = std::min(128, q - q0);
    const double* Cb = C + (size_t)q0 * ldc;
    double* Yb = Y + (int64_t)q0 * ldy;
    // 1: C upper triangular (first column block only); >= 2: rows >= mode - 2 of C are an identity block
    // (single column block only: the diagonal of the identity must start at column 0)
    // <= -2: rows >= -mode - 2 of C are an upper triangular block (same restriction)
    const int mode = (upper_tri && q0 == 0) ? 1 : ((ident_from >= 0 && q <= 128) ? (ident_is_tri ? -(2 + ident_from) : 2 + ident_from) : 0);
    if (qb <= 40)
      launch_blockmul<5>(st, n, V, ldv, p, Cb, ldc, qb, alpha, beta, Yb, ldy, mode);
    else if (qb <= 80)
      launch_blockmul<10>(st, n, V, ldv, p, Cb, ldc, qb, alpha, beta, Yb, ldy, mode);
    else
      launch_blockmul<16>(st, n, V, ldv, p, Cb, ldc, qb, alpha, beta, Yb, ldy, mode);
  }
  DLB_CUDA_CHECK(cudaGetLastError());
}

// U(n x m) <- U * T with T (m x m, ld m) upper triangular with explicit zeros below the
// diagonal (T = L^-T): the dtrmm('r','l','t','n') of ortho_cd (diaglib.f90:3327).  In place:
// column blocks are processed last-to-first so that a block only reads columns that have
// not been overwritten yet (U_new(:,j) depends on U(:,0..j) only).
void block_trmm_inplace(cudaStream_t st, int64_t n, double* U, int64_t ldu, int m, const double* T) {
  if (m <= 128) {
    block_mul(st, n, U, ldu, m, T, m, m, 1.0, 0.0, U, ldu, true);
    return;
  }
  const int nblk = (m + 127) / 128;
  for (int b = nblk - 1; b >= 0; --b) {
    const int q0 = b * 128, qb = std::min(128, m - q0);
    block_mul(st, n, U, ldu, q0 + qb, T + (size_t)q0 * m, m, qb, 1.0, 0.0, U + (int64_t)q0 * ldu, ldu, false);
  }
}

// Y = alpha V C + beta Y fused with the metric of the result, G (q x q, ldg) = Y^T Y (both
// triangles written).  Falls back to block_mul + gram_tn when the fused kernel does not apply
// (q > 40, unaligned data).  Used by ortho_cd / ortho_vs_x: every metric after the first one of an
// ortho_vs_x call comes out of the kernel that produced the block.
void block_mul_gram(cudaStream_t st, int num_sms, int64_t n, const double* V, int64_t ldv, int p, const double* C,
                    int ldc, int q, double alpha, double beta, double* Y, int64_t ldy, bool upper_tri, double* G,
                    int ldg, double* partial) {
  const bool al16 = aligned16(V) && (ldv % 2 == 0);
  const int p16 = (p + 15) / 16 * 16;
  const int PS = p16 + 4;
  constexpr int NQT = 5, QB = 40;
  const size_t smem = (2 * BMW_STAGES + (size_t)QB * PS + (size_t)BMW_STAGES * BM_KC * (8 * 16 + 4)) * sizeof(double);
  const int64_t ntiles = (n + 127) / 128;
  if (q <= QB && al16 && (n % 2 == 0) && !g_disable_ws && !g_disable_fused_gram && smem <= 200 * 1024 && ntiles >= 2) {
    static bool attr = false;
    if (!attr) {
      DLB_CUDA_CHECK(cudaFuncSetAttribute(blockmul_ws_kernel<NQT, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr = true;
    }
    int occ = 1;
    DLB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, blockmul_ws_kernel<NQT, 8, true>, 9 * 32, smem));
    const unsigned grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)num_sms * std::max(1, occ));
    blockmul_ws_kernel<NQT, 8, true><<<grid, 9 * 32, smem, st>>>(n, V, ldv, p, C, ldc, q, alpha, beta, Y, ldy, PS,
                                                                  upper_tri ? 1 : 0, partial, g_live, g_dbg);
    ++g_launches;
    const int tot = q * q;
    gram_reduce_kernel<<<(tot + 31) / 32, GRED_SL * 32, 0, st>>>(partial, (int)grid, QB, QB, q, q, 1, G, ldg, nullptr, g_live);
    ++g_launches;
    DLB_CUDA_CHECK(cudaGetLastError());
    return;
  }
  if (upper_tri && V == Y) block_trmm_inplace(st, n, Y, ldy, q, C);
  else block_mul(st, n, V, ldv, p, C, ldc, q, alpha, beta, Y, ldy, upper_tri);
  gram_tn(st, num_sms, n, Y, ldy, q, Y, ldy, q, G, ldg, true, partial);
}

}  // namespace dlb
