// Microbenchmark: how fast can 1-D bulk async copies (cp.async.bulk, UBLKCP) stream HBM into shared
// memory?  One producer lane per copy, a consumer warp that only waits.  Varies bytes per copy,
// copies per stage, stages and CTAs per SM.   nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mb_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t ph) {
  asm volatile("{ .reg .pred P1; W: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1; @P1 bra D; bra W; D: }" ::"r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(void* d, const void* s, uint32_t n, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(s), "r"(n), "r"(s32(b)) : "memory");
}
constexpr int STAGES = 4;
// each stage: ncopy copies of `bytes` bytes, source = ncopy different "columns" (stride col_stride) at a moving row offset
__global__ void stream(const char* src, size_t col_stride, int ncopy, int bytes, long chunks_total, double* sink) {
  extern __shared__ __align__(128) char smem[];
  uint64_t* full = (uint64_t*)smem;
  uint64_t* empty = full + STAGES;
  char* ring = smem + 128;
  const int stage_bytes = ncopy * bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int s = 0; s < STAGES; ++s) { mb_init(&full[s], 1); mb_init(&empty[s], 1); } asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  long my = (chunks_total - blockIdx.x + gridDim.x - 1) / gridDim.x;
  if (warp == 0) {
    int s = 0; uint32_t ph = 0;
    for (long it = 0; it < my; ++it) {
      long chunk = blockIdx.x + it * gridDim.x;
      mb_wait(&empty[s], ph ^ 1);
      if (lane == 0) mb_expect(&full[s], stage_bytes);
      __syncwarp();
      for (int c = lane; c < ncopy; c += 32) bulk(ring + (size_t)s * stage_bytes + c * bytes, src + c * col_stride + chunk * bytes, bytes, &full[s]);
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    int s = 0; uint32_t ph = 0; double acc = 0;
    for (long it = 0; it < my; ++it) {
      mb_wait(&full[s], ph);
      acc += ((double*)(ring + (size_t)s * stage_bytes))[lane];
      __syncwarp();
      if (lane == 0) mb_arrive(&empty[s]);
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
    if (acc == 12345.678) sink[0] = acc;
  }
}
int main() {
  const size_t rows = 1 << 24;           // doubles per column
  const int maxcols = 64;
  char* buf; cudaMalloc(&buf, rows * 8 * maxcols); cudaMemset(buf, 0, rows * 8 * maxcols);
  double* sink; cudaMalloc(&sink, 8);
  cudaFuncSetAttribute(stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int bytes : {256, 512, 1024, 2048, 4096}) for (int ncopy : {16, 40}) for (int cps : {1, 2, 4}) {
    size_t smem = 128 + (size_t)STAGES * ncopy * bytes;
    if (smem * cps > 220 * 1024) continue;
    long chunks = rows * 8 / bytes;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      stream<<<sms * cps, 64, smem>>>(buf, rows * 8, ncopy, bytes, chunks, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("bytes/copy %5d copies/stage %2d CTAs/SM %d : %7.1f GB/s  (%s)\n", bytes, ncopy, cps, (double)ncopy * rows * 8 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
