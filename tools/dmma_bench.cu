// Microbenchmark: FP64 tensor pipe (DMMA) issue rate on sm_100a for the mma.sync f64 shapes,
// versus plain DFMA.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_bench dmma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void k884(double* out, int iters) {
  double c[NACC][2];
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k16816(double* out, int iters) {
  double c[NACC][4];
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; ++i) b[i] = threadIdx.x * 2e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void kfma(double* out, int iters) {
  double c[NACC];
  for (int i = 0; i < NACC; ++i) c[i] = i;
  double a = threadIdx.x * 1e-3 + 1.0, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed mode: does the FP64 tensor pipe (DMMA) run beside the plain FP64 pipe (DFMA)?  Every warp
// issues NM independent DMMAs and NF independent DFMAs per iteration; if the two were separate
// pipes the combined rate would exceed either peak.  (round-2 question of the judge.)
template <int NM, int NF>
__global__ void kmix(double* out, int iters) {
  double c[NM][2], f[NF];
  for (int i = 0; i < NM; ++i) c[i][0] = c[i][1] = 0.0;
  for (int i = 0; i < NF; ++i) f[i] = i;
  double a = threadIdx.x * 1e-3 + 1.0, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < (NM > NF ? NM : NF); ++i) {
      if (i < NM)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
      if (i < NF) f[i] = fma(f[i], a, b);
    }
  }
  double s = 0;
  for (int i = 0; i < NM; ++i) s += c[i][0] + c[i][1];
  for (int i = 0; i < NF; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// warp-specialised mix: even warps only DMMA, odd warps only DFMA
__global__ void ksplit(double* out, int iters) {
  double c[8][2], f[16];
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  for (int i = 0; i < 16; ++i) f[i] = i;
  double a = threadIdx.x * 1e-3 + 1.0, b = 1e-9;
  if ((threadIdx.x >> 5) & 1) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = fma(f[i], a, b);
    }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  for (int i = 0; i < 16; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
double timeit(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    int grid = sms * 2, threads = warps * 32 / 2;
    if (threads < 32) threads = 32;
    double ms = timeit([&] { k884<8><<<grid, threads>>>(out, iters); });
    double fl = (double)grid * (threads / 32) * iters * 8 * 512.0;
    printf("m8n8k4   acc=8 warps/SM=%2d : %.2f TFLOP/s\n", 2 * threads / 32, fl / ms / 1e9);
    ms = timeit([&] { k884<2><<<grid, threads>>>(out, iters); });
    fl = (double)grid * (threads / 32) * iters * 2 * 512.0;
    printf("m8n8k4   acc=2 warps/SM=%2d : %.2f TFLOP/s\n", 2 * threads / 32, fl / ms / 1e9);
    ms = timeit([&] { k884<1><<<grid, threads>>>(out, iters); });
    fl = (double)grid * (threads / 32) * iters * 1 * 512.0;
    printf("m8n8k4   acc=1 warps/SM=%2d : %.2f TFLOP/s\n", 2 * threads / 32, fl / ms / 1e9);
    ms = timeit([&] { k16816<4><<<grid, threads>>>(out, iters / 4); });
    fl = (double)grid * (threads / 32) * (iters / 4) * 4 * 4096.0;
    printf("m16n8k16 acc=4 warps/SM=%2d : %.2f TFLOP/s\n", 2 * threads / 32, fl / ms / 1e9);
    ms = timeit([&] { kfma<16><<<grid, threads>>>(out, iters); });
    fl = (double)grid * threads * iters * 16 * 2.0;
    printf("DFMA     acc=16 warps/SM=%2d : %.2f TFLOP/s\n", 2 * threads / 32, fl / ms / 1e9);
  }
  // mixed DMMA + DFMA
  for (int warps : {8, 16, 32}) {
    int grid = sms * 2, threads = warps * 32 / 2;
    double ms = timeit([&] { kmix<8, 16><<<grid, threads>>>(out, iters); });
    double fl_m = (double)grid * (threads / 32) * iters * 8 * 512.0, fl_f = (double)grid * threads * iters * 16 * 2.0;
    printf("MIX 8 DMMA + 16 DFMA per warp-iteration, warps/SM=%2d : %.2f TFLOP/s total (DMMA %.2f + DFMA %.2f)\n",
           2 * threads / 32, (fl_m + fl_f) / ms / 1e9, fl_m / ms / 1e9, fl_f / ms / 1e9);
    ms = timeit([&] { kmix<8, 4><<<grid, threads>>>(out, iters); });
    fl_f = (double)grid * threads * iters * 4 * 2.0;
    printf("MIX 8 DMMA +  4 DFMA per warp-iteration, warps/SM=%2d : %.2f TFLOP/s total (DMMA %.2f + DFMA %.2f)\n",
           2 * threads / 32, (fl_m + fl_f) / ms / 1e9, fl_m / ms / 1e9, fl_f / ms / 1e9);
    ms = timeit([&] { ksplit<<<grid, threads>>>(out, iters); });
    fl_m = (double)grid * (threads / 64) * iters * 8 * 512.0;
    fl_f = (double)grid * (threads / 2) * iters * 16 * 2.0;
    printf("SPLIT even warps DMMA / odd warps DFMA, warps/SM=%2d : %.2f TFLOP/s total (DMMA %.2f + DFMA %.2f)\n",
           2 * threads / 32, (fl_m + fl_f) / ms / 1e9, fl_m / ms / 1e9, fl_f / ms / 1e9);
  }
  return 0;
}
