// microbench.cu — fp64 pipe throughput probes for design decisions (DFMA, DMMA shapes, exp/sqrt, I2F).
#include <cuda_runtime.h>
#include <stdio.h>
#include <math.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void k_dfma(double* out, int iters) {
  double a[8], x = 1.0000001, y = 1e-9 + threadIdx.x * 1e-12;
  for (int i = 0; i < 8; ++i) a[i] = i + threadIdx.x;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
  double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma884(double* out, int iters) {
  double c[8][2], a = 1.0 + threadIdx.x * 1e-9, b = 1e-3;
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma1688(double* out, int iters) {
  double c[4][4], a[4], b[2];
  for (int i = 0; i < 4; ++i) { a[i] = 1.0 + threadIdx.x * 1e-9; for (int j = 0; j < 4; ++j) c[i][j] = i; }
  b[0] = b[1] = 1e-3;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
  double s = 0; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma16816(double* out, int iters) {
  double c[4][4], a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-9;
  for (int i = 0; i < 4; ++i) { b[i] = 1e-3; for (int j = 0; j < 4; ++j) c[i][j] = i; }
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  double s = 0; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
__global__ void k_func(double* out, int iters) {
  double x[4];
  for (int i = 0; i < 4; ++i) x[i] = 0.3 + 0.1 * i + threadIdx.x * 1e-6;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (OP == 0) x[i] = exp(-x[i]) + 0.1;
      if (OP == 1) x[i] = sqrt(x[i]) + 0.1;
      if (OP == 2) x[i] = rsqrt(x[i]) * 0.5 + 0.1;
      if (OP == 3) x[i] = (double)(__double2int_rn(x[i] * 1000.0) ^ it) * 1e-6 + 0.1;  // F2I + I2F
      if (OP == 4) x[i] = 1.0 / (x[i] + 1.0);
    }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x[0] + x[1] + x[2] + x[3];
}
template <typename F>
static float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
  double* out; CK(cudaMalloc(&out, 148 * 8 * 256 * sizeof(double)));
  const int G = 148 * 8, T = 256, IT = 20000;
  double nthr = (double)G * T;
  float ms;
  ms = timeit([&] { k_dfma<<<G, T>>>(out, IT); });
  printf("DFMA        : %.2f TFLOP/s\n", nthr * IT * 8 * 2 / ms / 1e9);
  ms = timeit([&] { k_dmma884<<<G, T>>>(out, IT); });
  printf("DMMA m8n8k4 : %.2f TFLOP/s\n", nthr / 32 * IT * 8 * 512.0 / ms / 1e9);
  ms = timeit([&] { k_dmma1688<<<G, T>>>(out, IT); });
  printf("DMMA m16n8k8: %.2f TFLOP/s\n", nthr / 32 * IT * 4 * 2048.0 / ms / 1e9);
  ms = timeit([&] { k_dmma16816<<<G, T>>>(out, IT); });
  printf("DMMA m16n8k16: %.2f TFLOP/s\n", nthr / 32 * IT * 4 * 4096.0 / ms / 1e9);
  const char* names[5] = {"exp(double)", "sqrt(double)", "rsqrt(double)", "F2I+I2F double", "1/x double"};
  const int IT2 = 2000;
  ms = timeit([&] { k_func<0><<<G, T>>>(out, IT2); }); printf("%-16s: %.1f Gop/s\n", names[0], nthr * IT2 * 4 / ms / 1e6);
  ms = timeit([&] { k_func<1><<<G, T>>>(out, IT2); }); printf("%-16s: %.1f Gop/s\n", names[1], nthr * IT2 * 4 / ms / 1e6);
  ms = timeit([&] { k_func<2><<<G, T>>>(out, IT2); }); printf("%-16s: %.1f Gop/s\n", names[2], nthr * IT2 * 4 / ms / 1e6);
  ms = timeit([&] { k_func<3><<<G, T>>>(out, IT2); }); printf("%-16s: %.1f Gop/s\n", names[3], nthr * IT2 * 4 / ms / 1e6);
  ms = timeit([&] { k_func<4><<<G, T>>>(out, IT2); }); printf("%-16s: %.1f Gop/s\n", names[4], nthr * IT2 * 4 / ms / 1e6);
  return 0;
}
