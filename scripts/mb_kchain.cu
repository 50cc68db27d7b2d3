// Microbenchmarks behind the K-build design: fp64 / XU dependent-issue latencies and the throughput of the staged
// Matern52 evaluation on register data (no memory traffic), W elements per stage, with F2F or integer conversions.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/mb_kchain.bin scripts/mb_kchain.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void k_lat(double* out, long long* cyc, int mode) {
  double x = out[threadIdx.x] + 1.0, y = 1.0000001;
  float f = (float)x;
  long long t0 = clock64();
  if (mode == 0) {
#pragma unroll
    for (int i = 0; i < 256; ++i) x = fma(x, y, 1e-9);
  } else if (mode == 1) {
#pragma unroll
    for (int i = 0; i < 256; ++i) { f = (float)x; x = (double)f; }
  } else if (mode == 2) {
#pragma unroll
    for (int i = 0; i < 256; ++i) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(f));
  } else {
#pragma unroll
    for (int i = 0; i < 256; ++i) x = x * y;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x + f;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__device__ __forceinline__ float rsqrt_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int W, bool ICVT>
__device__ __forceinline__ void m52(const double (&xin)[W], double var, double var3, const double* tab, double (&out)[W]) {
  double x[W], g[W], h[W], r[W], d[W], u[W], pre[W];
#pragma unroll
  for (int q = 0; q < W; ++q) x[q] = __double2hiint(xin[q]) < 0x389a95a5 ? 5e-36 : xin[q];
#pragma unroll
  for (int q = 0; q < W; ++q) {
    if (ICVT) {
      const int hi = __double2hiint(x[q]) - 0x38000000;
      const unsigned fb = __funnelshift_l((unsigned)__double2loint(x[q]), (unsigned)hi, 3);
      const unsigned yb = __float_as_uint(rsqrt_approx(__uint_as_float(fb)));
      h[q] = __hiloint2double((int)(yb >> 3) + 0x37F00000, (int)(yb << 29));  // y0 / 2
    } else {
      h[q] = 0.5 * (double)rsqrt_approx((float)x[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < W; ++q) g[q] = (x[q] + x[q]) * h[q];
#pragma unroll
  for (int q = 0; q < W; ++q) r[q] = fma(-g[q], h[q], 0.5);
#pragma unroll
  for (int q = 0; q < W; ++q) g[q] = fma(g[q], r[q], g[q]);
#pragma unroll
  for (int q = 0; q < W; ++q) d[q] = fma(-g[q], g[q], x[q]);
#pragma unroll
  for (int q = 0; q < W; ++q) u[q] = fma(h[q], d[q], g[q]);
#pragma unroll
  for (int q = 0; q < W; ++q) pre[q] = fma(var3, x[q], fma(var, u[q], var));
  double sh[W], kd[W], rr[W], p[W];
  int k[W];
#pragma unroll
  for (int q = 0; q < W; ++q) sh[q] = fma(u[q], -92.33248261689366, 6755399441055744.0);
#pragma unroll
  for (int q = 0; q < W; ++q) { kd[q] = sh[q] - 6755399441055744.0; k[q] = __double2loint(sh[q]); }
#pragma unroll
  for (int q = 0; q < W; ++q) rr[q] = fma(kd[q], -0.01083042468962958, -u[q]);
#pragma unroll
  for (int q = 0; q < W; ++q) rr[q] = fma(kd[q], -6.619564634077006e-12, rr[q]);
#pragma unroll
  for (int q = 0; q < W; ++q) p[q] = fma(rr[q], 8.3333333333333332e-03, 4.1666666666666664e-02);
#pragma unroll
  for (int q = 0; q < W; ++q) p[q] = fma(p[q], rr[q], 1.6666666666666666e-01);
#pragma unroll
  for (int q = 0; q < W; ++q) p[q] = fma(p[q], rr[q], 0.5);
#pragma unroll
  for (int q = 0; q < W; ++q) p[q] = fma(p[q], rr[q], 1.0);
#pragma unroll
  for (int q = 0; q < W; ++q) p[q] = fma(p[q], rr[q], 1.0);
#pragma unroll
  for (int q = 0; q < W; ++q) {
    const double w = pre[q] * tab[k[q] & 63] * p[q];
    out[q] = __hiloint2double(__double2hiint(w) + ((k[q] >> 6) << 20), __double2loint(w));
  }
}

template <int W, bool ICVT>
__global__ void __launch_bounds__(256) k_m52(double* out, int iters, double var) {
  __shared__ double tab[64];
  if (threadIdx.x < 64) tab[threadIdx.x] = exp2(threadIdx.x / 64.0);
  __syncthreads();
  double x[W], o[W], acc = 0.0;
#pragma unroll
  for (int q = 0; q < W; ++q) x[q] = 0.1 + 0.01 * threadIdx.x + q;
  for (int it = 0; it < iters; ++it) {
    m52<W, ICVT>(x, var, var / 3, tab, o);
#pragma unroll
    for (int q = 0; q < W; ++q) { acc += o[q]; x[q] += 1e-3; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class F> float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
  double* out; long long* cyc;
  CK(cudaMalloc(&out, 148 * 8 * 256 * sizeof(double))); CK(cudaMemset(out, 0, 148 * 8 * 256 * sizeof(double)));
  CK(cudaMallocManaged(&cyc, 64));
  const char* nm[4] = {"DFMA dependent", "F2F d->f->d pair", "MUFU.RSQ dependent", "DMUL dependent"};
  for (int m = 0; m < 4; ++m) {
    k_lat<<<1, 32>>>(out, cyc, m); CK(cudaDeviceSynchronize());
    k_lat<<<1, 32>>>(out, cyc, m); CK(cudaDeviceSynchronize());
    printf("%-20s: %.1f cycles/op\n", nm[m], cyc[0] / 256.0);
  }
  const int IT = 2000;
  for (int per_sm = 1; per_sm <= 4; ++per_sm) {
    const int G = 148 * per_sm;
    const double n = (double)G * 256 * IT;
    float ms;
    ms = timeit([&] { k_m52<4, false><<<G, 256>>>(out, IT, 1.0); }); printf("CTAs/SM %d W=4 f2f : %.1f Gelem/s\n", per_sm, n * 4 / ms / 1e6);
    ms = timeit([&] { k_m52<4, true><<<G, 256>>>(out, IT, 1.0); });  printf("CTAs/SM %d W=4 icvt: %.1f Gelem/s\n", per_sm, n * 4 / ms / 1e6);
    ms = timeit([&] { k_m52<8, false><<<G, 256>>>(out, IT, 1.0); }); printf("CTAs/SM %d W=8 f2f : %.1f Gelem/s\n", per_sm, n * 8 / ms / 1e6);
    ms = timeit([&] { k_m52<8, true><<<G, 256>>>(out, IT, 1.0); });  printf("CTAs/SM %d W=8 icvt: %.1f Gelem/s\n", per_sm, n * 8 / ms / 1e6);
  }
  printf("floor at 21 fp64 ops/element (+2 bookkeeping): 18.4e12 / 23 = %.1f Gelem/s\n", 18.4e3 / 23);
  return 0;
}
