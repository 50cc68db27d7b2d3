// potrf.cu — blocked Cholesky and triangular solves for sm_100a.
//
// Replaces tf.linalg.cholesky (gpflow/models/gpr.py:102, posteriors.py:422,533,538,703,
// models/sgpr.py:201,207, conditionals/util.py:67, kullback_leiblers.py:107) and
// tf.linalg.triangular_solve (logdensities.py:150, conditionals/util.py:125,139, sgpr.py:204,264,
// posteriors.py:495-496,534,540,707,710, kullback_leiblers.py:114,152).
//
// Structure (row-major, lower): recursive blocked factorisation whose leaves are 128x128 diagonal
// blocks handled by ONE CTA entirely in shared memory (warp-cooperative 32x32 register Cholesky,
// per-row panel solves, 4x4 register-tiled updates) which also emits the INVERSE of the diagonal
// block; every off-diagonal operation — panel solve X = B L_jj^-T, trailing update C -= A A^T,
// forward/back substitution blocks — is then a dense GEMM (gemm.cu / gemm_tc.cu) with K >= 128.
// Rows below the square part (`rows > n`) ride along, so appending (Y-m)^T as extra rows yields
// alpha^T = (L^-1 (Y-m))^T without a separate TRSV (logdensities.py:150).
#include "common.cuh"

namespace gpk {

constexpr int LS = NB + 1;  // shared row stride of the leaf matrix

template <typename T> __device__ __forceinline__ T sqrt_t(T x);
template <> __device__ __forceinline__ double sqrt_t<double>(double x) { return sqrt(x); }
template <> __device__ __forceinline__ float sqrt_t<float>(float x) { return sqrtf(x); }

// ---- 32x32 diagonal block Cholesky by one warp: lane i owns row i in registers -------------------
template <typename T>
__device__ __forceinline__ int warp_chol32(T* S, int jb, T* dinvdiag) {
  const int lane = threadIdx.x & 31;
  T a[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) a[c] = c <= lane ? S[(jb + lane) * LS + jb + c] : T(0);
  int bad = 0;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    T d = __shfl_sync(0xffffffffu, a[k], k);
    if (!(d > T(0))) {  // non-positive or NaN pivot
      if (bad == 0) bad = k + 1;
      d = T(1);
    }
    const T piv = sqrt_t<T>(d);
    const T inv = T(1) / piv;
    a[k] = lane == k ? piv : a[k] * inv;  // l_ik for lane i >= k (rows above hold zeros)
    if (lane == k) dinvdiag[jb + k] = inv;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j > k) {
        const T ljk = __shfl_sync(0xffffffffu, a[k], j);
        if (lane >= j) a[j] -= a[k] * ljk;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 32; ++c)
    if (c <= lane) S[(jb + lane) * LS + jb + c] = a[c];
  return bad;
}

// ---- inverse of the 128x128 lower factor held in S (lower incl. diagonal) ---------------------------
// Result: strict lower part of Linv stored TRANSPOSED in the strict upper triangle of S
// (Linv[r][c] = S[c][r], r > c); diagonal of Linv in dinvdiag.  tmp: 3 * 32*32 scratch.
template <typename T>
__device__ void invert_lower_128(T* S, T* dinvdiag, T* tmp) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // (a) diagonal 32x32 blocks: warp J, lane j solves L_JJ x = e_j
  if (warp < 4) {
    const int jb = warp * 32;
    T x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = i == lane ? T(1) : T(0);
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      x[k] *= dinvdiag[jb + k];
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i > k) x[i] -= S[(jb + i) * LS + jb + k] * x[k];
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i > lane) S[(jb + lane) * LS + jb + i] = x[i];  // Linv[jb+i][jb+lane] -> transposed slot
  }
  __syncthreads();
  auto linv = [&](int r, int c) -> T {  // r >= c
    return r == c ? dinvdiag[r] : S[c * LS + r];
  };
  // (b) off-diagonal blocks by block distance d:  Linv_IJ = -Linv_II * sum_{K=J}^{I-1} L_IK Linv_KJ
  for (int d = 1; d < 4; ++d) {
    const int npairs = 4 - d;
    const int pair = tid >> 6, sub = tid & 63;       // 64 threads per 32x32 block, 4x4 micro-tiles
    const int r0 = (sub >> 3) * 4, c0 = (sub & 7) * 4;
    const int J = pair, I = pair + d;
    T acc[4][4];
    if (pair < npairs) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
      for (int K = J; K < I; ++K) {
        // L_IK [32x32] times Linv_KJ [32x32] (lower-triangular when K == J)
        for (int k = 0; k < 32; ++k) {
          T av[4], bv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) av[i] = S[(I * 32 + r0 + i) * LS + K * 32 + k];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int rr = K * 32 + k, cc = J * 32 + c0 + j;
            bv[j] = rr >= cc ? linv(rr, cc) : T(0);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] += av[i] * bv[j];
        }
      }
      T* tp = tmp + pair * 1024;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) tp[(r0 + i) * 32 + c0 + j] = acc[i][j];
    }
    __syncthreads();
    if (pair < npairs) {
      const T* tp = tmp + pair * 1024;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
      for (int k = 0; k < 32; ++k) {
        T av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = I * 32 + r0 + i, cc = I * 32 + k;
          av[i] = rr >= cc ? linv(rr, cc) : T(0);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = tp[k * 32 + c0 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] += av[i] * bv[j];
      }
      // Linv[I*32+r][J*32+c] = -acc  ->  S[J*32+c][I*32+r]
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) S[(J * 32 + c0 + j) * LS + I * 32 + r0 + i] = -acc[i][j];
    }
    __syncthreads();
  }
}

template <typename T>
__device__ __forceinline__ void write_dinv(const T* S, const T* dinvdiag, T* __restrict__ dinv) {
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    const int r = e / NB, c = e % NB;
    dinv[e] = r > c ? S[c * LS + r] : (r == c ? dinvdiag[r] : T(0));
  }
}

// ---- leaf: factor + invert one diagonal block (n <= 128) ------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256, 1)
potrf_leaf_kernel(T* __restrict__ A, int64_t lda, int n, T* __restrict__ dinv, int32_t* info, int info_base) {
  extern __shared__ __align__(16) unsigned char leaf_smem[];
  T* S = reinterpret_cast<T*>(leaf_smem);  // [128][129]
  T* dinvdiag = S + NB * LS;               // [128]
  T* tmp = dinvdiag + NB;                  // [3][32*32]
  const int tid = threadIdx.x;

  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e / NB, c = e % NB;
    T v = T(0);
    if (r < n && c <= r) v = A[(int64_t)r * lda + c];
    else if (r >= n && c == r) v = T(1);
    S[r * LS + c] = v;
  }
  __syncthreads();

  for (int J = 0; J < 4; ++J) {
    const int jb = J * 32;
    if (tid < 32) {
      const int bad = warp_chol32<T>(S, jb, dinvdiag);
      if (bad && tid == 0 && info) atomicCAS(info, 0, info_base + jb + bad);
    }
    __syncthreads();
    const int t0 = jb + 32, nr = NB - t0;
    if (nr > 0) {
      // panel: rows t0..127 solve x L_JJ^T = b  (one thread per row, x in registers)
      if (tid < nr) {
        const int r = t0 + tid;
        T x[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) x[c] = S[r * LS + jb + c];
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          x[k] *= dinvdiag[jb + k];
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (c > k) x[c] -= x[k] * S[(jb + c) * LS + jb + k];
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) S[r * LS + jb + c] = x[c];
      }
      __syncthreads();
      // trailing update of the remaining lower blocks with 4x4 register micro-tiles
      const int nt = nr / 4;
      for (int e = tid; e < nt * nt; e += 256) {
        const int ti = e / nt, tj = e % nt;
        if (tj > ti) continue;
        const int r0 = t0 + ti * 4, c0 = t0 + tj * 4;
        T acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
        for (int k = 0; k < 32; ++k) {
          T av[4], bv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) av[i] = S[(r0 + i) * LS + jb + k];
#pragma unroll
          for (int j = 0; j < 4; ++j) bv[j] = S[(c0 + j) * LS + jb + k];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] += av[i] * bv[j];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c0 + j <= r0 + i) S[(r0 + i) * LS + c0 + j] -= acc[i][j];
      }
      __syncthreads();
    }
  }

  // L back to global (lower part of the first n rows)
  for (int e = tid; e < n * NB; e += 256) {
    const int r = e / NB, c = e % NB;
    if (c <= r) A[(int64_t)r * lda + c] = S[r * LS + c];
  }
  __syncthreads();
  invert_lower_128<T>(S, dinvdiag, tmp);
  write_dinv<T>(S, dinvdiag, dinv);
}

// ---- standalone inverse of the diagonal blocks of a given factor (for trsm without cached dinv) -----
template <typename T>
__global__ void __launch_bounds__(256, 1)
trtri_diag_kernel(const T* __restrict__ L, int64_t ldl, int64_t n, T* __restrict__ dinv) {
  extern __shared__ __align__(16) unsigned char leaf_smem[];
  T* S = reinterpret_cast<T*>(leaf_smem);
  T* dinvdiag = S + NB * LS;
  T* tmp = dinvdiag + NB;
  const int tid = threadIdx.x;
  const int64_t b0 = (int64_t)blockIdx.x * NB;
  const int nb = (int)min((int64_t)NB, n - b0);
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e / NB, c = e % NB;
    T v = T(0);
    if (r < nb && c <= r) v = L[(b0 + r) * ldl + b0 + c];
    else if (r >= nb && c == r) v = T(1);
    S[r * LS + c] = v;
  }
  __syncthreads();
  if (tid < NB) dinvdiag[tid] = T(1) / S[tid * LS + tid];
  __syncthreads();
  invert_lower_128<T>(S, dinvdiag, tmp);
  write_dinv<T>(S, dinvdiag, dinv + (size_t)blockIdx.x * NB * NB);
}

template <typename T>
static size_t leaf_smem_bytes() { return (size_t)(NB * LS + NB + 3 * 1024) * sizeof(T); }

template <typename T>
static int leaf_attr() {
  static bool done = false;
  if (!done) {
    GPK_CUDA_OK(cudaFuncSetAttribute(potrf_leaf_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)leaf_smem_bytes<T>()));
    GPK_CUDA_OK(cudaFuncSetAttribute(trtri_diag_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)leaf_smem_bytes<T>()));
    done = true;
  }
  return 0;
}

static inline int64_t split_point(int64_t n) { return ((n / NB + 1) / 2) * NB; }

template <typename T>
static int potrf_rec(T* A, int64_t n, int64_t rows, int64_t lda, int32_t* info, T* dinv, int64_t col0,
                     cudaStream_t st) {
  if (n <= NB) {
    T* dblk = dinv + (size_t)(col0 / NB) * NB * NB;
    {
      ProfScope ps(PROF_LEAF, st);
      potrf_leaf_kernel<T><<<1, 256, leaf_smem_bytes<T>(), st>>>(A, lda, (int)n, dblk, info, (int)col0);
      GPK_LAUNCH_OK();
    }
    if (rows > n)  // rows below: X = B L^-T = B Linv^T, in place (single column tile)
      GPK_TRY(gemm_t<T>(0, 1, rows - n, n, n, T(1), A + n * lda, lda, dblk, NB, T(0), A + n * lda, lda, 0, st));
    return 0;
  }
  const int64_t n1 = split_point(n);
  GPK_TRY(potrf_rec<T>(A, n1, rows, lda, info, dinv, col0, st));
  // trailing update: A[n1:rows, n1:n] -= A[n1:rows, :n1] A[n1:n, :n1]^T  (lower tiles only)
  GPK_TRY(gemm_t<T>(0, 1, rows - n1, n - n1, n1, T(-1), A + n1 * lda, lda, A + n1 * lda, lda, T(1),
                    A + n1 * lda + n1, lda, GPK_GEMM_LOWER_ONLY, st));
  return potrf_rec<T>(A + n1 * lda + n1, n - n1, rows - n1, lda, info, dinv, col0 + n1, st);
}

template <typename T>
int potrf_t(T* A, int64_t n, int64_t rows, int64_t lda, int32_t* info, T* dinv, cudaStream_t st) {
  if (n <= 0) return 0;
  GPK_TRY(leaf_attr<T>());
  if (info) GPK_CUDA_OK(cudaMemsetAsync(info, 0, sizeof(int32_t), st));
  return potrf_rec<T>(A, n, rows, lda, info, dinv, 0, st);
}

template <typename T>
int trtri_diag_t(const T* L, int64_t n, int64_t ldl, T* dinv, cudaStream_t st) {
  if (n <= 0) return 0;
  GPK_TRY(leaf_attr<T>());
  const unsigned nblk = (unsigned)((n + NB - 1) / NB);
  trtri_diag_kernel<T><<<nblk, 256, leaf_smem_bytes<T>(), st>>>(L, ldl, n, dinv);
  GPK_LAUNCH_OK();
  return 0;
}

template <typename T>
static int trsm_rec(int trans, const T* L, int64_t n, int64_t ldl, T* B, int64_t nrhs, int64_t ldb, const T* dinv,
                    int64_t col0, cudaStream_t st) {
  if (n <= NB) {
    const T* dblk = dinv + (size_t)(col0 / NB) * NB * NB;
    // B <- Linv B  or  Linv^T B, in place (single row tile)
    return gemm_t<T>(trans ? 1 : 0, 0, n, nrhs, n, T(1), dblk, NB, B, ldb, T(0), B, ldb, 0, st);
  }
  const int64_t n1 = split_point(n);
  const T* L21 = L + n1 * ldl;
  const T* L22 = L + n1 * ldl + n1;
  T* B2 = B + n1 * ldb;
  if (!trans) {
    GPK_TRY(trsm_rec<T>(0, L, n1, ldl, B, nrhs, ldb, dinv, col0, st));
    GPK_TRY(gemm_t<T>(0, 0, n - n1, nrhs, n1, T(-1), L21, ldl, B, ldb, T(1), B2, ldb, 0, st));
    return trsm_rec<T>(0, L22, n - n1, ldl, B2, nrhs, ldb, dinv, col0 + n1, st);
  }
  GPK_TRY(trsm_rec<T>(1, L22, n - n1, ldl, B2, nrhs, ldb, dinv, col0 + n1, st));
  GPK_TRY(gemm_t<T>(1, 0, n1, nrhs, n - n1, T(-1), L21, ldl, B2, ldb, T(1), B, ldb, 0, st));
  return trsm_rec<T>(1, L, n1, ldl, B, nrhs, ldb, dinv, col0, st);
}

template <typename T>
int trsm_t(int trans, const T* L, int64_t n, int64_t ldl, T* B, int64_t nrhs, int64_t ldb, const T* dinv,
           cudaStream_t st) {
  if (n <= 0 || nrhs <= 0) return 0;
  return trsm_rec<T>(trans, L, n, ldl, B, nrhs, ldb, dinv, 0, st);
}

template int potrf_t<float>(float*, int64_t, int64_t, int64_t, int32_t*, float*, cudaStream_t);
template int potrf_t<double>(double*, int64_t, int64_t, int64_t, int32_t*, double*, cudaStream_t);
template int trtri_diag_t<float>(const float*, int64_t, int64_t, float*, cudaStream_t);
template int trtri_diag_t<double>(const double*, int64_t, int64_t, double*, cudaStream_t);
template int trsm_t<float>(int, const float*, int64_t, int64_t, float*, int64_t, int64_t, const float*, cudaStream_t);
template int trsm_t<double>(int, const double*, int64_t, int64_t, double*, int64_t, int64_t, const double*,
                            cudaStream_t);

}  // namespace gpk
