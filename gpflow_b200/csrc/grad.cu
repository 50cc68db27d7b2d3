// grad.cu — device backward pass of GPR.log_marginal_likelihood.
//
// The reference gets d(LML)/d(theta) from TensorFlow autodiff through gpflow/models/gpr.py:91-107 (driven by
// gpflow/optimizers/scipy.py:78-228 via training_loss_closure, models/training_mixins.py:43-78).  Here the adjoint is
// written out:
//     dLML/dK = G = 1/2 (alpha alpha^T - P K^-1),   alpha = K^-1 (Y - m),   K = kernel(X) + sigma_n^2 I
//     dLML/dtheta = sum_ij G_ij dK_ij/dtheta ,      dLML/dsigma_n^2 = tr G
// with K^-1 = L^-T L^-1 from the factor the forward pass leaves behind:
//   1. alpha  = L^-T beta            (beta^T = the extra rows of the factorisation; one trsm)
//   2. L^-1   in place (recursive block inversion [A 0; C D]^-1 = [A^-1 0; -D^-1 C A^-1, D^-1]; the 128x128 diagonal
//             blocks are the block inverses potrf already produced; two triangular x dense GEMMs per level)
//   3. K^-1   = L^-T L^-1, lower triangle (recursive: C11 = lauum(A11) + A21^T A21, C21 = A22^T A21, C22 = lauum(A22))
//   4. one K-build-shaped pass over the lower-triangle tiles that re-evaluates k and dk/ds per element (s = scaled
//      squared distance, by direct differences), forms G_ij on the fly from alpha and K^-1, and reduces
//      sum G (.) dK/dtheta per parameter: registers -> warp shuffles -> one atomicAdd per CTA and parameter.
// Steps 2-3 run on the DMMA GEMM with the triangular operand's zero k-range skipped (GPK_GEMM_A_LOWER): 2 N^3 / 3 flops.
// Covered kernels: a single stationary leaf (SquaredExponential, Matern12/32/52, Exponential) with a scalar or ARD
// lengthscale; the Python layer raises NotImplementedError for anything else.
#include "internal.cuh"

namespace gpk {

constexpr int GR_MAXD = 32;
struct GradKern {
  int type;            // GPK_K_RBF / MATERN12 / MATERN32 / MATERN52 / EXPONENTIAL
  int nd;              // active dims
  int ard;             // 0: scalar lengthscale (one gradient slot), 1: nd slots
  double variance;
  int dims[GR_MAXD];
  double inv_l[GR_MAXD];  // 1 / lengthscale_d
};

__device__ __forceinline__ void k_and_dkds(int type, double s, double var, double& k, double& dkds) {
  // s = scaled squared distance; k(s) and dk/ds as gpflow/kernels/stationaries.py:209-210,250-251,270-271,290-292,311-313
  // (the 1e-36 clip before the square root passes no gradient when active, like tf.maximum)
  if (type == GPK_K_RBF) {
    k = var * exp(-0.5 * s);
    dkds = -0.5 * k;
    return;
  }
  const bool clipped = !(s > 1e-36);
  const double r = sqrt(clipped ? 1e-36 : s);
  if (type == GPK_K_MATERN12) {
    k = var * exp(-r);
    dkds = clipped ? 0.0 : -k / (2.0 * r);
  } else if (type == GPK_K_EXPONENTIAL) {
    k = var * exp(-0.5 * r);
    dkds = clipped ? 0.0 : -k / (4.0 * r);
  } else if (type == GPK_K_MATERN32) {
    const double s3 = 1.7320508075688772935, e = exp(-s3 * r);
    k = var * (1.0 + s3 * r) * e;
    dkds = clipped ? 0.0 : -1.5 * var * e;
  } else {  // MATERN52
    const double s5 = 2.2360679774997896964, e = exp(-s5 * r);
    k = var * (1.0 + s5 * r + (5.0 / 3.0) * r * r) * e;
    dkds = clipped ? 0.0 : -(5.0 / 6.0) * var * (1.0 + s5 * r) * e;
  }
}

constexpr int GT = 64;  // tile edge

// gout: [0] d/dvariance, [1] d/dnoise_variance, [2 ...] d/dlengthscale (1 slot, or nd slots with ARD)
template <int ND>
__global__ void __launch_bounds__(256)
gpr_grad_kernel(GradKern gk, const double* __restrict__ X, int64_t N, int64_t ldx, const double* __restrict__ alpha,
                int P, const double* __restrict__ Kinv, int64_t ldk, double* __restrict__ gout) {
  __shared__ double xa[GT][ND + 1], xb[GT][ND + 1];
  __shared__ double red[8][ND + 2];
  // lower-triangular tile index -> (ti, tj), tj <= ti
  const int64_t t = blockIdx.x;
  int64_t ti = (int64_t)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while (ti * (ti + 1) / 2 > t) --ti;
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  const int64_t tj = t - ti * (ti + 1) / 2;
  const int tid = threadIdx.x, tr = tid >> 4, tc = tid & 15;
  const int nd = gk.nd;
  for (int e = tid; e < GT * ND; e += 256) {
    const int r = e / ND, d = e % ND;
    const int64_t ra = ti * GT + r, rb = tj * GT + r;
    const double sc = d < nd ? gk.inv_l[d] : 0.0;
    const int col = d < nd ? gk.dims[d] : 0;
    xa[r][d] = (ra < N && d < nd) ? X[ra * ldx + col] * sc : 0.0;
    xb[r][d] = (rb < N && d < nd) ? X[rb * ldx + col] * sc : 0.0;
  }
  __syncthreads();
  double gv = 0.0, gn = 0.0, gl[ND];
#pragma unroll
  for (int d = 0; d < ND; ++d) gl[d] = 0.0;
  double gls = 0.0;
#pragma unroll 1
  for (int a = 0; a < 4; ++a) {
    const int r = tr + 16 * a;
    const int64_t i = ti * GT + r;
#pragma unroll 1
    for (int b = 0; b < 4; ++b) {
      const int c = tc + 16 * b;
      const int64_t j = tj * GT + c;
      if (i >= N || j > i) continue;
      double s = 0.0;
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        const double df = xa[r][d] - xb[c][d];
        s = fma(df, df, s);
      }
      double k, dkds;
      k_and_dkds(gk.type, s, gk.variance, k, dkds);
      double aa = 0.0;
      for (int p = 0; p < P; ++p) aa = fma(alpha[i * P + p], alpha[j * P + p], aa);
      const double G = 0.5 * (aa - (double)P * Kinv[i * ldk + j]);
      const double Ge = i == j ? G : 2.0 * G;  // the strict lower part stands for both (i,j) and (j,i)
      gv = fma(Ge, k, gv);
      if (i == j) gn += G;
      const double w = Ge * dkds * -2.0;
      if (gk.ard) {
#pragma unroll
        for (int d = 0; d < ND; ++d) {
          const double df = xa[r][d] - xb[c][d];
          gl[d] = fma(w, df * df, gl[d]);   // ds/dl_d = -2 diff_d^2 / l_d^3; the 1/l_d factor is applied at the end
        }
      } else {
        gls = fma(w, s, gls);               // ds/dl = -2 s / l
      }
    }
  }
  // CTA reduction: shuffles, then one atomicAdd per parameter
  const int lane = tid & 31, wp = tid >> 5;
  gv = warp_sum(gv);
  gn = warp_sum(gn);
  gls = warp_sum(gls);
#pragma unroll
  for (int d = 0; d < ND; ++d) gl[d] = warp_sum(gl[d]);
  if (lane == 0) {
    red[wp][0] = gv;
    red[wp][1] = gn;
#pragma unroll
    for (int d = 0; d < ND; ++d) red[wp][2 + d] = gk.ard ? gl[d] : (d == 0 ? gls : 0.0);
  }
  __syncthreads();
  if (tid < ND + 2) {
    double v = 0.0;
    for (int w2 = 0; w2 < 8; ++w2) v += red[w2][tid];
    if (tid == 0) atomicAdd(gout + 0, v / gk.variance);
    else if (tid == 1) atomicAdd(gout + 1, v);
    else {
      const int d = tid - 2;
      if (gk.ard) { if (d < nd) atomicAdd(gout + 2 + d, v * gk.inv_l[d]); }
      else if (d == 0) atomicAdd(gout + 2, v * gk.inv_l[0]);
    }
  }
}

// ---- L^-1 in place (lower), diagonal 128-blocks taken from the block inverses of the factorisation ------------
__global__ void put_dinv_kernel(double* __restrict__ L, int64_t ldl, int64_t n, const double* __restrict__ dinv) {
  const int64_t b0 = (int64_t)blockIdx.x * NB;
  const double* src = dinv + (size_t)blockIdx.x * NB * NB;
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    const int r = e / NB, c = e % NB;
    // the strict upper part of the block is zeroed: lauum reads the block as a dense operand, and the workspace above
    // the diagonal tiles of the K-build is uninitialised (0 x NaN)
    if (b0 + r < n && b0 + c < n) L[(b0 + r) * ldl + b0 + c] = c <= r ? src[r * NB + c] : 0.0;
  }
}

static inline int64_t split128(int64_t n) { return ((n / NB + 1) / 2) * NB; }

static int trtri_rec(double* L, int64_t n, int64_t ldl, double* tmp, cudaStream_t st) {
  if (n <= NB) return 0;  // diagonal blocks are already inverses
  const int64_t n1 = split128(n), n2 = n - n1;
  GPK_TRY(trtri_rec(L, n1, ldl, tmp, st));
  GPK_TRY(trtri_rec(L + n1 * ldl + n1, n2, ldl, tmp, st));
  double* L21 = L + n1 * ldl;
  double* A22 = L + n1 * ldl + n1;
  // Tt [n1, n2] = L11inv^T L21^T   (= (L21 L11inv)^T), the zero k-range of the triangular operand skipped
  GPK_TRY(gemm_t<double>(1, 1, n1, n2, n1, 1.0, L, ldl, L21, ldl, 0.0, tmp, n2, GPK_GEMM_A_LOWER, st));
  // L21 <- - L22inv (Tt)^T
  GPK_TRY(gemm_t<double>(0, 1, n2, n1, n2, -1.0, A22, ldl, tmp, n2, 0.0, L21, ldl, GPK_GEMM_A_LOWER, st));
  return 0;
}

// C (lower) = A^T A for lower-triangular A, out of place
static int lauum_rec(const double* A, int64_t n, int64_t lda, double* C, int64_t ldc, cudaStream_t st) {
  if (n <= NB)
    return gemm_t<double>(1, 0, n, n, n, 1.0, A, lda, A, lda, 0.0, C, ldc, GPK_GEMM_A_LOWER | GPK_GEMM_LOWER_ONLY, st);
  const int64_t n1 = split128(n), n2 = n - n1;
  const double* A21 = A + n1 * lda;
  const double* A22 = A + n1 * lda + n1;
  GPK_TRY(lauum_rec(A, n1, lda, C, ldc, st));
  GPK_TRY(lauum_rec(A22, n2, lda, C + n1 * ldc + n1, ldc, st));
  GPK_TRY(gemm_t<double>(1, 0, n1, n1, n2, 1.0, A21, lda, A21, lda, 1.0, C, ldc, GPK_GEMM_LOWER_ONLY, st));
  GPK_TRY(gemm_t<double>(1, 0, n2, n1, n2, 1.0, A22, lda, A21, lda, 0.0, C + n1 * ldc, ldc, GPK_GEMM_A_LOWER, st));
  return 0;
}

// K^-1 (lower triangle) from the factor L and its block inverses; L is overwritten by L^-1.
// tmp: (n/2 + 128)^2 doubles.
int potri_lower(double* L, int64_t n, int64_t ldl, const double* dinv, double* Kinv, int64_t ldk, double* tmp,
                cudaStream_t st) {
  const unsigned nblk = (unsigned)((n + NB - 1) / NB);
  put_dinv_kernel<<<nblk, 256, 0, st>>>(L, ldl, n, dinv);
  GPK_LAUNCH_OK();
  GPK_TRY(trtri_rec(L, n, ldl, tmp, st));
  return lauum_rec(L, n, ldl, Kinv, ldk, st);
}

int gpr_grad_launch(const gpk_knode* nodes, int n_nodes, const int32_t* dims, const double* ard, const double* X,
                    int64_t N, int64_t ldx, int64_t D, const double* alpha, int P, const double* Kinv, int64_t ldk,
                    double* gout, cudaStream_t st) {
  GPK_CHECK_ARG(n_nodes == 1, "gpr_lml_grad: the device backward covers a single stationary leaf kernel");
  const gpk_knode& nd = nodes[0];
  GPK_CHECK_ARG(nd.op == GPK_K_RBF || nd.op == GPK_K_MATERN12 || nd.op == GPK_K_MATERN32 || nd.op == GPK_K_MATERN52 ||
                    nd.op == GPK_K_EXPONENTIAL,
                "gpr_lml_grad: kernel op %d has no device backward", nd.op);
  GradKern gk;
  memset(&gk, 0, sizeof(gk));
  gk.type = nd.op;
  gk.variance = nd.variance;
  gk.nd = nd.n_dims > 0 ? nd.n_dims : (int)D;
  GPK_CHECK_ARG(gk.nd <= GR_MAXD, "gpr_lml_grad: more than %d active dims", GR_MAXD);
  gk.ard = nd.n_ard > 0 ? 1 : 0;
  for (int d = 0; d < gk.nd; ++d) {
    gk.dims[d] = nd.n_dims > 0 ? dims[nd.dims_off + d] : d;
    gk.inv_l[d] = 1.0 / (nd.n_ard > 0 ? ard[nd.ard_off + d] : nd.lengthscale);
  }
  const int64_t nt = (N + GT - 1) / GT;
  const unsigned grid = (unsigned)(nt * (nt + 1) / 2);
  ProfScope ps(PROF_KBUILD, st);
  if (gk.nd <= 8) gpr_grad_kernel<8><<<grid, 256, 0, st>>>(gk, X, N, ldx, alpha, P, Kinv, ldk, gout);
  else if (gk.nd <= 16) gpr_grad_kernel<16><<<grid, 256, 0, st>>>(gk, X, N, ldx, alpha, P, Kinv, ldk, gout);
  else gpr_grad_kernel<32><<<grid, 256, 0, st>>>(gk, X, N, ldx, alpha, P, Kinv, ldk, gout);
  GPK_LAUNCH_OK();
  return 0;
}

}  // namespace gpk
