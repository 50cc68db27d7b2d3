// kbuild.cu — fused pairwise covariance builder (K-build) for sm_100a.
//
// One pass over the output: each CTA owns a 64x64 tile of K, stages the (weighted) active columns
// of its X / X2 row blocks in shared memory, forms the Gram term with a 4x4 register micro-tile,
// turns it into the scaled squared distance with the norm expansion the reference uses
// (gpflow/utilities/ops.py:105-122, on inputs scaled as in kernels/stationaries.py:77-79), applies
// every leaf function (stationaries.py:209-313, statics.py:57-91, linears.py:60-68) and folds the
// Sum/Product tree (kernels/base.py:281-314) in registers, adds the diagonal shift
// (utilities/model_utils.py:33-38, covariances/kuus.py:33) and writes K once with 16-byte
// vector stores.  The reference materialises one [N,N2] temporary per elementwise op instead.
//
// Algorithmic HBM bytes per launch: T*(N*N2 + (N+N2)*D)  (GPK_LOWER: T*(N(N+1)/2 + N*D)).
#include <math.h>
#include <stdarg.h>

#include <vector>

#include "common.cuh"

namespace gpk {

constexpr int KB_MAXG = 4;      // gram groups (distinct (active_dims, weights) sets)
constexpr int KB_MAXL = 12;     // leaves
constexpr int KB_MAXDIMS = 256; // total active dims over groups
constexpr int KB_MAXOPS = 32;
constexpr int KB_OP_ADD = 0xFE, KB_OP_MUL = 0xFF;
constexpr int KB_TILE = 64;     // output tile edge
constexpr int KB_KC = 32;       // dims staged per chunk

struct KProg {
  int n_groups, n_leaves, n_ops, symmetric;
  int g_ndims[KB_MAXG], g_off[KB_MAXG], g_weighted[KB_MAXG];
  int l_type[KB_MAXL], l_group[KB_MAXL];
  double l_scale[KB_MAXL], l_var[KB_MAXL], l_alpha[KB_MAXL];
  unsigned char ops[KB_MAXOPS];
  short dims[KB_MAXDIMS];
  double w[KB_MAXDIMS];
};
static_assert(sizeof(KProg) < 4000, "KProg must fit the kernel parameter space");

// ---------------------------------------------------------------------------------------------
// host: flatten the reference-shaped node list into groups / leaves / postfix ops
// ---------------------------------------------------------------------------------------------
static bool is_leaf_op(int op) { return op >= GPK_K_RBF && op <= GPK_K_CONSTANT; }
static bool uses_gram(int op) { return op <= GPK_K_LINEAR; }

static int emit_ops(const gpk_knode* nodes, int idx, const std::vector<int>& leaf_of, KProg& p, int& depth,
                    int& max_depth) {
  const gpk_knode& nd = nodes[idx];
  if (is_leaf_op(nd.op)) {
    if (p.n_ops >= KB_MAXOPS) return -1;
    p.ops[p.n_ops++] = (unsigned char)leaf_of[idx];
    depth++;
    if (depth > max_depth) max_depth = depth;
    return 0;
  }
  if (nd.n_children < 1 || nd.n_children > GPK_MAX_CHILDREN) return -1;
  for (int c = 0; c < nd.n_children; ++c) {
    if (nd.child[c] < 0 || nd.child[c] >= idx) return -1;  // children precede parents
    if (emit_ops(nodes, nd.child[c], leaf_of, p, depth, max_depth)) return -1;
    if (c > 0) {
      if (p.n_ops >= KB_MAXOPS) return -1;
      p.ops[p.n_ops++] = nd.op == GPK_K_SUM ? KB_OP_ADD : KB_OP_MUL;
      depth--;
    }
  }
  return 0;
}

int compile_kprog(const gpk_knode* nodes, int n_nodes, const int32_t* dims, const double* ard, int64_t D,
                  KProg& p) {
  memset(&p, 0, sizeof(p));
  GPK_CHECK_ARG(nodes && n_nodes > 0, "kbuild: empty kernel expression");
  GPK_CHECK_ARG(D > 0 && D < 32768, "kbuild: bad input dimension D=%lld", (long long)D);
  std::vector<int> leaf_of(n_nodes, -1);
  int tot_dims = 0;
  for (int i = 0; i < n_nodes; ++i) {
    const gpk_knode& nd = nodes[i];
    GPK_CHECK_ARG(nd.op >= GPK_K_RBF && nd.op <= GPK_K_PRODUCT, "kbuild: unknown kernel op %d", nd.op);
    if (!is_leaf_op(nd.op)) continue;
    GPK_CHECK_ARG(p.n_leaves < KB_MAXL, "kbuild: more than %d leaf kernels", KB_MAXL);
    int l = p.n_leaves++;
    leaf_of[i] = l;
    p.l_type[l] = nd.op;
    p.l_var[l] = nd.variance;
    p.l_alpha[l] = nd.alpha;
    p.l_scale[l] = 1.0;
    p.l_group[l] = -1;
    if (!uses_gram(nd.op)) continue;
    int nd_dims = nd.n_dims > 0 ? nd.n_dims : (int)D;
    GPK_CHECK_ARG(nd.n_dims == 0 || dims != nullptr, "kbuild: active dims given without index array");
    GPK_CHECK_ARG(nd.n_ard == 0 || (nd.n_ard == nd_dims && ard != nullptr),
                  "kbuild: size of ARD parameter (%d) does not match active dims (%d)", nd.n_ard, nd_dims);
    // weights of this leaf's gram term
    std::vector<double> w;
    if (nd.n_ard > 0) {
      w.resize(nd_dims);
      for (int d = 0; d < nd_dims; ++d) {
        double a = ard[nd.ard_off + d];
        w[d] = nd.op == GPK_K_LINEAR ? a : 1.0 / a;  // stationary: X/l on BOTH sides (stationaries.py:77-79)
      }
      if (nd.op == GPK_K_LINEAR) p.l_var[l] = 1.0;
    } else if (nd.op != GPK_K_LINEAR) {
      p.l_scale[l] = 1.0 / (nd.lengthscale * nd.lengthscale);
    }
    // find or create the group
    int g = -1;
    for (int c = 0; c < p.n_groups && g < 0; ++c) {
      const int wmode = w.empty() ? 0 : (nd.op == GPK_K_LINEAR ? 1 : 2);
      if (p.g_ndims[c] != nd_dims || p.g_weighted[c] != wmode) continue;
      bool same = true;
      for (int d = 0; d < nd_dims && same; ++d) {
        int col = nd.n_dims > 0 ? dims[nd.dims_off + d] : d;
        same = p.dims[p.g_off[c] + d] == col && (w.empty() || p.w[p.g_off[c] + d] == w[d]);
      }
      if (same) g = c;
    }
    if (g < 0) {
      GPK_CHECK_ARG(p.n_groups < KB_MAXG, "kbuild: more than %d distinct (active_dims, ARD) groups", KB_MAXG);
      GPK_CHECK_ARG(tot_dims + nd_dims <= KB_MAXDIMS, "kbuild: more than %d active dims in total", KB_MAXDIMS);
      g = p.n_groups++;
      p.g_ndims[g] = nd_dims;
      p.g_off[g] = tot_dims;
      p.g_weighted[g] = w.empty() ? 0 : (nd.op == GPK_K_LINEAR ? 1 : 2);  // 1: A side only, 2: both sides
      for (int d = 0; d < nd_dims; ++d) {
        int col = nd.n_dims > 0 ? dims[nd.dims_off + d] : d;
        GPK_CHECK_ARG(col >= 0 && col < D, "kbuild: active dim %d out of range [0,%lld)", col, (long long)D);
        p.dims[tot_dims + d] = (short)col;
        p.w[tot_dims + d] = w.empty() ? 1.0 : w[d];
      }
      tot_dims += nd_dims;
    }
    p.l_group[l] = g;
  }
  int depth = 0, max_depth = 0;
  GPK_CHECK_ARG(emit_ops(nodes, n_nodes - 1, leaf_of, p, depth, max_depth) == 0,
                "kbuild: malformed or too large kernel expression (max %d postfix ops)", KB_MAXOPS);
  GPK_CHECK_ARG(max_depth <= 4, "kbuild: kernel expression nests deeper than the 4-entry evaluation stack");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// device
// ---------------------------------------------------------------------------------------------
template <typename T>
struct KMath;
template <>
struct KMath<double> {
  static __device__ __forceinline__ double exp_(double x) { return exp(x); }
  static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
  static __device__ __forceinline__ double pow_(double x, double y) { return pow(x, y); }
};
template <>
struct KMath<float> {
  static __device__ __forceinline__ float exp_(float x) { return expf(x); }
  static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
  static __device__ __forceinline__ float pow_(float x, float y) { return powf(x, y); }
};

// value of one leaf given the (weighted) gram term `dot` and the row / column norms
template <typename T>
__device__ __forceinline__ T leaf_value(int type, T dot, T na, T nb, T scale, T var, T alpha, bool on_diag) {
  using M = KMath<T>;
  if (type == GPK_K_LINEAR) return var * dot;
  if (type == GPK_K_CONSTANT) return var;
  if (type == GPK_K_WHITE) return on_diag ? var : T(0);
  T r2 = scale * (na + nb - T(2) * dot);  // ops.py:113-122 — may be slightly negative
  if (type == GPK_K_RBF) return var * M::exp_(T(-0.5) * r2);                 // stationaries.py:210
  if (type == GPK_K_RQ) return var * M::pow_(T(1) + T(0.5) * r2 / alpha, -alpha);  // :238
  T r = M::sqrt_(fmax(r2, T(1e-36)));                                         // :114
  if (type == GPK_K_MATERN52) {                                               // :311-313
    const T s5 = T(2.23606797749978969641);
    return var * (T(1) + s5 * r + T(5.0 / 3.0) * r * r) * M::exp_(-s5 * r);
  }
  if (type == GPK_K_MATERN32) {                                               // :290-292
    const T s3 = T(1.73205080756887729353);
    return var * (T(1) + s3 * r) * M::exp_(-s3 * r);
  }
  if (type == GPK_K_MATERN12) return var * M::exp_(-r);                       // :270-271
  return var * M::exp_(T(-0.5) * r);                                          // Exponential :250-251
}

// element e (= r*4+c) of a register-resident 4x4 tile without dynamic register indexing
template <typename T>
__device__ __forceinline__ T sel16(const T (&d)[4][4], int e) {
  T v = d[0][0];
#pragma unroll
  for (int q = 1; q < 16; ++q)
    if (e == q) v = d[q >> 2][q & 3];
  return v;
}

template <typename T, int NG>
__global__ void __launch_bounds__(256)
kbuild_kernel(const __grid_constant__ KProg prog, const T* __restrict__ X, int64_t N, int64_t ldx,
              const T* __restrict__ X2, int64_t N2, int64_t ldx2, T* __restrict__ K, int64_t ldk, int lower,
              T diag_scalar, const T* __restrict__ diag_vec, int vec_ok) {
  const int bx = blockIdx.x, by = blockIdx.y;
  if (lower && bx > by) return;
  __shared__ __align__(16) T sA[KB_KC][KB_TILE];
  __shared__ __align__(16) T sB[KB_KC][KB_TILE];
  __shared__ T sNa[NG][KB_TILE];
  __shared__ T sNb[NG][KB_TILE];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)by * KB_TILE, col0 = (int64_t)bx * KB_TILE;
  const bool sym = prog.symmetric != 0;
  const T* Xb = sym ? X : X2;
  const int64_t ldb = sym ? ldx : ldx2;

  T dots[NG][4][4];
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) dots[g][r][c] = T(0);

  if (tid < KB_TILE) {
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      sNa[g][tid] = T(0);
      sNb[g][tid] = T(0);
    }
  }

#pragma unroll
  for (int g = 0; g < NG; ++g) {
    if (g >= prog.n_groups) break;
    const int nd = prog.g_ndims[g], off = prog.g_off[g];
    for (int d0 = 0; d0 < nd; d0 += KB_KC) {
      const int kc = min(KB_KC, nd - d0);
      __syncthreads();
      // stage: sA[d][r] = w_d * X[row0+r, dims[d]],  sB[d][c] = X2[col0+c, dims[d]]
      for (int e = tid; e < kc * KB_TILE; e += 256) {
        const int d = e % kc, r = e / kc;  // consecutive threads walk one row's dims (same cache lines)
        const int col = prog.dims[off + d0 + d];
        const int64_t gr = row0 + r, gc = col0 + r;
        T a = gr < N ? X[gr * ldx + col] : T(0);
        T b = gc < N2 ? Xb[gc * ldb + col] : T(0);
        const T wv = T(prog.w[off + d0 + d]);
        sA[d][r] = a * wv;
        sB[d][r] = prog.g_weighted[g] == 2 ? b * wv : b;
      }
      __syncthreads();
      // squared norms of the staged (scaled) rows / columns: same expression on both sides
      if (tid < 2 * KB_TILE) {
        const int r = tid & (KB_TILE - 1);
        T acc = T(0);
        if (tid < KB_TILE) {
          for (int d = 0; d < kc; ++d) { const T a = sA[d][r]; acc = fma(a, a, acc); }
          sNa[g][r] += acc;
        } else {
          for (int d = 0; d < kc; ++d) { const T b = sB[d][r]; acc = fma(b, b, acc); }
          sNb[g][r] += acc;
        }
      }
      // gram micro-tile
      for (int d = 0; d < kc; ++d) {
        T a[4], b[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) a[r] = sA[d][ty * 4 + r];
#pragma unroll
        for (int c = 0; c < 4; ++c) b[c] = sB[d][tx * 4 + c];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) dots[g][r][c] = fma(a[r], b[c], dots[g][r][c]);
      }
    }
  }
  __syncthreads();

  // epilogue: leaf functions + postfix Sum/Product fold.  The 16 elements of the micro-tile are
  // processed by ONE rolled loop (values staged in a small local array): unrolling it replicates
  // the exp/sqrt/pow code 16x (180 KB of SASS) and makes the kernel instruction-fetch bound.
  T vals[16];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) vals[r * 4 + c] = NG == 1 ? dots[0][r][c] : T(0);
  const int n_ops = prog.n_ops;
#pragma unroll 1
  for (int e = 0; e < 16; ++e) {
    const int r = e >> 2, c = e & 3;
    const int64_t gi = row0 + ty * 4 + r, gj = col0 + tx * 4 + c;
    const bool on_diag = sym && gi == gj;
    T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0);
    for (int o = 0; o < n_ops; ++o) {
      const int op = prog.ops[o];
      if (op < KB_MAXL) {
        const int g = prog.l_group[op];
        T dot = T(0), na = T(0), nb = T(0);
        if (NG == 1) {
          dot = vals[e];
          if (g == 0) { na = sNa[0][ty * 4 + r]; nb = sNb[0][tx * 4 + c]; }
        } else {
#pragma unroll
          for (int gg = 0; gg < NG; ++gg)
            if (g == gg) {
              dot = sel16(dots[gg], e);
              na = sNa[gg][ty * 4 + r];
              nb = sNb[gg][tx * 4 + c];
            }
        }
        T v = leaf_value<T>(prog.l_type[op], dot, na, nb, T(prog.l_scale[op]), T(prog.l_var[op]),
                            T(prog.l_alpha[op]), on_diag);
        s3 = s2; s2 = s1; s1 = s0; s0 = v;
      } else {
        s0 = op == KB_OP_ADD ? s1 + s0 : s1 * s0;
        s1 = s2; s2 = s3;
      }
    }
    if (on_diag) s0 += diag_scalar + (diag_vec ? diag_vec[gi] : T(0));
    vals[e] = s0;
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t gi = row0 + ty * 4 + r;
    if (gi < N) {
      const int64_t gj0 = col0 + tx * 4;
      T* dst = K + gi * ldk + gj0;
      if (vec_ok && gj0 + 3 < N2) {
        if (sizeof(T) == 8) {
          reinterpret_cast<double2*>(dst)[0] = make_double2((double)vals[r * 4 + 0], (double)vals[r * 4 + 1]);
          reinterpret_cast<double2*>(dst)[1] = make_double2((double)vals[r * 4 + 2], (double)vals[r * 4 + 3]);
        } else {
          reinterpret_cast<float4*>(dst)[0] =
              make_float4((float)vals[r * 4 + 0], (float)vals[r * 4 + 1], (float)vals[r * 4 + 2], (float)vals[r * 4 + 3]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (gj0 + c < N2) dst[c] = vals[r * 4 + c];
      }
    }
  }
}

template <typename T>
__global__ void kdiag_kernel(const __grid_constant__ KProg prog, const T* __restrict__ X, int64_t N, int64_t ldx,
                             T* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0);
  for (int o = 0; o < prog.n_ops; ++o) {
    const int op = prog.ops[o];
    if (op < KB_MAXL) {
      T v = T(prog.l_var[op]);  // stationaries.py:82-83, statics.py:41-42
      if (prog.l_type[op] == GPK_K_LINEAR) {  // linears.py:67-68
        const int g = prog.l_group[op];
        T acc = T(0);
        for (int d = 0; d < prog.g_ndims[g]; ++d) {
          T x = X[i * ldx + prog.dims[prog.g_off[g] + d]];
          acc += T(prog.w[prog.g_off[g] + d]) * x * x;
        }
        v *= acc;
      }
      s3 = s2; s2 = s1; s1 = s0; s0 = v;
    } else {
      s0 = op == KB_OP_ADD ? s1 + s0 : s1 * s0;
      s1 = s2; s2 = s3;
    }
  }
  out[i] = s0;
}


// =============================================================================================
// Fast path: ONE stationary leaf (RBF / Matern12 / Matern32 / Matern52 / Exponential) — the common
// case and BASELINE config 2.  fp64 on B200 is compute-bound (exp(double) 835 Gop/s, sqrt 1164 Gop/s
// measured vs 18.3 T DFMA/s), so this path
//   * evaluates only lower-triangle tiles of a symmetric K and, for GPK_FULL, writes the mirrored
//     tile through a shared-memory transpose (compute once, store twice: HBM-bound);
//   * uses a table-driven exp (2^(j/64) table + degree-5 polynomial, ~12 DFMA instead of ~22) and a
//     MUFU-seeded Newton square root (~9 instead of ~16);
//   * launches a 1-D grid over the needed tiles only.
// =============================================================================================
__constant__ double c_exp2_tab[64];
static double h_exp2_tab[64];
static bool h_exp2_tab_ready = false;

__device__ __forceinline__ double fast_exp_neg(double x, const double* __restrict__ tab) {  // x <= 0, tab in SHARED memory
  if (x < -708.0) return 0.0;
  const double t = x * 92.33248261689366;  // 64 / ln 2
  const double sh = t + 6755399441055744.0;              // 1.5 * 2^52: round-to-nearest integer in the low bits
  const double kd = sh - 6755399441055744.0;
  const int k = __double2loint(sh);
  double r = fma(kd, -0.01083042468962958, x);           // ln2/64 hi (low 22 mantissa bits zero: kd*hi exact)
  r = fma(kd, -6.619564634077006e-12, r);                // ln2/64 lo
  double p = fma(r, 8.3333333333333332e-03, 4.1666666666666664e-02);
  p = fma(p, r, 1.6666666666666666e-01);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const int j = k & 63, n = k >> 6;                      // k = 64 n + j, floor semantics for negative k
  const double two_n = __longlong_as_double((long long)(n + 1023) << 52);
  return tab[j] * p * two_n;  // lanes hit different entries: shared memory, never __constant__
}
__device__ __forceinline__ double fast_sqrt_pos(double x) {  // x >= 1e-36
  double y = (double)__frsqrt_rn((float)x);  // MUFU.RSQ seed
  double e = fma(-x * y, y, 1.0);
  y = fma(0.5 * y, e, y);
  e = fma(-x * y, y, 1.0);
  y = fma(0.5 * y, e, y);
  double s = x * y;
  return fma(0.5 * y, fma(-s, s, x), s);
}

template <typename T> struct FastMath;
template <> struct FastMath<double> {
  static __device__ __forceinline__ double exp_neg(double x, const double* tab) { return fast_exp_neg(x, tab); }
  static __device__ __forceinline__ double sqrt_pos(double x) { return fast_sqrt_pos(x); }
};
template <> struct FastMath<float> {
  static __device__ __forceinline__ float exp_neg(float x, const double*) { return __expf(x); }
  static __device__ __forceinline__ float sqrt_pos(float x) { return sqrtf(x); }
};

template <typename T, int TYPE>
__device__ __forceinline__ T stationary_value(T r2, T var, const double* tab) {
  using F = FastMath<T>;
  if (TYPE == GPK_K_RBF) return var * F::exp_neg(fmin(T(-0.5) * r2, T(0)), tab);
  const T r2c = fmax(r2, T(1e-36));
  const T r = F::sqrt_pos(r2c);
  if (TYPE == GPK_K_MATERN52) {
    const T s5 = T(2.23606797749978969641);
    return var * fma(T(5.0 / 3.0), r2c, fma(s5, r, T(1))) * F::exp_neg(-s5 * r, tab);
  }
  if (TYPE == GPK_K_MATERN32) {
    const T s3 = T(1.73205080756887729353);
    return var * fma(s3, r, T(1)) * F::exp_neg(-s3 * r, tab);
  }
  if (TYPE == GPK_K_MATERN12) return var * F::exp_neg(-r, tab);
  return var * F::exp_neg(T(-0.5) * r, tab);  // Exponential
}

// Four elements at a time, stage by stage: every arithmetic step is issued for the four independent
// elements back to back, so the long dependent chains (sqrt: ~8, exp: ~12 fp64 ops) overlap instead of
// each waiting out the fp64 latency alone.
__device__ __forceinline__ float rsqrt_approx(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int TYPE>
__device__ __forceinline__ void stationary_value4(const double (&r2in)[4], double var, const double* __restrict__ tab,
                                                  double (&out)[4]) {
  double arg[4], pre[4];
  if (TYPE == GPK_K_RBF) {
#pragma unroll
    for (int q = 0; q < 4; ++q) { arg[q] = fmin(-0.5 * r2in[q], 0.0); pre[q] = var; }
  } else {
    double x[4], y[4], e[4], r[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) x[q] = fmax(r2in[q], 1e-36);
#pragma unroll
    for (int q = 0; q < 4; ++q) y[q] = (double)rsqrt_approx((float)x[q]);
    // y ~ x^-1/2 to 2^-22 (MUFU) -> one Newton step (2^-43) -> r = x*y and one Heron correction of r
    // with the refined y: relative error ~2^-85 before rounding
#pragma unroll
    for (int q = 0; q < 4; ++q) e[q] = fma(-x[q] * y[q], y[q], 1.0);
#pragma unroll
    for (int q = 0; q < 4; ++q) y[q] = fma(0.5 * y[q], e[q], y[q]);
#pragma unroll
    for (int q = 0; q < 4; ++q) r[q] = x[q] * y[q];
#pragma unroll
    for (int q = 0; q < 4; ++q) r[q] = fma(0.5 * y[q], fma(-r[q], r[q], x[q]), r[q]);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (TYPE == GPK_K_MATERN52) {
        const double s5 = 2.23606797749978969641;
        pre[q] = var * fma(5.0 / 3.0, x[q], fma(s5, r[q], 1.0));
        arg[q] = -s5 * r[q];
      } else if (TYPE == GPK_K_MATERN32) {
        const double s3 = 1.73205080756887729353;
        pre[q] = var * fma(s3, r[q], 1.0);
        arg[q] = -s3 * r[q];
      } else if (TYPE == GPK_K_MATERN12) {
        pre[q] = var;
        arg[q] = -r[q];
      } else {
        pre[q] = var;
        arg[q] = -0.5 * r[q];
      }
    }
  }
  // exp(arg), arg <= 0
  double sh[4], kd[4], rr[4], p[4];
  int k[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) sh[q] = fma(arg[q], 92.33248261689366, 6755399441055744.0);
#pragma unroll
  for (int q = 0; q < 4; ++q) { kd[q] = sh[q] - 6755399441055744.0; k[q] = __double2loint(sh[q]); }
#pragma unroll
  for (int q = 0; q < 4; ++q) rr[q] = fma(kd[q], -0.01083042468962958, arg[q]);
#pragma unroll
  for (int q = 0; q < 4; ++q) rr[q] = fma(kd[q], -6.619564634077006e-12, rr[q]);
#pragma unroll
  for (int q = 0; q < 4; ++q) p[q] = fma(rr[q], 8.3333333333333332e-03, 4.1666666666666664e-02);
#pragma unroll
  for (int q = 0; q < 4; ++q) p[q] = fma(p[q], rr[q], 1.6666666666666666e-01);
#pragma unroll
  for (int q = 0; q < 4; ++q) p[q] = fma(p[q], rr[q], 0.5);
#pragma unroll
  for (int q = 0; q < 4; ++q) p[q] = fma(p[q], rr[q], 1.0);
#pragma unroll
  for (int q = 0; q < 4; ++q) p[q] = fma(p[q], rr[q], 1.0);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const double two_n = __longlong_as_double((long long)((k[q] >> 6) + 1023) << 52);
    const double v = pre[q] * tab[k[q] & 63] * p[q] * two_n;
    out[q] = arg[q] < -708.0 ? 0.0 : v;
  }
}

constexpr int KF_TS = KB_TILE + 1;  // transpose staging stride

template <typename T, int TYPE>
__global__ void __launch_bounds__(256, 3)
kbuild_fast_kernel(const __grid_constant__ KProg prog, const T* __restrict__ X, int64_t N, int64_t ldx,
                   const T* __restrict__ X2, int64_t N2, int64_t ldx2, T* __restrict__ K, int64_t ldk, int mode,
                   T diag_scalar, const T* __restrict__ diag_vec, int vec_ok) {
  // mode: 0 rectangular (2-D tile index from the 1-D grid), 1 symmetric lower-only, 2 symmetric full (mirror)
  extern __shared__ __align__(16) unsigned char kf_smem[];
  T* sA = reinterpret_cast<T*>(kf_smem);            // [KB_KC][KB_TILE]
  T* sB = sA + KB_KC * KB_TILE;                     // [KB_KC][KB_TILE]
  T* sNa = sB + KB_KC * KB_TILE;                    // [KB_TILE]
  T* sNb = sNa + KB_TILE;                           // [KB_TILE]
  T* sT = sNb + KB_TILE;                            // [KB_TILE][KF_TS] transpose staging (mode 2)
  __shared__ double s_tab[64];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  if (tid < 64) s_tab[tid] = c_exp2_tab[tid];
  int64_t by, bx;
  if (mode == 0) {
    const int64_t ntx = (N2 + KB_TILE - 1) / KB_TILE;
    by = blockIdx.x / ntx;
    bx = blockIdx.x % ntx;
  } else {  // lower-triangle tile enumeration: t = by (by + 1) / 2 + bx, bx <= by
    const int64_t t = blockIdx.x;
    by = (int64_t)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while (by * (by + 1) / 2 > t) --by;
    while ((by + 1) * (by + 2) / 2 <= t) ++by;
    bx = t - by * (by + 1) / 2;
  }
  const int64_t row0 = by * KB_TILE, col0 = bx * KB_TILE;
  const bool sym = mode != 0;
  const T* Xb = sym ? X : X2;
  const int64_t ldb = sym ? ldx : ldx2;

  T dots[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) dots[r][c] = T(0);
  if (tid < KB_TILE) { sNa[tid] = T(0); sNb[tid] = T(0); }
  const int nd = prog.g_ndims[0];
  for (int d0 = 0; d0 < nd; d0 += KB_KC) {
    const int kc = min(KB_KC, nd - d0);
    __syncthreads();
    for (int e = tid; e < kc * KB_TILE; e += 256) {
      const int d = e % kc, r = e / kc;
      const int col = prog.dims[d0 + d];
      const int64_t gr = row0 + r, gc = col0 + r;
      const T a = gr < N ? X[gr * ldx + col] : T(0);
      const T b = gc < N2 ? Xb[gc * ldb + col] : T(0);
      const T wv = T(prog.w[d0 + d]);   // 1 or 1/lengthscale_d, applied to both sides
      sA[d * KB_TILE + r] = a * wv;
      sB[d * KB_TILE + r] = b * wv;
    }
    __syncthreads();
    if (tid < 2 * KB_TILE) {  // squared norms of the staged rows / columns, identical expression
      const int r = tid & (KB_TILE - 1);
      const T* src = tid < KB_TILE ? sA : sB;
      T acc = T(0);
      for (int d = 0; d < kc; ++d) { const T v = src[d * KB_TILE + r]; acc = fma(v, v, acc); }
      if (tid < KB_TILE) sNa[r] += acc; else sNb[r] += acc;
    }
    for (int d = 0; d < kc; ++d) {
      T a[4], b[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = sA[d * KB_TILE + ty * 4 + r];
#pragma unroll
      for (int c = 0; c < 4; ++c) b[c] = sB[d * KB_TILE + tx * 4 + c];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) dots[r][c] = fma(a[r], b[c], dots[r][c]);
    }
  }
  __syncthreads();

  const T scale = T(prog.l_scale[0]), var = T(prog.l_var[0]);
  T vals[16];
  T na[4], nb[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) { na[r] = sNa[ty * 4 + r]; nb[r] = sNb[tx * 4 + r]; }
  // fully unrolled (no local array, 32-bit index math); the sum of norms is formed FIRST so that
  // (i,j) and (j,i) round identically and a symmetric K comes out bit-symmetric
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    if (sizeof(T) == 8) {
      double r2[4], o[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) r2[c] = (double)scale * fma(-2.0, (double)dots[r][c], (double)na[r] + (double)nb[c]);
      stationary_value4<TYPE>(r2, (double)var, s_tab, o);
#pragma unroll
      for (int c = 0; c < 4; ++c) vals[r * 4 + c] = (T)o[c];
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const T sn = na[r] + nb[c];
        const T r2 = scale * fma(T(-2), dots[r][c], sn);
        vals[r * 4 + c] = stationary_value<T, TYPE>(r2, var, s_tab);
      }
    }
  }
  if (sym && bx == by) {  // diagonal shift: only diagonal tiles carry diagonal elements
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (ty * 4 + r == tx * 4 + c) {
          const int64_t gi = row0 + ty * 4 + r;
          vals[r * 4 + c] += diag_scalar + ((diag_vec && gi < N) ? diag_vec[gi] : T(0));
        }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t gi = row0 + ty * 4 + r;
    if (gi < N) {
      const int64_t gj0 = col0 + tx * 4;
      T* dst = K + gi * ldk + gj0;
      if (vec_ok && gj0 + 3 < N2) {
        if (sizeof(T) == 8) {
          reinterpret_cast<double2*>(dst)[0] = make_double2((double)vals[r * 4 + 0], (double)vals[r * 4 + 1]);
          reinterpret_cast<double2*>(dst)[1] = make_double2((double)vals[r * 4 + 2], (double)vals[r * 4 + 3]);
        } else {
          reinterpret_cast<float4*>(dst)[0] =
              make_float4((float)vals[r * 4 + 0], (float)vals[r * 4 + 1], (float)vals[r * 4 + 2], (float)vals[r * 4 + 3]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (gj0 + c < N2) dst[c] = vals[r * 4 + c];
      }
    }
  }
  if (mode == 2 && bx < by) {  // mirrored tile K[col0.., row0..] = tile^T, staged for coalesced rows
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) sT[(tx * 4 + c) * KF_TS + ty * 4 + r] = vals[r * 4 + c];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int tr = ty + 16 * r;                   // row of the transposed tile
      const int64_t gi = col0 + tr;
      if (gi < N2) {
        const int64_t gj0 = row0 + tx * 4;
        T* dst = K + gi * ldk + gj0;
        const T* src = sT + tr * KF_TS + tx * 4;
        if (vec_ok && gj0 + 3 < N) {
          if (sizeof(T) == 8) {
            reinterpret_cast<double2*>(dst)[0] = make_double2((double)src[0], (double)src[1]);
            reinterpret_cast<double2*>(dst)[1] = make_double2((double)src[2], (double)src[3]);
          } else {
            reinterpret_cast<float4*>(dst)[0] = make_float4((float)src[0], (float)src[1], (float)src[2], (float)src[3]);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (gj0 + c < N) dst[c] = src[c];
        }
      }
    }
  }
}

static bool fast_path_ok(const KProg& p) {
  if (p.n_leaves != 1 || p.n_ops != 1 || p.n_groups != 1) return false;
  const int t = p.l_type[0];
  return t == GPK_K_RBF || t == GPK_K_MATERN12 || t == GPK_K_MATERN32 || t == GPK_K_MATERN52 || t == GPK_K_EXPONENTIAL;
}

template <typename T, int TYPE>
static int kbuild_fast_go(const KProg& p, const void* X, int64_t N, int64_t ldx, const void* X2, int64_t N2, int64_t ldx2,
                          void* K, int64_t ldk, int mode, double diag_scalar, const void* diag_vec, int vec_ok,
                          cudaStream_t st) {
  const size_t smem = (size_t)(2 * KB_KC * KB_TILE + 2 * KB_TILE + KB_TILE * KF_TS) * sizeof(T);
  static bool attr = false;
  if (!attr) {
    GPK_CUDA_OK(cudaFuncSetAttribute(kbuild_fast_kernel<T, TYPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const int64_t nty = (N + KB_TILE - 1) / KB_TILE, ntx = (N2 + KB_TILE - 1) / KB_TILE;
  const int64_t ntiles = mode == 0 ? nty * ntx : nty * (nty + 1) / 2;
  GPK_CHECK_ARG(ntiles < (1ll << 31), "kbuild: too many tiles");
  kbuild_fast_kernel<T, TYPE><<<(unsigned)ntiles, 256, smem, st>>>(p, (const T*)X, N, ldx, (const T*)X2, N2, ldx2, (T*)K,
                                                                    ldk, mode, (T)diag_scalar, (const T*)diag_vec, vec_ok);
  GPK_LAUNCH_OK();
  return 0;
}

template <typename T>
static int kbuild_fast_launch(const KProg& p, const void* X, int64_t N, int64_t ldx, const void* X2, int64_t N2,
                              int64_t ldx2, void* K, int64_t ldk, int lower, double diag_scalar, const void* diag_vec,
                              cudaStream_t st) {
  if (!h_exp2_tab_ready) {
    for (int j = 0; j < 64; ++j) h_exp2_tab[j] = exp2((double)j / 64.0);
    GPK_CUDA_OK(cudaMemcpyToSymbol(c_exp2_tab, h_exp2_tab, sizeof(h_exp2_tab)));
    h_exp2_tab_ready = true;
  }
  ProfScope ps(PROF_KBUILD, st);
  const int vec_ok = ((uintptr_t)K % 16 == 0) && ((ldk * sizeof(T)) % 16 == 0);
  const int mode = p.symmetric ? (lower ? 1 : 2) : 0;
#define GPK_KF(TY) return kbuild_fast_go<T, TY>(p, X, N, ldx, X2, N2, ldx2, K, ldk, mode, diag_scalar, diag_vec, vec_ok, st)
  switch (p.l_type[0]) {
    case GPK_K_RBF: GPK_KF(GPK_K_RBF);
    case GPK_K_MATERN12: GPK_KF(GPK_K_MATERN12);
    case GPK_K_MATERN32: GPK_KF(GPK_K_MATERN32);
    case GPK_K_MATERN52: GPK_KF(GPK_K_MATERN52);
    default: GPK_KF(GPK_K_EXPONENTIAL);
  }
#undef GPK_KF
}

template <typename T>
static int kbuild_launch(const KProg& p, const void* X, int64_t N, int64_t ldx, const void* X2, int64_t N2,
                         int64_t ldx2, void* K, int64_t ldk, int lower, double diag_scalar, const void* diag_vec,
                         cudaStream_t st) {
  ProfScope ps(PROF_KBUILD, st);
  dim3 grid((unsigned)((N2 + KB_TILE - 1) / KB_TILE), (unsigned)((N + KB_TILE - 1) / KB_TILE));
  const int vec_ok = ((uintptr_t)K % 16 == 0) && ((ldk * sizeof(T)) % 16 == 0);
#define GPK_KB_GO(NG)                                                                                            \
  kbuild_kernel<T, NG><<<grid, 256, 0, st>>>(p, (const T*)X, N, ldx, (const T*)X2, N2, ldx2, (T*)K, ldk, lower,   \
                                              (T)diag_scalar, (const T*)diag_vec, vec_ok)
  switch (p.n_groups) {
    case 0:
    case 1: GPK_KB_GO(1); break;
    case 2: GPK_KB_GO(2); break;
    case 3: GPK_KB_GO(3); break;
    default: GPK_KB_GO(4); break;
  }
#undef GPK_KB_GO
  GPK_LAUNCH_OK();
  return 0;
}

int kbuild_impl(const gpk_knode* nodes, int n_nodes, const int32_t* dims, const double* ard, const void* X,
                int64_t N, int64_t ldx, const void* X2, int64_t N2, int64_t ldx2, int64_t D, void* K, int64_t ldk,
                int dtype, int uplo, double diag_scalar, const void* diag_vec, cudaStream_t st) {
  GPK_CHECK_ARG(dtype == GPK_F32 || dtype == GPK_F64, "kbuild: bad dtype %d", dtype);
  GPK_CHECK_ARG(X && K, "kbuild: null X or K");
  const bool sym = X2 == nullptr;
  if (sym) { N2 = N; ldx2 = ldx; }
  GPK_CHECK_ARG(N >= 0 && N2 >= 0 && ldx >= D && ldx2 >= D && ldk >= N2, "kbuild: bad shape/stride");
  GPK_CHECK_ARG(sym || (uplo == GPK_FULL && diag_scalar == 0.0 && diag_vec == nullptr),
                "kbuild: uplo=LOWER / diagonal shift need the symmetric form (X2 == NULL)");
  if (N == 0 || N2 == 0) return 0;
  KProg p;
  GPK_TRY(compile_kprog(nodes, n_nodes, dims, ard, D, p));
  p.symmetric = sym ? 1 : 0;
  static const bool no_fast = getenv("GPK_KBUILD_GENERIC") != nullptr;
  if (fast_path_ok(p) && !no_fast) {
    if (dtype == GPK_F64)
      return kbuild_fast_launch<double>(p, X, N, ldx, X2, N2, ldx2, K, ldk, uplo == GPK_LOWER, diag_scalar, diag_vec, st);
    return kbuild_fast_launch<float>(p, X, N, ldx, X2, N2, ldx2, K, ldk, uplo == GPK_LOWER, diag_scalar, diag_vec, st);
  }
  if (dtype == GPK_F64)
    return kbuild_launch<double>(p, X, N, ldx, X2, N2, ldx2, K, ldk, uplo == GPK_LOWER, diag_scalar, diag_vec, st);
  return kbuild_launch<float>(p, X, N, ldx, X2, N2, ldx2, K, ldk, uplo == GPK_LOWER, diag_scalar, diag_vec, st);
}

int kdiag_impl(const gpk_knode* nodes, int n_nodes, const int32_t* dims, const double* ard, const void* X,
               int64_t N, int64_t ldx, int64_t D, void* out, int dtype, cudaStream_t st) {
  GPK_CHECK_ARG(dtype == GPK_F32 || dtype == GPK_F64, "kdiag: bad dtype %d", dtype);
  GPK_CHECK_ARG(X && out && ldx >= D, "kdiag: bad arguments");
  if (N == 0) return 0;
  KProg p;
  GPK_TRY(compile_kprog(nodes, n_nodes, dims, ard, D, p));
  const unsigned blocks = (unsigned)((N + 255) / 256);
  if (dtype == GPK_F64)
    kdiag_kernel<double><<<blocks, 256, 0, st>>>(p, (const double*)X, N, ldx, (double*)out);
  else
    kdiag_kernel<float><<<blocks, 256, 0, st>>>(p, (const float*)X, N, ldx, (float*)out);
  GPK_LAUNCH_OK();
  return 0;
}

}  // namespace gpk
