// planes.cuh — the digit-plane store of one fp64 factorisation (shared by gemm_tc.cu and potrf.cu).
//
// The tcgen05 trailing update (gemm_tc.cu) consumes L as S signed 7-bit digit planes per element,
//     l_ik = 2^(e_i - 6) * sum_s 2^(-7 s) d_s(i,k),   d_s in [-64, 64]  (int8),
// stored PRE-TILED in the canonical no-swizzle K-major UMMA image: tile (rb, kb) = rows [128 rb, 128 rb + 128) x
// columns [32 kb, 32 kb + 32) holds S consecutive planes of 4096 bytes.  Row block rb only ever needs the k-blocks left
// of its diagonal block (kb < 4 rb; extra rows below the square part need all of them), so the tiles are packed
// triangularly: tile (rb, kb) starts at (plane_prefix(rb) + kb) * S * 4096 bytes.
//
// Static scales: |L_ik| <= sqrt(A_ii) for a positive-definite A, so e_i = ilogb(sqrt(A_ii)) + 1 is valid for every
// entry of row i BEFORE the row exists.  The panel-solve kernel (potrf.cu) therefore emits the planes of the columns it
// has just finished directly from shared memory, and no slicing pass over L is needed (the price -- rows of L are often
// well below sqrt(A_ii), so a few leading digit bits are unused -- is measured in scripts/static_scale_study.py:
// max |dL| / max |L| = 1e-11 at N = 2048 with S = 7).  Rows below the square part (the (Y - m)^T rows that ride along)
// have no such bound and are sliced with their running row maximum before each update (slice_rows_kernel).
#pragma once
#include "common.cuh"

namespace gpk {

constexpr int TC_BM = 128, TC_BN = 64, TC_KB = 32;   // CTA tile of the update, bytes (= int8 elements) per k-step
constexpr int TC_ATILE = TC_BM * TC_KB;              // 4096 B per digit plane of a 128-row tile
constexpr int TC_BTILE = TC_BN * TC_KB;              // 2048 B (one half of a 128-row tile)
constexpr int TC_MAXS = 8;

// byte offset of element (row r in [0,128), k in [0,32)) inside one digit-plane tile:
// canonical no-swizzle K-major UMMA layout ((8,n),2):((1,SBO),LBO) in 16-byte units, LBO = 8, SBO = 16
__device__ __host__ __forceinline__ int tc_tile_off(int r, int k) {
  return (r >> 3) * 256 + (k >> 4) * 128 + (r & 7) * 16 + (k & 15);
}

// number of k-block tiles stored before row block rb; nbk = number of 128-column blocks of the square part
__device__ __host__ __forceinline__ int64_t plane_prefix(int64_t rb, int64_t nbk) {
  const int64_t q = rb < nbk + 1 ? rb : nbk + 1;
  return 2 * q * (q - 1) + (rb - q) * 4 * nbk;
}

// Conversions without the XU pipe (F2I / I2F / FRND run at 16 per clock per SM, a quarter of the fp64 FMA rate):
//  * digit extraction: t = v + 1.5 * 2^52 rounds v (|v| < 2^31) to the nearest integer (ties to even, like rint) in the
//    fp64 adder; the integer sits in the low word of t, the rounded value is t - 1.5 * 2^52;
//  * int32 -> double: the bit pattern 0x43300000'(x ^ 0x80000000) is 2^52 + 2^31 + x exactly.
__device__ __forceinline__ double tc_round_digit(double v, int& digit) {
  const double magic = 6755399441055744.0;  // 1.5 * 2^52
  const double t = v + magic;
  digit = __double2loint(t);
  return t - magic;
}
__device__ __forceinline__ double tc_int_to_double(int x) {
  return __hiloint2double(0x43300000, x ^ 0x80000000) - 4503601774854144.0;  // 2^52 + 2^31
}

struct TcPlanes {
  int8_t* planes = nullptr;   // digit planes, triangular tile packing
  double* rowscale = nullptr; // [rows_pad]: 2^(e_i - 6)
  int* err = nullptr;         // device word: protocol error code of the tcgen05 kernel (bounded waits)
  int S = 7;
  int64_t nbk = 0;            // 128-column blocks of the square part
  int64_t n_sq = 0;           // rows >= n_sq are "extra" rows (dynamic scales)
  bool is_static = true;      // false: every update re-slices its operand rows (GPK_TC_STATIC=0)
  bool rect = false;          // experiment (GPK_TC_RECT=1): rectangular instead of triangular tile packing
  __device__ __host__ int8_t* tile(int64_t rb, int64_t kb) const {
    const int64_t pre = rect ? rb * 4 * nbk : plane_prefix(rb, nbk);
    return planes + (size_t)(pre + kb) * S * TC_ATILE;
  }
};

size_t tc_planes_bytes(int64_t n, int64_t rows);
TcPlanes tc_planes_layout(void* ws, int64_t n, int64_t rows, int S);
// rowscale[i] = 2^(ilogb(sqrt(A_ii)) + 1 - 6) for the square rows (reads the ORIGINAL diagonal: call before factorising)
int tc_row_exponents(const double* A, int64_t lda, const TcPlanes& pl, cudaStream_t st);
// slices rows [row0, row0 + nrows) (global indices) of the k-range [k0, k0 + K) with their running row maxima
int tc_slice_rows(const double* P, int64_t ld, int64_t row0, int64_t nrows, int64_t k0, int64_t K, const TcPlanes& pl,
                  cudaStream_t st);
// C[m, n] -= L[r0 : r0 + m, k0 : k0 + K] L[r0 : r0 + n, k0 : k0 + K]^T from the plane store (lower tiles only if `lower`)
int syrk_tc_planes(double* C, int64_t ldc, int64_t m, int64_t n, const TcPlanes& pl, int64_t r0, int64_t k0, int64_t K,
                   int lower, cudaStream_t st, const GemmOpts* opts = nullptr);

}  // namespace gpk
