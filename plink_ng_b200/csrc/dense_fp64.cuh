// dense_fp64.cuh - the small fp64 matrix products around the PCA factorizations (no cuBLAS in this library):
//   GemmTN :  C (p x n)    = A^T B            A: rows x p, B: rows x n   (reduction over the long dimension, split
//                                              over CTAs; partial sums are added in a fixed order -> bit-reproducible)
//   GemmNN :  O (rows x n) = [O -] A C [diag] A: rows x p, C: p x n      (block Gram-Schmidt update / back-transform)
// All matrices column-major.  64 x 64 output tiles, 256 threads, 4 x 4 accumulators per thread, K tiles of 32 staged
// in shared memory (rows of the tile padded to 65 doubles: conflict-free stores and loads).  These replace the
// dgemm calls inside the reference's SvdRectFused / CalcPca path (2.0/plink2_matrix.cc:1039, :734); they are
// CUDA-core fp64 work, a few 1e10..1e11 flops per --pca approx run.
#pragma once
#include "common.cuh"

namespace pl2 {

constexpr uint32_t kDgTile = 64, kDgK = 32, kDgPad = 65, kDgThreads = 256;

// partial[split][n][p] (column-major p x n per split) = A[r0:r1, i]^T B[r0:r1, j]
static __global__ void __launch_bounds__(kDgThreads) dgemm_tn_kernel(const double* __restrict__ a, uint64_t lda, uint32_t p, const double* __restrict__ b, uint64_t ldb, uint32_t n, uint32_t rows, uint32_t rows_per_split, double* __restrict__ partial) {
  __shared__ double sa[kDgK * kDgPad], sb[kDgK * kDgPad];
  const uint32_t i0 = blockIdx.x * kDgTile, j0 = blockIdx.y * kDgTile;
  const uint32_t r_begin = blockIdx.z * rows_per_split, r_end = min(rows, r_begin + rows_per_split);
  const uint32_t ti = threadIdx.x & 15, tj = threadIdx.x >> 4;
  double acc[4][4] = {};
  for (uint32_t r0 = r_begin; r0 < r_end; r0 += kDgK) {
    // 32 rows x 64 columns per operand: lane -> row (coalesced), 8 warps x 8 iterations -> columns
    for (uint32_t e = threadIdx.x; e < kDgK * kDgTile; e += kDgThreads) {
      const uint32_t kk = e & 31, cc = e >> 5;
      const uint32_t r = r0 + kk;
      sa[kk * kDgPad + cc] = (r < r_end && i0 + cc < p) ? a[static_cast<uint64_t>(i0 + cc) * lda + r] : 0.0;
      sb[kk * kDgPad + cc] = (r < r_end && j0 + cc < n) ? b[static_cast<uint64_t>(j0 + cc) * ldb + r] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (uint32_t kk = 0; kk < kDgK; ++kk) {
      double av[4], bv[4];
#pragma unroll
      for (uint32_t x = 0; x < 4; ++x) {
        av[x] = sa[kk * kDgPad + ti + 16 * x];
        bv[x] = sb[kk * kDgPad + tj + 16 * x];
      }
#pragma unroll
      for (uint32_t x = 0; x < 4; ++x)
#pragma unroll
        for (uint32_t y = 0; y < 4; ++y) acc[x][y] = fma(av[x], bv[y], acc[x][y]);
    }
    __syncthreads();
  }
  double* out = partial + static_cast<uint64_t>(blockIdx.z) * p * n;
#pragma unroll
  for (uint32_t x = 0; x < 4; ++x)
#pragma unroll
    for (uint32_t y = 0; y < 4; ++y) {
      const uint32_t i = i0 + ti + 16 * x, j = j0 + tj + 16 * y;
      if (i < p && j < n) out[static_cast<uint64_t>(j) * p + i] = acc[x][y];
    }
}

// c[j * ldc + i] = sum over splits (fixed order)
static __global__ void __launch_bounds__(256) dgemm_tn_reduce_kernel(const double* __restrict__ partial, uint32_t splits, uint32_t p, uint32_t n, double* __restrict__ c, uint64_t ldc) {
  const uint64_t idx = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<uint64_t>(p) * n) return;
  double s = 0.0;
  for (uint32_t k = 0; k < splits; ++k) s += partial[static_cast<uint64_t>(k) * p * n + idx];
  c[(idx / p) * ldc + (idx % p)] = s;
}

// out[r, j] = (subtract ? out[r, j] - s : s * colscale[j]),  s = sum_k a[r, k] c[k, j]
static __global__ void __launch_bounds__(kDgThreads) dgemm_nn_kernel(const double* __restrict__ a, uint64_t lda, uint32_t rows, uint32_t p, const double* __restrict__ c, uint64_t ldc, uint32_t n, double* __restrict__ out, uint64_t ldo, int subtract, const double* __restrict__ colscale) {
  __shared__ double sa[kDgK * kDgPad], sc[kDgK * kDgPad];
  const uint32_t r0 = blockIdx.x * kDgTile, j0 = blockIdx.y * kDgTile;
  const uint32_t ti = threadIdx.x & 15, tj = threadIdx.x >> 4;
  double acc[4][4] = {};
  for (uint32_t k0 = 0; k0 < p; k0 += kDgK) {
    for (uint32_t e = threadIdx.x; e < kDgK * kDgTile; e += kDgThreads) {
      {  // A tile: 64 rows x 32 k, lane -> row
        const uint32_t rr = e & 63, kk = e >> 6;
        sa[kk * kDgPad + rr] = (r0 + rr < rows && k0 + kk < p) ? a[static_cast<uint64_t>(k0 + kk) * lda + r0 + rr] : 0.0;
      }
      {  // C tile: 32 k x 64 columns, lane -> k
        const uint32_t kk = e & 31, cc = e >> 5;
        sc[kk * kDgPad + cc] = (k0 + kk < p && j0 + cc < n) ? c[static_cast<uint64_t>(j0 + cc) * ldc + k0 + kk] : 0.0;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (uint32_t kk = 0; kk < kDgK; ++kk) {
      double av[4], cv[4];
#pragma unroll
      for (uint32_t x = 0; x < 4; ++x) {
        av[x] = sa[kk * kDgPad + ti + 16 * x];
        cv[x] = sc[kk * kDgPad + tj + 16 * x];
      }
#pragma unroll
      for (uint32_t x = 0; x < 4; ++x)
#pragma unroll
        for (uint32_t y = 0; y < 4; ++y) acc[x][y] = fma(av[x], cv[y], acc[x][y]);
    }
    __syncthreads();
  }
#pragma unroll
  for (uint32_t x = 0; x < 4; ++x)
#pragma unroll
    for (uint32_t y = 0; y < 4; ++y) {
      const uint32_t r = r0 + ti + 16 * x, j = j0 + tj + 16 * y;
      if (r < rows && j < n) {
        double* dst = &out[static_cast<uint64_t>(j) * ldo + r];
        *dst = subtract ? (*dst - acc[x][y]) : (colscale ? acc[x][y] * colscale[j] : acc[x][y]);
      }
    }
}

// Host helpers (all work on ctx->stream).  d_partial must hold DgemmTNPartialDoubles(...) doubles.
static inline uint32_t DgemmTNSplits(const Ctx* c, uint32_t p, uint32_t n, uint32_t rows) {
  const uint32_t tiles = DivUpU32(p, kDgTile) * DivUpU32(n, kDgTile);
  const uint32_t want = DivUpU32(4 * static_cast<uint32_t>(c->sm_count), tiles);
  return std::max(1u, std::min({want, DivUpU32(rows, 8 * kDgK), 512u}));
}
static inline uint64_t DgemmTNPartialDoubles(const Ctx* c, uint32_t p, uint32_t n, uint32_t rows) { return static_cast<uint64_t>(DgemmTNSplits(c, p, n, rows)) * p * n; }

static inline int DgemmTN(Ctx* c, const double* a, uint64_t lda, uint32_t p, const double* b, uint64_t ldb, uint32_t n, uint32_t rows, double* d_partial, double* out, uint64_t ldc) {
  if (!p || !n) return 0;
  uint32_t splits = DgemmTNSplits(c, p, n, rows);
  const uint32_t rows_per_split = RoundUpU32(DivUpU32(rows, splits), kDgK);
  splits = DivUpU32(rows, rows_per_split);
  dgemm_tn_kernel<<<dim3(DivUpU32(p, kDgTile), DivUpU32(n, kDgTile), splits), kDgThreads, 0, c->stream>>>(a, lda, p, b, ldb, n, rows, rows_per_split, d_partial);
  dgemm_tn_reduce_kernel<<<static_cast<uint32_t>(DivUpU64(static_cast<uint64_t>(p) * n, 256)), 256, 0, c->stream>>>(d_partial, splits, p, n, out, ldc);
  c->launches += 2;
  return cudaGetLastError() != cudaSuccess;
}

static inline int DgemmNN(Ctx* c, const double* a, uint64_t lda, uint32_t rows, uint32_t p, const double* cm, uint64_t ldc, uint32_t n, double* out, uint64_t ldo, bool subtract, const double* colscale) {
  if (!rows || !n) return 0;
  dgemm_nn_kernel<<<dim3(DivUpU32(rows, kDgTile), DivUpU32(n, kDgTile)), kDgThreads, 0, c->stream>>>(a, lda, rows, p, cm, ldc, n, out, ldo, subtract ? 1 : 0, colscale);
  c->launches++;
  return cudaGetLastError() != cudaSuccess;
}

}  // namespace pl2
