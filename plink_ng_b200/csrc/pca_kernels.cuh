// pca_kernels.cuh - streaming fp64 passes of `--pca approx` (CalcPca, 2.0/plink2_matrix_calc.cc:5697-5941).
//
// Y is the M x N standardized genotype matrix (ExpandCenteredVarmaj with variance_standardize = 1,
// missing -> 0), never materialised: every element is a per-variant 4-entry table lookup of the
// resident 2-bit genotype.  Two skinny products replace the reference's dgemm calls:
//   pca_xa_kernel  : H[v][c]  = sum_i Y[v][i] * G[i][c]            (CalcPcaXaThread / first half of Xtxa, :5233, :5264)
//   pca_xtb_kernel : B[i][c] += sum_v Y[v][i] * H[v][c]            (second half of Xtxa :5235, CalcPcaXtbThread :5298)
// Both are register-tiled 4 x 4 (variants/samples x columns) fp64 FMA kernels with the small dense
// operand staged in shared memory.  Algorithmic work: N*M*C FMAs per call.
#pragma once
#include "common.cuh"
#include "cp_async.cuh"

namespace pl2 {

constexpr uint32_t kPcaColsMax = 40;   // columns per launch (multiple of 4; 128 x 40 doubles of smem)

// H (column-major, ld = h_ld) [v_global][c] for the variants of this batch.
// grid.x = variant tiles of 128; block = 32 lanes x ceil(C/4) warps; each thread 4 variants x 4 columns
// (C even; a trailing half quad reads two slack doubles and is masked on store).
static __global__ void pca_xa_kernel(const uint8_t* __restrict__ raw, uint32_t pitch, uint32_t sample_ct_padded, uint32_t variant_ct, const double* __restrict__ ztab /* [variant][4] */, const double* __restrict__ g /* row-major [sample][g_ld] */, uint32_t g_ld, uint32_t col0, uint32_t cols,
                                     double* __restrict__ h, uint64_t h_ld, uint64_t v_global0) {
  extern __shared__ __align__(16) double s_g[];  // [128 samples][cols]
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t cw = threadIdx.x >> 5;  // column quad
  const uint32_t v0 = blockIdx.x * 128 + lane * 4;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  double z[4][4];
  const uint8_t* rows[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const uint32_t v = min(v0 + a, variant_ct - 1);
    rows[a] = raw + static_cast<uint64_t>(v) * pitch;
#pragma unroll
    for (int k = 0; k < 4; ++k) z[a][k] = (v0 + a < variant_ct) ? ztab[4ull * v + k] : 0.0;
  }
  for (uint32_t s0 = 0; s0 < sample_ct_padded; s0 += 128) {
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < 128 * cols; i += blockDim.x) {
      const uint32_t s = i / cols, c = i % cols;
      s_g[i] = g[static_cast<uint64_t>(s0 + s) * g_ld + col0 + c];
    }
    __syncthreads();
#pragma unroll 1
    for (uint32_t w = 0; w < 8; ++w) {  // 8 words of 16 samples
      uint32_t word[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) word[a] = __ldg(reinterpret_cast<const uint32_t*>(rows[a] + (s0 / 4) + 4 * w));
#pragma unroll
      for (uint32_t k = 0; k < 16; ++k) {
        const double* gp = &s_g[(16 * w + k) * cols + 4 * cw];
        const double2 g01 = *reinterpret_cast<const double2*>(gp);
        const double2 g23 = *reinterpret_cast<const double2*>(gp + 2);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const uint32_t code = (word[a] >> (2 * k)) & 3u;
          const double zz = (code & 2u) ? ((code & 1u) ? 0.0 : z[a][2]) : ((code & 1u) ? z[a][1] : z[a][0]);
          acc[a][0] = fma(zz, g01.x, acc[a][0]);
          acc[a][1] = fma(zz, g01.y, acc[a][1]);
          acc[a][2] = fma(zz, g23.x, acc[a][2]);
          acc[a][3] = fma(zz, g23.y, acc[a][3]);
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    if (v0 + a < variant_ct) {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const uint32_t c = 4 * cw + b;
        if (c < cols) h[static_cast<uint64_t>(col0 + c) * h_ld + v_global0 + v0 + a] = acc[a][b];
      }
    }
  }
}

// out[i][c] += sum_v Y[v][i] * H[v][c] for this batch's variants.  H column-major (ld = h_ld).
// out element (i, c) lives at out[i * out_rs + c * out_cs].
// grid.x = sample tiles of 128; block = 32 lanes (4 samples each) x (C/4) warps.
static __global__ void pca_xtb_kernel(const uint8_t* __restrict__ raw, uint32_t pitch, uint32_t sample_ct, uint32_t variant_ct, const double* __restrict__ ztab, const double* __restrict__ h, uint64_t h_ld, uint64_t v_global0, uint32_t col0, uint32_t cols,
                                      double* __restrict__ out, uint64_t out_rs, uint64_t out_cs) {
  extern __shared__ __align__(16) double s_h[];  // [128 variants][cols] then [128][4] z tables
  double* s_z = s_h + 128 * cols;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t cw = threadIdx.x >> 5;
  const uint32_t i0 = blockIdx.x * 128 + lane * 4;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  const uint8_t* col = raw + blockIdx.x * 32 + lane;  // byte holding this lane's 4 samples
  for (uint32_t vb = 0; vb < variant_ct; vb += 128) {
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < 128 * cols; i += blockDim.x) {
      const uint32_t c = i / 128, v = i % 128;
      s_h[v * cols + c] = (vb + v < variant_ct) ? h[static_cast<uint64_t>(col0 + c) * h_ld + v_global0 + vb + v] : 0.0;
    }
    for (uint32_t i = threadIdx.x; i < 512; i += blockDim.x) s_z[i] = (vb + i / 4 < variant_ct) ? ztab[4ull * vb + i] : 0.0;
    __syncthreads();
    const uint32_t vend = min(128u, variant_ct - vb);
#pragma unroll 4
    for (uint32_t v = 0; v < vend; ++v) {
      const uint32_t byte = __ldg(col + static_cast<uint64_t>(vb + v) * pitch);
      const double2 h01 = *reinterpret_cast<const double2*>(&s_h[v * cols + 4 * cw]);
      const double2 h23 = *reinterpret_cast<const double2*>(&s_h[v * cols + 4 * cw + 2]);
      const double z0 = s_z[4 * v], z1 = s_z[4 * v + 1], z2 = s_z[4 * v + 2];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const uint32_t code = (byte >> (2 * a)) & 3u;
        const double zz = (code & 2u) ? ((code & 1u) ? 0.0 : z2) : ((code & 1u) ? z1 : z0);
        acc[a][0] = fma(zz, h01.x, acc[a][0]);
        acc[a][1] = fma(zz, h01.y, acc[a][1]);
        acc[a][2] = fma(zz, h23.x, acc[a][2]);
        acc[a][3] = fma(zz, h23.y, acc[a][3]);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    if (i0 + a < sample_ct) {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const uint32_t c = 4 * cw + b;
        if (c < cols) out[static_cast<uint64_t>(i0 + a) * out_rs + static_cast<uint64_t>(col0 + c) * out_cs] += acc[a][b];
      }
    }
  }
}

static __global__ void scale_kernel(double* __restrict__ x, uint64_t n, double s) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) x[i] *= s;
}

}  // namespace pl2
