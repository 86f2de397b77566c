// umma_probe.cuh - single-CTA tcgen05 probe: runs int8 UMMAs over caller-supplied shared-memory
// images with caller-supplied descriptor fields and returns the TMEM accumulator.  Used by
// pl2gpu_selftest_umma (production layout vs scalar reference) and by tests/ to pin the operand
// layout the production kernels rely on.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace pl2 {

struct UmmaProbeParams {
  uint32_t a_bytes, b_bytes;        // image sizes (A at smem offset 0, B at b_smem_off)
  uint32_t b_smem_off;              // multiple of 128
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  uint32_t a_step_bytes, b_step_bytes;  // descriptor start-address advance per k-step
  uint32_t k_steps;
  uint32_t idesc;
  uint32_t n;                       // columns to read back (multiple of 16, <= 256)
};

constexpr uint32_t kProbeSmemBytes = 96 * 1024;

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const uint8_t* __restrict__ a_img, const uint8_t* __restrict__ b_img, UmmaProbeParams prm, int32_t* __restrict__ d_out /* [128][n] */) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_base_slot;
  const uint32_t tid = threadIdx.x;
  const uint32_t warp = tid >> 5;
  const uint32_t lane = tid & 31;
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;
  uint8_t* sm = smem + (smem_base - smem_u32(smem));
  for (uint32_t i = tid; i < prm.a_bytes; i += 128) sm[i] = a_img[i];
  for (uint32_t i = tid; i < prm.b_bytes; i += 128) sm[prm.b_smem_off + i] = b_img[i];
  if (tid == 0) {
    mbar_init(&bar_done, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<256>(&tmem_base_slot);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;
  if (tid == 0) {
    for (uint32_t ks = 0; ks < prm.k_steps; ++ks) {
      const uint64_t da = make_smem_desc(smem_base + ks * prm.a_step_bytes, prm.a_lbo, prm.a_sbo);
      const uint64_t db = make_smem_desc(smem_base + prm.b_smem_off + ks * prm.b_step_bytes, prm.b_lbo, prm.b_sbo);
      umma_i8_ss(tmem_base, da, db, prm.idesc, ks ? 1u : 0u);
    }
    umma_commit(&bar_done);
  }
  mbar_wait(&bar_done, 0);
  tc_fence_after_sync();
  for (uint32_t c0 = 0; c0 < prm.n; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem_base + ((32u * warp) << 16) + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (uint32_t c = 0; c < 16; ++c) d_out[static_cast<uint64_t>(32 * warp + lane) * prm.n + c0 + c] = static_cast<int32_t>(v[c]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc<256>(tmem_base);
  }
}


// TS probe: A rows come from global memory ([128][k_steps*32] int8, K-major), are written to TMEM
// with tcgen05.st (8 columns per k-step at column 256 + 8*ks) and multiplied with the smem B image.
__global__ void __launch_bounds__(128, 1)
umma_probe_ts_kernel(const uint8_t* __restrict__ a_rows, const uint8_t* __restrict__ b_img, UmmaProbeParams prm, int32_t* __restrict__ d_out /* [128][n] */) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_base_slot;
  const uint32_t tid = threadIdx.x;
  const uint32_t warp = tid >> 5;
  const uint32_t lane = tid & 31;
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;
  uint8_t* sm = smem + (smem_base - smem_u32(smem));
  for (uint32_t i = tid; i < prm.b_bytes; i += 128) sm[i] = b_img[i];
  if (tid == 0) {
    mbar_init(&bar_done, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_slot);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t row_bytes = prm.k_steps * 32;
  for (uint32_t ks = 0; ks < prm.k_steps; ++ks) {
    uint32_t v[8];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a_rows + static_cast<uint64_t>(tid) * row_bytes + ks * 32);
#pragma unroll
    for (uint32_t c = 0; c < 8; ++c) v[c] = src[c];
    tmem_st8(tmem_base + ((32u * warp) << 16) + 256 + 8 * ks, v);
  }
  tmem_st_wait();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (tid == 0) {
    for (uint32_t ks = 0; ks < prm.k_steps; ++ks) {
      const uint64_t db = make_smem_desc(smem_base + ks * prm.b_step_bytes, prm.b_lbo, prm.b_sbo);
      umma_i8_ts(tmem_base, tmem_base + 256 + 8 * ks, db, prm.idesc, ks ? 1u : 0u);
    }
    umma_commit(&bar_done);
  }
  mbar_wait(&bar_done, 0);
  tc_fence_after_sync();
  for (uint32_t c0 = 0; c0 < prm.n; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem_base + ((32u * warp) << 16) + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (uint32_t c = 0; c < 16; ++c) d_out[static_cast<uint64_t>(32 * warp + lane) * prm.n + c0 + c] = static_cast<int32_t>(v[c]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace pl2
