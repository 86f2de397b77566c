// grm_kernels.cuh - device kernels of the GRM path (CalcGrm, 2.0/plink2_matrix_calc.cc:4555).
//
// Replaces ExpandCenteredVarmaj + the dsyrk/dgemm accumulation (CalcGrmThread / CalcGrmPartThread,
// :4285-4327) and the CalcMissingMatrix popcount pass (:4404-4553) with ONE exact int8 tcgen05
// contraction per variant batch:
//
//   G_ij * obs_ij = sum_v z_iv z_jv,  z_iv = s_v (g_iv - c_v m_iv),  s_v = 1/sqrt(2 p_v q_v), c_v = 2 q_v,
//                 = sum_v g_iv L1_jv + m_iv L2_jv,   L1_jv = w_v (g_jv - c_v m_jv),  L2_jv = -c_v L1_jv,  w_v = s_v^2
//   obs_ij        = sum_v m_iv m_jv      (= M - miss_i - miss_j + bothmiss_ij, :4769-4788)
//
// g (ALT dosage 0/1/2, missing -> 0) and m (non-missing indicator) are exact small integers.  The
// real-valued per-variant 3-entry tables L1_v(g), L2_v(g) are written in fixed point with scale
// 2^F and split into four balanced base-256 digits (int8 in [-128,127]), so
//   sum_v g_iv L_jv = 2^-F * sum_k 256^k * (sum_v g_iv d_k,jv)        <- four exact int32 accumulators
// and the only error is the 2^-(F+1) rounding of each table entry (F chosen per batch from the
// largest |L|: 32 significant bits; see DESIGN.md for the error bound).
//
// Tile = 128 rows (I side: planes g, m) x 96 cols (J side: 8 digit planes + m).  TMEM columns:
// [0,384) digit accumulators D_0..D_3, [384,480) obs counts.
#pragma once
#include "common.cuh"
#include "geno_expand.cuh"
#include "umma.cuh"

namespace pl2 {

constexpr uint32_t kGrmLimbs = 4;
constexpr uint32_t kGrmTabStride = 16;          // uint32 tables per variant (9 used, 64-byte rows)
constexpr uint32_t kGrmKc = 64;                 // variants per stage = two UMMA k-steps
constexpr uint32_t kGrmStages = 3;
constexpr uint32_t kGrmLookahead = 3;
constexpr uint32_t kGrmSuperI = 2 * kTileRows;  // g, m
constexpr uint32_t kGrmSuperJ = 9 * kTileCols;  // L1_0..3, L2_0..3, m
constexpr uint32_t kGrmLboI = operand_lbo(kGrmSuperI);  // 2048
constexpr uint32_t kGrmLboJ = operand_lbo(kGrmSuperJ);  // 6912
constexpr uint32_t kGrmStageBytesI = kGrmSuperI * kGrmKc;  // 8192
constexpr uint32_t kGrmStageBytesJ = kGrmSuperJ * kGrmKc;  // 27648
constexpr uint32_t kGrmStageBytes = kGrmStageBytesI + kGrmStageBytesJ;
constexpr uint32_t kGrmSmemBytes = kGrmStages * kGrmStageBytes + 1024;
constexpr uint32_t kGrmProducerThreads = 256;
constexpr uint32_t kGrmThreads = kGrmProducerThreads + 32;
constexpr uint32_t kGrmTileWords = kTileRows * kTileCols;  // per-tile accumulator entries

// ---- per-variant digit tables: tab[v][p] byte c = digit of plane p for genotype code c.
// planes 0..3 = L1 digits (least significant first), 4..7 = L2 digits, 8 = m (constant).
// lvals[v][0..2] = L1_v(g = 0,1,2) as doubles (host-prepared, see grm.cu); scale = 2^F.
__global__ void grm_tables_kernel(const double* __restrict__ lvals /* [variant][6]: L1(0,1,2), L2(0,1,2) */, uint32_t variant_ct, uint32_t variant_ct_padded, double scale, uint32_t* __restrict__ tab) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= variant_ct_padded) return;
  uint32_t out[kGrmTabStride];
#pragma unroll
  for (uint32_t p = 0; p < kGrmTabStride; ++p) out[p] = 0;
  out[8] = kTabNonmiss;
  if (v < variant_ct) {
#pragma unroll
    for (uint32_t which = 0; which < 2; ++which) {
#pragma unroll
      for (uint32_t g = 0; g < 3; ++g) {
        long long x = __double2ll_rn(lvals[6ull * v + 3 * which + g] * scale);
#pragma unroll
        for (uint32_t k = 0; k < kGrmLimbs; ++k) {
          // balanced base-256 digit in [-128, 127]
          long long d = ((x + 128) & 255) - 128;
          x = (x - d) >> 8;
          out[4 * which + k] |= (static_cast<uint32_t>(d) & 0xFFu) << (8 * g);
        }
      }
    }
  }
#pragma unroll
  for (uint32_t p = 0; p < kGrmTabStride; ++p) tab[static_cast<uint64_t>(v) * kGrmTabStride + p] = out[p];
}

__global__ void __launch_bounds__(kGrmThreads, 1)
grm_tc_kernel(const uint8_t* __restrict__ raw, uint32_t pitch, uint32_t variant_ct_padded /* multiple of kGrmKc */, const uint32_t* __restrict__ tab, double inv_scale, const uint32_t* __restrict__ tile_order, const uint32_t* __restrict__ tile_rt, const uint32_t* __restrict__ tile_tc, double* __restrict__ acc_g, int32_t* __restrict__ acc_obs) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full[kGrmStages];
  __shared__ __align__(8) uint64_t bar_empty[kGrmStages];
  __shared__ __align__(8) uint64_t bar_acc;
  __shared__ uint32_t tmem_base_slot;

  const uint32_t tid = threadIdx.x;
  const uint32_t warp = tid >> 5;
  const uint32_t lane = tid & 31;
  const uint32_t tile = tile_order[blockIdx.x];
  const uint32_t i0 = tile_rt[tile] * kTileRows;
  const uint32_t j0 = tile_tc[tile] * kTileCols;
  const uint32_t stage_iters = variant_ct_padded / kGrmKc;
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;

  if (tid == 0) {
    for (uint32_t s = 0; s < kGrmStages; ++s) {
      mbar_init(&bar_full[s], kGrmProducerThreads);
      mbar_init(&bar_empty[s], 1);
    }
    mbar_init(&bar_acc, 1);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc<512>(&tmem_base_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp < 8) {
    // ---------------- producers ----------------
    // thread = (variant k = tid % 64, role = tid / 64).  role 0: the 128 row-side samples (32 bytes)
    // -> planes g, m.  role r in 1..3: the 96 col-side samples (24 bytes) -> digit planes 3(r-1)..3(r-1)+2
    // through the per-variant tables.
    const uint32_t k = tid & 63;
    const uint32_t role = tid >> 6;
    const bool is_i = role == 0;
    const uint8_t* src = raw + static_cast<uint64_t>(k) * pitch + (is_i ? (i0 / 4) : (j0 / 4));
    const uint64_t stage_stride = static_cast<uint64_t>(kGrmKc) * pitch;
    const uint32_t plane0 = is_i ? 0u : 3u * (role - 1);
    const uint32_t dst_k = is_i ? operand_offset(k, 0, kGrmLboI) : (kGrmStageBytesI + operand_offset(k, plane0 * 6, kGrmLboJ));

    struct Row {
      uint32_t w[8];
      uint32_t t[3];
    };
    auto load_row = [&](uint32_t it) -> Row {
      Row r;
#pragma unroll
      for (uint32_t q = 0; q < 8; ++q) r.w[q] = 0xFFFFFFFFu;
      r.t[0] = r.t[1] = r.t[2] = 0;
      if (it < stage_iters) {
        const uint8_t* p = src + it * stage_stride;
        if (is_i) {
          const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
          const uint4 b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
          r.w[0] = a.x; r.w[1] = a.y; r.w[2] = a.z; r.w[3] = a.w;
          r.w[4] = b.x; r.w[5] = b.y; r.w[6] = b.z; r.w[7] = b.w;
        } else {
          const uint2 a = __ldg(reinterpret_cast<const uint2*>(p));
          const uint2 b = __ldg(reinterpret_cast<const uint2*>(p) + 1);
          const uint2 c = __ldg(reinterpret_cast<const uint2*>(p) + 2);
          r.w[0] = a.x; r.w[1] = a.y; r.w[2] = b.x; r.w[3] = b.y; r.w[4] = c.x; r.w[5] = c.y;
          const uint32_t* trow = tab + (static_cast<uint64_t>(it) * kGrmKc + k) * kGrmTabStride + plane0;
          r.t[0] = __ldg(trow);
          r.t[1] = __ldg(trow + 1);
          r.t[2] = __ldg(trow + 2);
        }
      }
      return r;
    };
    auto sts16 = [](uint32_t addr, const uint4& v) { asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); };

    Row pre[kGrmLookahead];
#pragma unroll
    for (uint32_t d = 0; d < kGrmLookahead; ++d) pre[d] = load_row(d);

    for (uint32_t it0 = 0; it0 < stage_iters; it0 += kGrmLookahead) {
#pragma unroll
      for (uint32_t d = 0; d < kGrmLookahead; ++d) {
        const uint32_t it = it0 + d;
        if (it < stage_iters) {
          const uint32_t s = it % kGrmStages;
          const uint32_t ph = (it / kGrmStages) & 1;
          const Row cur = pre[d];
          pre[d] = load_row(it + kGrmLookahead);
          mbar_wait(&bar_empty[s], ph ^ 1);
          const uint32_t dst = smem_base + s * kGrmStageBytes + dst_k;
          if (is_i) {
#pragma unroll
            for (uint32_t q = 0; q < 8; ++q) {
              const Sel4 sel = make_selectors(cur.w[q]);
              sts16(dst + q * kCoreBytes, expand16(kTabDosage, sel));
              sts16(dst + (8 + q) * kCoreBytes, expand16(kTabNonmiss, sel));
            }
          } else {
#pragma unroll
            for (uint32_t q = 0; q < 6; ++q) {
              const Sel4 sel = make_selectors(cur.w[q]);
              sts16(dst + q * kCoreBytes, expand16(cur.t[0], sel));
              sts16(dst + (6 + q) * kCoreBytes, expand16(cur.t[1], sel));
              sts16(dst + (12 + q) * kCoreBytes, expand16(cur.t[2], sel));
            }
          }
          fence_proxy_async_smem();
          mbar_arrive(&bar_full[s]);
        }
      }
    }

    // ---------------- epilogue ----------------
    mbar_wait(&bar_acc, 0);
    tc_fence_after_sync();
    const uint32_t lane_grp = warp & 3;
    const uint32_t col_half = warp >> 2;
    const uint32_t rpos = 32 * lane_grp + lane;
    const uint32_t rsample = (rpos & ~15u) + PosToSample(rpos & 15u);
    double* g_tile = acc_g + static_cast<uint64_t>(tile) * kGrmTileWords + rsample;
    int32_t* o_tile = acc_obs + static_cast<uint64_t>(tile) * kGrmTileWords + rsample;
    const uint32_t taddr = tmem_base + ((32u * lane_grp) << 16);
#pragma unroll 1
    for (uint32_t chunk = 0; chunk < 3; ++chunk) {
      const uint32_t c0 = col_half * 48 + chunk * 16;  // J position group
      uint32_t d0[16], d1[16], d2[16], d3[16], nn[16];
      tmem_ld16(taddr + c0, d0);
      tmem_ld16(taddr + 96 + c0, d1);
      tmem_ld16(taddr + 192 + c0, d2);
      tmem_ld16(taddr + 288 + c0, d3);
      tmem_ld16(taddr + 384 + c0, nn);
      tmem_ld_wait();
#pragma unroll
      for (uint32_t c = 0; c < 16; ++c) {
        const uint32_t csample = c0 + PosToSample(c);
        const long long tot = static_cast<long long>(static_cast<int32_t>(d0[c])) + (static_cast<long long>(static_cast<int32_t>(d1[c])) << 8) +
                              (static_cast<long long>(static_cast<int32_t>(d2[c])) << 16) + (static_cast<long long>(static_cast<int32_t>(d3[c])) << 24);
        g_tile[static_cast<uint64_t>(csample) * kTileRows] += static_cast<double>(tot) * inv_scale;
        o_tile[static_cast<uint64_t>(csample) * kTileRows] += static_cast<int32_t>(nn[c]);
      }
    }
    tc_fence_before_sync();
  } else {
    if (lane == 0) {
      constexpr uint32_t idesc_n192 = make_idesc_i8(128, 192, true, true);
      constexpr uint32_t idesc_n96 = make_idesc_i8(128, 96, true, true);
      for (uint32_t it = 0; it < stage_iters; ++it) {
        const uint32_t s = it % kGrmStages;
        const uint32_t ph = (it / kGrmStages) & 1;
        mbar_wait(&bar_full[s], ph);
        tc_fence_after_sync();
#pragma unroll
        for (uint32_t kk = 0; kk < kGrmKc / 32; ++kk) {
          const uint32_t si = smem_base + s * kGrmStageBytes + kk * 4 * kGrmLboI;
          const uint32_t sj = smem_base + s * kGrmStageBytes + kGrmStageBytesI + kk * 4 * kGrmLboJ;
          const uint32_t acc = (it | kk) ? 1u : 0u;
          const uint64_t a_g = make_smem_desc(si, kGrmLboI, kCoreBytes);
          const uint64_t a_m = make_smem_desc(si + 8 * kCoreBytes, kGrmLboI, kCoreBytes);
          const uint64_t b_l1_01 = make_smem_desc(sj, kGrmLboJ, kCoreBytes);
          const uint64_t b_l1_23 = make_smem_desc(sj + 12 * kCoreBytes, kGrmLboJ, kCoreBytes);
          const uint64_t b_l2_01 = make_smem_desc(sj + 24 * kCoreBytes, kGrmLboJ, kCoreBytes);
          const uint64_t b_l2_23 = make_smem_desc(sj + 36 * kCoreBytes, kGrmLboJ, kCoreBytes);
          const uint64_t b_m = make_smem_desc(sj + 48 * kCoreBytes, kGrmLboJ, kCoreBytes);
          umma_i8_ss(tmem_base + 0, a_g, b_l1_01, idesc_n192, acc);
          umma_i8_ss(tmem_base + 192, a_g, b_l1_23, idesc_n192, acc);
          umma_i8_ss(tmem_base + 0, a_m, b_l2_01, idesc_n192, 1u);
          umma_i8_ss(tmem_base + 192, a_m, b_l2_23, idesc_n192, 1u);
          umma_i8_ss(tmem_base + 384, a_m, b_m, idesc_n96, acc);
        }
        umma_commit(&bar_empty[s]);
      }
      umma_commit(&bar_acc);
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 8) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---- finalisation: rows [r0, r1) of the lower triangle INCLUDING the diagonal into the reference's
// in-memory layout grm[(j - r0) * row_stride + i], i <= j (CalcGrm :4630, :4769-4788); optionally
// the per-pair observation counts as float (the .grm.N.bin payload, :4985-5019).
__global__ void __launch_bounds__(256)
grm_finalize_kernel(const double* __restrict__ acc_g, const int32_t* __restrict__ acc_obs, const uint32_t* __restrict__ tile_rt, const uint32_t* __restrict__ tile_tc, uint32_t sample_ct, uint32_t r0, uint32_t r1, uint64_t row_stride, int use_obs, double variant_ct_recip, double* __restrict__ out_g, float* __restrict__ out_obs) {
  __shared__ double s_g[16][kTileCols + 1];
  __shared__ int32_t s_o[16][kTileCols + 1];
  const uint32_t tile = blockIdx.x >> 3;
  const uint32_t sub = blockIdx.x & 7;
  const uint32_t row_base = tile_rt[tile] * kTileRows + sub * 16;
  if (row_base >= r1 || row_base + 16 <= r0) return;
  const uint32_t col_base = tile_tc[tile] * kTileCols;
  if (col_base > row_base + 15) return;
  const uint32_t r = threadIdx.x & 15;
  for (uint32_t c = threadIdx.x >> 4; c < kTileCols; c += 16) {
    const uint64_t off = static_cast<uint64_t>(tile) * kGrmTileWords + static_cast<uint64_t>(c) * kTileRows + sub * 16 + r;
    s_g[r][c] = acc_g[off];
    s_o[r][c] = acc_obs[off];
  }
  __syncthreads();
  for (uint32_t idx = threadIdx.x; idx < 16 * kTileCols; idx += 256) {
    const uint32_t rr = idx / kTileCols, cl = idx % kTileCols;
    const uint32_t j = row_base + rr, i = col_base + cl;
    if (j < r0 || j >= r1 || j >= sample_ct || i > j) continue;
    const uint64_t o = static_cast<uint64_t>(j - r0) * row_stride + i;
    const double g = s_g[rr][cl];
    // reference: `/= u31tod(obs)` per entry, or `*= 1.0 / variant_ct` (:4769-4788)
    out_g[o] = use_obs ? (g / static_cast<double>(s_o[rr][cl])) : (g * variant_ct_recip);
    if (out_obs) out_obs[o] = static_cast<float>(s_o[rr][cl]);
  }
}

}  // namespace pl2
