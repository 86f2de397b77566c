// grm_kernels.cuh - device kernels of the GRM path (CalcGrm, 2.0/plink2_matrix_calc.cc:4555).
//
// Replaces ExpandCenteredVarmaj + the dsyrk/dgemm accumulation (CalcGrmThread / CalcGrmPartThread,
// :4285-4327) and the CalcMissingMatrix popcount pass (:4404-4553) with ONE exact int8 tcgen05
// contraction per variant batch:
//
//   G_ij * obs_ij = sum_v z_iv z_jv,   z_v(g) = intercept_v + g * slope_v (0 for missing)   [PopulateRescaledDosage]
//                 = sum_v g_iv L1_v(g_jv) + m_iv L2_v(g_jv),   L1_v(g) = slope_v z_v(g),  L2_v(g) = intercept_v z_v(g)
//   obs_ij        = sum_v m_iv m_jv      (= M - miss_i - miss_j + bothmiss_ij, :4769-4788)
//
// g (ALT dosage 0/1/2, missing -> 0) and m (non-missing indicator) are exact small integers.  The
// real-valued per-variant 3-entry tables L1_v, L2_v are written in fixed point with scale 2^F
// (|L| 2^F < 2^38) and split into FIVE balanced base-256 digits (int8 in [-128,127]):
//   sum_v g_iv L_v(g_jv) = 2^-F * sum_k 256^k * (sum_v g_iv d_k,v(g_jv))     <- five exact int32 accumulators
// so the only error is the 2^-(F+1) rounding of each table entry: 40 significant bits relative to
// the largest |L| of the batch (DESIGN.md has the bound; measured < 1e-10 absolute on G).
//
// Tile = 128 rows (I side: planes g, m) x 80 cols (J side: 10 digit planes + m).  TMEM columns:
// [0,400) digit accumulators D_0..D_4, [400,480) obs counts.
#pragma once
#include "common.cuh"
#include "geno_expand.cuh"
#include "umma.cuh"

namespace pl2 {

constexpr uint32_t kGrmTileCols = 80;
constexpr uint32_t kGrmSamplePad = 640;         // lcm(128, 80)
constexpr uint32_t kGrmLimbs = 5;
constexpr uint32_t kGrmFixedBits = 38;          // |L| * 2^F < 2^38
constexpr uint32_t kGrmPlanesJ = 2 * kGrmLimbs + 1;
constexpr uint32_t kGrmGroupsJ = kGrmTileCols / 16;  // 16-sample groups per J plane
constexpr uint32_t kGrmTabStride = 16;          // uint32 tables per variant (11 used, 64-byte rows)
constexpr uint32_t kGrmKc = 64;                 // variants per stage = two UMMA k-steps
constexpr uint32_t kGrmStages = 3;
constexpr uint32_t kGrmLookahead = 3;
constexpr uint32_t kGrmSuperI = 2 * kTileRows;            // g, m
constexpr uint32_t kGrmSuperJ = kGrmPlanesJ * kGrmTileCols;  // 880
constexpr uint32_t kGrmLboI = operand_lbo(kGrmSuperI);    // 2048
constexpr uint32_t kGrmLboJ = operand_lbo(kGrmSuperJ);    // 7040
constexpr uint32_t kGrmStageBytesI = kGrmSuperI * kGrmKc;
constexpr uint32_t kGrmStageBytesJ = kGrmSuperJ * kGrmKc;
constexpr uint32_t kGrmStageBytes = kGrmStageBytesI + kGrmStageBytesJ;
constexpr uint32_t kGrmSmemBytes = kGrmStages * kGrmStageBytes + 1024;
constexpr uint32_t kGrmProducerThreads = 256;
constexpr uint32_t kGrmThreads = kGrmProducerThreads + 32;
constexpr uint32_t kGrmTileWords = kTileRows * kGrmTileCols;  // per-tile accumulator entries
static_assert(kGrmSmemBytes <= 232448, "GRM pipeline exceeds the 227 KB shared-memory opt-in limit");
static_assert((kGrmLimbs + 1) * kGrmTileCols <= 512, "GRM accumulators exceed TMEM");

// ---- per-variant digit tables: tab[v][p], byte c = digit of plane p for genotype code c.
// planes 0..4 = L1 digits (least significant first), 5..9 = L2 digits, 10 = m (constant).
// lvals[v] = {L1(0), L1(1), L1(2), L2(0), L2(1), L2(2)} as doubles (host-prepared, grm.cu).
__global__ void grm_tables_kernel(const double* __restrict__ lvals, uint32_t variant_ct, uint32_t variant_ct_padded, double scale, uint32_t* __restrict__ tab) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= variant_ct_padded) return;
  uint32_t out[kGrmTabStride];
#pragma unroll
  for (uint32_t p = 0; p < kGrmTabStride; ++p) out[p] = 0;
  out[2 * kGrmLimbs] = kTabNonmiss;
  if (v < variant_ct) {
#pragma unroll
    for (uint32_t which = 0; which < 2; ++which) {
#pragma unroll
      for (uint32_t g = 0; g < 3; ++g) {
        long long x = __double2ll_rn(lvals[6ull * v + 3 * which + g] * scale);
#pragma unroll
        for (uint32_t k = 0; k < kGrmLimbs; ++k) {
          const long long d = ((x + 128) & 255) - 128;  // balanced base-256 digit in [-128, 127]
          x = (x - d) >> 8;
          out[kGrmLimbs * which + k] |= (static_cast<uint32_t>(d) & 0xFFu) << (8 * g);
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
  const uint32_t warp = uniform_warp_idx();
  const uint32_t lane = tid & 31;
  const uint32_t tile = tile_order[blockIdx.x];
  const uint32_t i0 = tile_rt[tile] * kTileRows;
  const uint32_t j0 = tile_tc[tile] * kGrmTileCols;
  const uint32_t stage_iters = variant_ct_padded / kGrmKc;
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;

  if (tid == 0) {
    for (uint32_t s = 0; s < kGrmStages; ++s) {
      mbar_init(&bar_full[s], kGrmProducerThreads / 32);
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
    // thread = (variant k = tid % 64, role = tid / 64).  role 0: the 128 row-side samples (one
    // 32-byte sector) -> planes g, m.  roles 1..3: the 80 col-side samples (20 bytes) -> digit planes
    // {0..3}, {4..7}, {8,9,10} through the per-variant tables.
    const uint32_t k = tid & 63;
    const uint32_t role = tid >> 6;
    const bool is_i = role == 0;
    const uint8_t* src = raw + static_cast<uint64_t>(k) * pitch + (is_i ? (i0 / 4) : (j0 / 4));
    const uint64_t stage_stride = static_cast<uint64_t>(kGrmKc) * pitch;
    const uint32_t plane0 = is_i ? 0u : 4u * (role - 1);
    const uint32_t plane_ct = is_i ? 0u : (role == 3 ? 3u : 4u);
    const uint32_t dst_k = is_i ? operand_offset(k, 0, kGrmLboI) : (kGrmStageBytesI + operand_offset(k, plane0 * kGrmGroupsJ, kGrmLboJ));

    struct Row {
      uint32_t w[8];
      uint32_t t[4];
    };
    auto load_row = [&](uint32_t it) -> Row {
      Row r;
#pragma unroll
      for (uint32_t q = 0; q < 8; ++q) r.w[q] = 0xFFFFFFFFu;
      r.t[0] = r.t[1] = r.t[2] = r.t[3] = 0;
      if (it < stage_iters) {
        const uint8_t* p = src + it * stage_stride;
        if (is_i) {
          const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
          const uint4 b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
          r.w[0] = a.x; r.w[1] = a.y; r.w[2] = a.z; r.w[3] = a.w;
          r.w[4] = b.x; r.w[5] = b.y; r.w[6] = b.z; r.w[7] = b.w;
        } else {
          const uint32_t* q32 = reinterpret_cast<const uint32_t*>(p);
#pragma unroll
          for (uint32_t q = 0; q < kGrmGroupsJ; ++q) r.w[q] = __ldg(q32 + q);
          const uint4 tt = __ldg(reinterpret_cast<const uint4*>(tab + (static_cast<uint64_t>(it) * kGrmKc + k) * kGrmTabStride + plane0));
          r.t[0] = tt.x; r.t[1] = tt.y; r.t[2] = tt.z; r.t[3] = tt.w;
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
            for (uint32_t q = 0; q < kGrmGroupsJ; ++q) {
              const Sel4 sel = make_selectors(cur.w[q]);
              sts16(dst + q * kCoreBytes, expand16(cur.t[0], sel));
              sts16(dst + (kGrmGroupsJ + q) * kCoreBytes, expand16(cur.t[1], sel));
              sts16(dst + (2 * kGrmGroupsJ + q) * kCoreBytes, expand16(cur.t[2], sel));
              if (plane_ct == 4) sts16(dst + (3 * kGrmGroupsJ + q) * kCoreBytes, expand16(cur.t[3], sel));
            }
          }
          fence_proxy_async_smem();
          mbar_arrive_warp(&bar_full[s], lane);
        }
      }
    }

    // ---------------- epilogue ----------------
    mbar_wait(&bar_acc, 0);
    tc_fence_after_sync();
    const uint32_t lane_grp = warp & 3;
    const uint32_t rpos = 32 * lane_grp + lane;
    const uint32_t rsample = (rpos & ~15u) + PosToSample(rpos & 15u);
    double* g_tile = acc_g + static_cast<uint64_t>(tile) * kGrmTileWords + rsample;
    int32_t* o_tile = acc_obs + static_cast<uint64_t>(tile) * kGrmTileWords + rsample;
    const uint32_t taddr = tmem_base + ((32u * lane_grp) << 16);
    // 5 column groups of 16: warps 0-3 take groups {0,2,4}, warps 4-7 take {1,3}
#pragma unroll 1
    for (uint32_t grp = warp >> 2; grp < kGrmGroupsJ; grp += 2) {
      const uint32_t c0 = grp * 16;
      uint32_t d0[16], d1[16], d2[16], d3[16], d4[16], nn[16];
      tmem_ld16(taddr + c0, d0);
      tmem_ld16(taddr + kGrmTileCols + c0, d1);
      tmem_ld16(taddr + 2 * kGrmTileCols + c0, d2);
      tmem_ld16(taddr + 3 * kGrmTileCols + c0, d3);
      tmem_ld16(taddr + 4 * kGrmTileCols + c0, d4);
      tmem_ld16(taddr + 5 * kGrmTileCols + c0, nn);
      tmem_ld_wait();
#pragma unroll
      for (uint32_t c = 0; c < 16; ++c) {
        const uint32_t csample = c0 + PosToSample(c);
        const long long tot = static_cast<long long>(static_cast<int32_t>(d0[c])) + (static_cast<long long>(static_cast<int32_t>(d1[c])) << 8) +
                              (static_cast<long long>(static_cast<int32_t>(d2[c])) << 16) + (static_cast<long long>(static_cast<int32_t>(d3[c])) << 24) +
                              (static_cast<long long>(static_cast<int32_t>(d4[c])) << 32);
        g_tile[static_cast<uint64_t>(csample) * kTileRows] += static_cast<double>(tot) * inv_scale;
        o_tile[static_cast<uint64_t>(csample) * kTileRows] += static_cast<int32_t>(nn[c]);
      }
    }
    tc_fence_before_sync();
  } else {
    // ---------------- UMMA issuer: whole warp loops, one elected lane issues (umma.cuh) ----------------
    constexpr uint32_t idesc_n160 = make_idesc_i8(128, 2 * kGrmTileCols, true, true);
    constexpr uint32_t idesc_n80 = make_idesc_i8(128, kGrmTileCols, true, true);
    constexpr uint32_t kPlaneStep = (kGrmGroupsJ * kCoreBytes) >> 4;  // J plane step inside a k-group (descriptor units)
    const uint32_t tmem_u = uniform_u32(tmem_base);
    const uint64_t desc_i = make_smem_desc(smem_base, kGrmLboI, kCoreBytes);
    const uint64_t desc_j = make_smem_desc(smem_base + kGrmStageBytesI, kGrmLboJ, kCoreBytes);
    uint32_t s = 0, ph = 0;
    for (uint32_t it = 0; it < stage_iters; ++it) {
      mbar_wait(&bar_full[s], ph);
      tc_fence_after_sync();
      if (elect_one_sync()) {
#pragma unroll
        for (uint32_t kk = 0; kk < kGrmKc / 32; ++kk) {
          const uint32_t acc = (it | kk) ? 1u : 0u;
          const uint64_t a_g = desc_i + ((s * kGrmStageBytes + kk * 4 * kGrmLboI) >> 4);
          const uint64_t a_m = a_g + ((8 * kCoreBytes) >> 4);
          const uint64_t bj = desc_j + ((s * kGrmStageBytes + kk * 4 * kGrmLboJ) >> 4);
          umma_i8_ss(tmem_u + 0, a_g, bj, idesc_n160, acc);                                       // g x [d1_0 d1_1]
          umma_i8_ss(tmem_u + 2 * kGrmTileCols, a_g, bj + 2 * kPlaneStep, idesc_n160, acc);       // g x [d1_2 d1_3]
          umma_i8_ss(tmem_u + 4 * kGrmTileCols, a_g, bj + 4 * kPlaneStep, idesc_n80, acc);        // g x d1_4
          umma_i8_ss(tmem_u + 0, a_m, bj + 5 * kPlaneStep, idesc_n160, 1u);                       // m x [d2_0 d2_1]
          umma_i8_ss(tmem_u + 2 * kGrmTileCols, a_m, bj + 7 * kPlaneStep, idesc_n160, 1u);        // m x [d2_2 d2_3]
          umma_i8_ss(tmem_u + 4 * kGrmTileCols, a_m, bj + 9 * kPlaneStep, idesc_n80, 1u);         // m x d2_4
          umma_i8_ss(tmem_u + 5 * kGrmTileCols, a_m, bj + 10 * kPlaneStep, idesc_n80, acc);       // m x m = obs
        }
        umma_commit(&bar_empty[s]);
      }
      __syncwarp();
      if (++s == kGrmStages) {
        s = 0;
        ph ^= 1;
      }
    }
    if (elect_one_sync()) umma_commit(&bar_acc);
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
  __shared__ double s_g[16][kGrmTileCols + 1];
  __shared__ int32_t s_o[16][kGrmTileCols + 1];
  const uint32_t tile = blockIdx.x >> 3;
  const uint32_t sub = blockIdx.x & 7;
  const uint32_t row_base = tile_rt[tile] * kTileRows + sub * 16;
  if (row_base >= r1 || row_base + 16 <= r0) return;
  const uint32_t col_base = tile_tc[tile] * kGrmTileCols;
  if (col_base > row_base + 15) return;
  const uint32_t r = threadIdx.x & 15;
  for (uint32_t c = threadIdx.x >> 4; c < kGrmTileCols; c += 16) {
    const uint64_t off = static_cast<uint64_t>(tile) * kGrmTileWords + static_cast<uint64_t>(c) * kTileRows + sub * 16 + r;
    s_g[r][c] = acc_g[off];
    s_o[r][c] = acc_obs[off];
  }
  __syncthreads();
  for (uint32_t idx = threadIdx.x; idx < 16 * kGrmTileCols; idx += 256) {
    const uint32_t rr = idx / kGrmTileCols, cl = idx % kGrmTileCols;
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
