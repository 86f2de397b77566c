// king_ts_kernel.cuh - KING pair counts, "TS" tensor kernel: the row-side (A) operand is expanded
// from a SAMPLE-major copy of the genotype block straight into tensor memory (tcgen05.st), so only
// the column-side (B) operand goes through shared memory.  In the SS kernel (king_kernels.cuh) every
// operand byte is written to and read back from shared memory (~200 B/clk wanted vs 128 B/clk
// available, profiles/r01_ncu_king_v1.md); here shared-memory traffic drops to ~100 B/clk and the
// tensor pipe becomes the limiter.
//
// Tile = 128 rows x 80 cols.  TMEM columns: [0,400) accumulators TT|TH, HT|HH, SS (int32),
// [400,496) four A slots of 24 columns (planes T, H, S; 8 columns = 32 K-bytes per lane).
// Same products and the same raw accumulator semantics as king_tc_kernel (tile width 80).
#pragma once
#include "common.cuh"
#include "geno_expand.cuh"
#include "umma.cuh"

namespace pl2 {

constexpr uint32_t kTsCols = 80;
constexpr uint32_t kTsSamplePad = 640;  // lcm(128, 80)
constexpr uint32_t kTsGroupsJ = kTsCols / 16;
constexpr uint32_t kTsAccCols = 5 * kTsCols;   // 400
constexpr uint32_t kTsTileAccWords = kTsAccCols * kTileRows;
constexpr uint32_t kTsASlots = 4;
constexpr uint32_t kTsASlotCols = 24;
constexpr uint32_t kTsKcJ = 64;                // variants per shared-memory stage (two k-steps)
constexpr uint32_t kTsStagesJ = 4;
constexpr uint32_t kTsLboJ = (3 * kTsCols / 16) * kCoreBytes + 64;  // 1984: +64 keeps the K-permuted rows bank-conflict free
constexpr uint32_t kTsStageBytesJ = (kTsKcJ / 8) * kTsLboJ;         // 15872
constexpr uint32_t kTsSmemBytes = kTsStagesJ * kTsStageBytesJ + 1024;
constexpr uint32_t kTsThreads = 288;

// ---- 2-bit matrix transpose: raw[variant][pitch] -> rawT[sample][pitch_t] (pitch_t = variants/4 bytes).
// One CTA = 64 variants x 64 samples through a shared-memory byte tile.
static __global__ void __launch_bounds__(256) geno_transpose_kernel(const uint8_t* __restrict__ raw, uint32_t pitch, uint8_t* __restrict__ raw_t, uint32_t pitch_t) {
  __shared__ uint8_t tile[64][68];
  const uint32_t v0 = blockIdx.x * 64, s0 = blockIdx.y * 64;
  const uint32_t t = threadIdx.x;
  {
    const uint32_t v = t >> 2, sw = t & 3;
    const uint32_t w = *reinterpret_cast<const uint32_t*>(raw + static_cast<uint64_t>(v0 + v) * pitch + s0 / 4 + 4 * sw);
#pragma unroll
    for (uint32_t j = 0; j < 16; ++j) tile[v][16 * sw + j] = static_cast<uint8_t>((w >> (2 * j)) & 3u);
  }
  __syncthreads();
  {
    const uint32_t s = t >> 2, vw = t & 3;
    uint32_t w = 0;
#pragma unroll
    for (uint32_t j = 0; j < 16; ++j) w |= static_cast<uint32_t>(tile[16 * vw + j][s]) << (2 * j);
    *reinterpret_cast<uint32_t*>(raw_t + static_cast<uint64_t>(s0 + s) * pitch_t + v0 / 4 + 4 * vw) = w;
  }
}

__global__ void __launch_bounds__(kTsThreads, 1)
king_ts_kernel(const uint8_t* __restrict__ raw, uint32_t pitch, const uint8_t* __restrict__ raw_t, uint32_t pitch_t, uint32_t variant_ct_padded /* multiple of 64 */, const uint32_t* __restrict__ tile_order, const uint32_t* __restrict__ tile_rt, const uint32_t* __restrict__ tile_tc, int32_t* __restrict__ raw_acc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full_a[kTsASlots];
  __shared__ __align__(8) uint64_t bar_empty_a[kTsASlots];
  __shared__ __align__(8) uint64_t bar_full_b[kTsStagesJ];
  __shared__ __align__(8) uint64_t bar_empty_b[kTsStagesJ];
  __shared__ __align__(8) uint64_t bar_acc;
  __shared__ uint32_t tmem_base_slot;

  const uint32_t tid = threadIdx.x;
  const uint32_t warp = tid >> 5;
  const uint32_t lane = tid & 31;
  const uint32_t tile = tile_order[blockIdx.x];
  const uint32_t i0 = tile_rt[tile] * kTileRows;
  const uint32_t j0 = tile_tc[tile] * kTsCols;
  const uint32_t stage_iters = variant_ct_padded / kTsKcJ;
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;

  if (tid == 0) {
    for (uint32_t s = 0; s < kTsASlots; ++s) {
      mbar_init(&bar_full_a[s], 128);
      mbar_init(&bar_empty_a[s], 1);
    }
    for (uint32_t s = 0; s < kTsStagesJ; ++s) {
      mbar_init(&bar_full_b[s], 128);
      mbar_init(&bar_empty_b[s], 1);
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
    // Two groups of four warps (group = warp / 4).  Group g expands the ROW side of k-steps ks = 2 it + g
    // straight into tensor memory (thread = TMEM lane = sample i0 + 32 (warp % 4) + lane) and the COLUMN
    // side of the shared-memory stages it with it % 2 == g (thread = variant k of the stage, half of its
    // 20-byte row).  The tcgen05.st of one k-step is overlapped with the column-side work and with the
    // expansion of the group's next k-step; tcgen05.wait::st + arrive come last.
    const uint32_t grp = warp >> 2;
    const uint32_t lq = warp & 3;
    const uint32_t row = 32 * lq + lane;
    const uint8_t* src_i = raw_t + static_cast<uint64_t>(i0 + row) * pitch_t;
    const uint32_t taddr_lane = tmem_base + ((32u * lq) << 16) + kTsAccCols;
    const uint32_t t = row;                        // 0..127 inside the group
    const uint32_t k = t & 63;
    const uint32_t half = t >> 6;                  // 0: words 0..2, 1: words 3..4
    const uint32_t w0 = half ? 3u : 0u;
    const uint32_t wn = half ? 2u : 3u;
    const uint8_t* src_j = raw + static_cast<uint64_t>(k) * pitch + j0 / 4 + 4 * w0;
    const uint64_t stage_stride = static_cast<uint64_t>(kTsKcJ) * pitch;
    // K rows of the column side are stored in the PRMT position order of the row side
    // (geno_expand.cuh): variant k of a 16-variant group sits at row SampleToPos(k % 16)
    const uint32_t kpos = (k & ~15u) + SampleToPos(k & 15u);
    const uint32_t dst_k = (kpos >> 3) * kTsLboJ + (kpos & 7) * 16 + w0 * kCoreBytes;

    auto load_i = [&](uint32_t it) -> uint2 {  // row-side words of k-step 2 it + grp
      const uint32_t ks = 2 * it + grp;
      return (it < stage_iters) ? __ldg(reinterpret_cast<const uint2*>(src_i + 8ull * ks)) : make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
    };
    struct RowJ {
      uint32_t w[3];
    };
    auto load_j = [&](uint32_t it) -> RowJ {  // column-side words of stage it
      RowJ r;
      r.w[0] = r.w[1] = r.w[2] = 0xFFFFFFFFu;
      if (it < stage_iters) {
        const uint32_t* p = reinterpret_cast<const uint32_t*>(src_j + it * stage_stride);
        r.w[0] = __ldg(p);
        r.w[1] = __ldg(p + 1);
        if (wn == 3) r.w[2] = __ldg(p + 2);
      }
      return r;
    };
    struct ExpI {
      uint32_t v[3][8];
    };
    auto expand_i = [&](const uint2& w) -> ExpI {
      ExpI e;
      const Sel4 s0 = make_selectors(w.x), s1 = make_selectors(w.y);
      const uint32_t tabs[3] = {kTabHet, kTabHom, kTabSgn};
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        const uint4 a = expand16(tabs[p], s0), b = expand16(tabs[p], s1);
        e.v[p][0] = a.x; e.v[p][1] = a.y; e.v[p][2] = a.z; e.v[p][3] = a.w;
        e.v[p][4] = b.x; e.v[p][5] = b.y; e.v[p][6] = b.z; e.v[p][7] = b.w;
      }
      return e;
    };

    constexpr uint32_t kLa = 4;  // iterations of lookahead (row side: 4 k-steps of this group; column side: 2 stages)
    uint2 pre_i[kLa];
    RowJ pre_j[kLa / 2];
#pragma unroll
    for (uint32_t d = 0; d < kLa; ++d) pre_i[d] = load_i(d);
#pragma unroll
    for (uint32_t d = 0; d < kLa / 2; ++d) pre_j[d] = load_j(2 * d + grp);
    ExpI cur = expand_i(pre_i[0]);
    for (uint32_t it0 = 0; it0 < stage_iters; it0 += kLa) {
#pragma unroll
      for (uint32_t d = 0; d < kLa; ++d) {
        const uint32_t it = it0 + d;
        if (it < stage_iters) {
          const uint32_t ks = 2 * it + grp;
          const uint32_t slot = ks % kTsASlots;
          pre_i[d] = load_i(it + kLa);
          mbar_wait(&bar_empty_a[slot], ((ks / kTsASlots) & 1) ^ 1);
          tc_fence_after_sync();
          const uint32_t ta = taddr_lane + slot * kTsASlotCols;
          tmem_st8(ta, cur.v[0]);
          tmem_st8(ta + 8, cur.v[1]);
          tmem_st8(ta + 16, cur.v[2]);
          if ((d & 1) == grp) {  // this group's turn on the column side (it0 is a multiple of 4)
            const RowJ rj = pre_j[d >> 1];
            pre_j[d >> 1] = load_j(it + kLa);
            const uint32_t sb = it % kTsStagesJ;
            mbar_wait(&bar_empty_b[sb], ((it / kTsStagesJ) & 1) ^ 1);
            const uint32_t dst = smem_base + sb * kTsStageBytesJ + dst_k;
#pragma unroll
            for (uint32_t q = 0; q < 3; ++q) {
              if (q < wn) {
                const Sel4 sel = make_selectors(rj.w[q]);
                const uint4 vt = expand16(kTabHet, sel), vh = expand16(kTabHom, sel), vs = expand16(kTabSgn, sel);
                const uint32_t a0 = dst + q * kCoreBytes;
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a0), "r"(vt.x), "r"(vt.y), "r"(vt.z), "r"(vt.w) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a0 + kTsGroupsJ * kCoreBytes), "r"(vh.x), "r"(vh.y), "r"(vh.z), "r"(vh.w) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a0 + 2 * kTsGroupsJ * kCoreBytes), "r"(vs.x), "r"(vs.y), "r"(vs.z), "r"(vs.w) : "memory");
              }
            }
            fence_proxy_async_smem();
            mbar_arrive(&bar_full_b[sb]);
          }
          const ExpI nxt = expand_i(pre_i[(d + 1) % kLa]);  // words of iteration it + 1 (already resident)
          tmem_st_wait();
          tc_fence_before_sync();
          mbar_arrive(&bar_full_a[slot]);
          cur = nxt;
        }
      }
    }
  } else {
    // ---------------- UMMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t idesc_n160 = make_idesc_i8(128, 2 * kTsCols, false, true);
      constexpr uint32_t idesc_n80 = make_idesc_i8(128, kTsCols, false, true);
      for (uint32_t it = 0; it < stage_iters; ++it) {
        const uint32_t sb = it % kTsStagesJ;
        mbar_wait(&bar_full_b[sb], (it / kTsStagesJ) & 1);
#pragma unroll
        for (uint32_t kk = 0; kk < 2; ++kk) {
          const uint32_t ks = 2 * it + kk;
          const uint32_t slot = ks % kTsASlots;
          mbar_wait(&bar_full_a[slot], (ks / kTsASlots) & 1);
          tc_fence_after_sync();
          const uint32_t acc = ks ? 1u : 0u;
          const uint32_t bj = smem_base + sb * kTsStageBytesJ + kk * 4 * kTsLboJ;
          const uint64_t b_th = make_smem_desc(bj, kTsLboJ, kCoreBytes);
          const uint64_t b_s = make_smem_desc(bj + 2 * kTsGroupsJ * kCoreBytes, kTsLboJ, kCoreBytes);
          const uint32_t ta = tmem_base + kTsAccCols + slot * kTsASlotCols;
          umma_i8_ts(tmem_base + 0, ta, b_th, idesc_n160, acc);
          umma_i8_ts(tmem_base + 2 * kTsCols, ta + 8, b_th, idesc_n160, acc);
          umma_i8_ts(tmem_base + 4 * kTsCols, ta + 16, b_s, idesc_n80, acc);
          umma_commit(&bar_empty_a[slot]);
        }
        umma_commit(&bar_empty_b[sb]);
      }
      umma_commit(&bar_acc);
    }
    __syncwarp();
  }

  if (warp < 8) {
    // ---------------- epilogue: TMEM -> raw accumulators (+=) ----------------
    mbar_wait(&bar_acc, 0);
    tc_fence_after_sync();
    const uint32_t lane_grp = warp & 3;
    const uint32_t rsample = 32 * lane_grp + lane;  // rows are in natural sample order here
    int32_t* acc_tile = raw_acc + static_cast<uint64_t>(tile) * kTsTileAccWords + rsample;
    const uint32_t chunk_begin = (warp < 4) ? 0u : 13u, chunk_end = (warp < 4) ? 13u : 25u;
#pragma unroll 1
    for (uint32_t chunk = chunk_begin; chunk < chunk_end; ++chunk) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((32u * lane_grp) << 16) + 16 * chunk, v);
      tmem_ld_wait();
      const uint32_t q = chunk / kTsGroupsJ;
      const uint32_t cgrp = chunk % kTsGroupsJ;
      int32_t* base = acc_tile + static_cast<uint64_t>(q * kTsCols + cgrp * 16) * kTileRows;
#pragma unroll
      for (uint32_t c = 0; c < 16; ++c) {
        int32_t* p = base + PosToSample(c) * kTileRows;
        *p += static_cast<int32_t>(v[c]);
      }
    }
    tc_fence_before_sync();
  }
  __syncthreads();
  if (warp == 8) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace pl2
