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
#include "geno_tile.cuh"
#include "umma.cuh"

namespace pl2 {

constexpr uint32_t kTsGroupsJ = kTsCols / 16;
constexpr uint32_t kTsAccCols = 5 * kTsCols;   // 400
constexpr uint32_t kTsTileAccWords = kTsAccCols * kTileRows;
constexpr uint32_t kTsASlots = 4;
constexpr uint32_t kTsASlotCols = 24;
constexpr uint32_t kTsStagesJ = 4;
constexpr uint32_t kTsLboJ = (3 * kTsCols / 16) * kCoreBytes + 64;  // 1984: +64 keeps the K-permuted rows bank-conflict free
constexpr uint32_t kTsStageBytesJ = (kTsKcJ / 8) * kTsLboJ;         // 15872
constexpr uint32_t kTsSmemBytes = kTsStagesJ * kTsStageBytesJ + 1024;
constexpr uint32_t kTsRowWarps = 8;
constexpr uint32_t kTsColWarps = 10;           // 5 words x 64 variants per stage
constexpr uint32_t kTsThreads = 32 * (kTsRowWarps + kTsColWarps + 1);  // + the UMMA issuer warp

__global__ void __launch_bounds__(kTsThreads, 1)
king_ts_kernel(const uint8_t* __restrict__ raw_j, const uint8_t* __restrict__ raw_i, uint32_t variant_ct_padded /* multiple of 256 */, const uint32_t* __restrict__ tile_order, const uint32_t* __restrict__ tile_rt, const uint32_t* __restrict__ tile_tc, int32_t* __restrict__ raw_acc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full_a[kTsASlots];
  __shared__ __align__(8) uint64_t bar_empty_a[kTsASlots];
  __shared__ __align__(8) uint64_t bar_full_b[kTsStagesJ];
  __shared__ __align__(8) uint64_t bar_empty_b[kTsStagesJ];
  __shared__ __align__(8) uint64_t bar_acc;
  __shared__ uint32_t tmem_base_slot;

  const uint32_t tid = threadIdx.x;
  const uint32_t warp = uniform_warp_idx();
  const uint32_t lane = tid & 31;
  const uint32_t tile = tile_order[blockIdx.x];
  const uint32_t rt = tile_rt[tile];
  const uint32_t ct = tile_tc[tile];
  const uint32_t stage_iters = variant_ct_padded / kTsKcJ;
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;

  if (tid == 0) {
    for (uint32_t s = 0; s < kTsASlots; ++s) {
      mbar_init(&bar_full_a[s], 4);             // one arrival per row-side warp of the owning group
      mbar_init(&bar_empty_a[s], 1);
    }
    for (uint32_t s = 0; s < kTsStagesJ; ++s) {
      mbar_init(&bar_full_b[s], kTsColWarps);
      mbar_init(&bar_empty_b[s], 1);
    }
    mbar_init(&bar_acc, 1);
    mbar_fence_init();
  }
  if (warp == kTsRowWarps + kTsColWarps) tmem_alloc<512>(&tmem_base_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;

  const uint32_t thread_zero = tid * (variant_ct_padded >> 31);  // 0 (a batch never has 2^31 variants)
  const uint32_t tab_t = table_reg(kTabHet, thread_zero), tab_h = table_reg(kTabHom, thread_zero), tab_s = table_reg(kTabSgn, thread_zero);

  if (warp < kTsRowWarps) {
    // ---------------- row-side producers: 2-bit words -> registers -> tensor memory ----------------
    // Two groups of four warps (group = warp / 4); group g owns k-steps ks = 2 n + g and the A slots
    // ks % 4 in {g, g + 2}.  Thread = TMEM lane = sample 128 rt + 32 (warp % 4) + lane.  The words of
    // the next k-step are expanded while the UMMAs of the previous ones run.
    const uint32_t grp = warp >> 2;
    const uint32_t lq = warp & 3;
    const uint32_t row = 32 * lq + lane;
    const uint8_t* src_i = raw_i + static_cast<uint64_t>(rt) * (2 * stage_iters) * 1024 + row * 8;
    const uint32_t taddr_lane = tmem_base + ((32u * lq) << 16) + kTsAccCols;
    auto load_i = [&](uint32_t n) -> uint2 {
      return (n < stage_iters) ? __ldg(reinterpret_cast<const uint2*>(src_i + 1024ull * (2 * n + grp))) : make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
    };
    struct ExpI {
      uint32_t v[3][8];
    };
    auto expand_i = [&](const uint2& w) -> ExpI {
      ExpI e;
      const Sel4 s0 = make_selectors(w.x), s1 = make_selectors(w.y);
      const uint32_t tabs[3] = {tab_t, tab_h, tab_s};
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        const uint4 a = expand16(tabs[p], s0), b = expand16(tabs[p], s1);
        e.v[p][0] = a.x; e.v[p][1] = a.y; e.v[p][2] = a.z; e.v[p][3] = a.w;
        e.v[p][4] = b.x; e.v[p][5] = b.y; e.v[p][6] = b.z; e.v[p][7] = b.w;
      }
      return e;
    };
    constexpr uint32_t kLa = 4;
    uint2 pre_i[kLa];
#pragma unroll
    for (uint32_t d = 0; d < kLa; ++d) pre_i[d] = load_i(d);
    ExpI cur = expand_i(pre_i[0]);
    for (uint32_t n0 = 0; n0 < stage_iters; n0 += kLa) {  // stage_iters is a multiple of 4
#pragma unroll
      for (uint32_t d = 0; d < kLa; ++d) {
        const uint32_t n = n0 + d;
        const uint32_t ks = 2 * n + grp;
        const uint32_t slot = ks % kTsASlots;
        pre_i[d] = load_i(n + kLa);
        mbar_wait(&bar_empty_a[slot], ((ks / kTsASlots) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t ta = taddr_lane + slot * kTsASlotCols;
        tmem_st8(ta, cur.v[0]);
        tmem_st8(ta + 8, cur.v[1]);
        tmem_st8(ta + 16, cur.v[2]);
        tmem_st_wait();
        tc_fence_before_sync();
        mbar_arrive_warp(&bar_full_a[slot], lane);
        cur = expand_i(pre_i[(d + 1) % kLa]);
      }
    }
  } else if (warp < kTsRowWarps + kTsColWarps) {
    // ---------------- column-side producers: 2-bit words -> int8 planes in shared memory ----------------
    // Thread = (word w of the 20-byte row, variant k of the 64-variant stage).
    const uint32_t t = tid - 32 * kTsRowWarps;     // 0..319
    const uint32_t k = t & 63;
    const uint32_t w = t >> 6;
    const uint8_t* src_j = raw_j + static_cast<uint64_t>(ct) * stage_iters * (kTsKcJ * 20) + k * 20 + 4 * w;
    // K rows are stored in the PRMT position order of the row side (geno_expand.cuh): variant k of a
    // 16-variant group sits at row SampleToPos(k % 16)
    const uint32_t kpos = (k & ~15u) + SampleToPos(k & 15u);
    const uint32_t dst_k = (kpos >> 3) * kTsLboJ + (kpos & 7) * 16 + w * kCoreBytes;
    auto load_j = [&](uint32_t it) -> uint32_t {
      return (it < stage_iters) ? __ldg(reinterpret_cast<const uint32_t*>(src_j + static_cast<uint64_t>(it) * (kTsKcJ * 20))) : 0xFFFFFFFFu;
    };
    constexpr uint32_t kLa = 4;
    uint32_t pre_j[kLa];
#pragma unroll
    for (uint32_t d = 0; d < kLa; ++d) pre_j[d] = load_j(d);
    for (uint32_t it0 = 0; it0 < stage_iters; it0 += kLa) {
#pragma unroll
      for (uint32_t d = 0; d < kLa; ++d) {
        const uint32_t it = it0 + d;
        const Sel4 sel = make_selectors(pre_j[d]);
        pre_j[d] = load_j(it + kLa);
        const uint4 vt = expand16(tab_t, sel), vh = expand16(tab_h, sel), vs = expand16(tab_s, sel);
        const uint32_t sb = d % kTsStagesJ;
        mbar_wait(&bar_empty_b[sb], ((it / kTsStagesJ) & 1) ^ 1);
        const uint32_t a0 = smem_base + sb * kTsStageBytesJ + dst_k;
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a0), "r"(vt.x), "r"(vt.y), "r"(vt.z), "r"(vt.w) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a0 + kTsGroupsJ * kCoreBytes), "r"(vh.x), "r"(vh.y), "r"(vh.z), "r"(vh.w) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a0 + 2 * kTsGroupsJ * kCoreBytes), "r"(vs.x), "r"(vs.y), "r"(vs.z), "r"(vs.w) : "memory");
        fence_proxy_async_smem();
        mbar_arrive_warp(&bar_full_b[sb], lane);
      }
    }
  } else {
    // ---------------- UMMA issuer: whole warp loops, one elected lane issues (umma.cuh) ----------------
    // One outer iteration = the 4 shared-memory stages = 8 k-steps = two rounds of the 4 A slots, so
    // every slot index, A parity and descriptor offset is a compile-time constant.
    static_assert(kTsStagesJ == 4 && kTsASlots == 4, "issuer unrolling assumes 4 stages / 4 slots");
    constexpr uint32_t idesc_n160 = make_idesc_i8(128, 2 * kTsCols, false, true);
    constexpr uint32_t idesc_n80 = make_idesc_i8(128, kTsCols, false, true);
    const uint32_t tmem_u = uniform_u32(tmem_base);
    const uint64_t desc0 = make_smem_desc(smem_base, kTsLboJ, kCoreBytes);
    for (uint32_t it0 = 0; it0 < stage_iters; it0 += kTsStagesJ) {
      const uint32_t ph_b = (it0 / kTsStagesJ) & 1;
#pragma unroll
      for (uint32_t sb = 0; sb < kTsStagesJ; ++sb) {
        mbar_wait(&bar_full_b[sb], ph_b);
#pragma unroll
        for (uint32_t kk = 0; kk < 2; ++kk) {
          const uint32_t kq = 2 * sb + kk;            // k-step inside the outer iteration
          const uint32_t slot = kq % kTsASlots;
          mbar_wait(&bar_full_a[slot], (kq / kTsASlots) & 1);
          tc_fence_after_sync();
          if (elect_one_sync()) {
            const uint32_t acc = (it0 | kq) ? 1u : 0u;
            const uint64_t b_th = desc0 + ((sb * kTsStageBytesJ + kk * 4 * kTsLboJ) >> 4);
            const uint64_t b_s = b_th + ((2 * kTsGroupsJ * kCoreBytes) >> 4);
            const uint32_t ta = tmem_u + kTsAccCols + slot * kTsASlotCols;
            umma_i8_ts(tmem_u + 0, ta, b_th, idesc_n160, acc);
            umma_i8_ts(tmem_u + 2 * kTsCols, ta + 8, b_th, idesc_n160, acc);
            umma_i8_ts(tmem_u + 4 * kTsCols, ta + 16, b_s, idesc_n80, acc);
            umma_commit(&bar_empty_a[slot]);
            if (kk == 1) umma_commit(&bar_empty_b[sb]);
          }
          __syncwarp();
        }
      }
    }
    if (elect_one_sync()) umma_commit(&bar_acc);
    __syncwarp();
  }

  if (warp < kTsRowWarps) {
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
  if (warp == kTsRowWarps + kTsColWarps) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace pl2
