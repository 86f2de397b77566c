// king_pairs_kernel.cuh - KING counts for an explicit LIST of sample pairs (`--king-table-subset`,
// CalcKingTableSubset, 2.0/plink2_matrix_calc.cc:3224; IncrKingSubset / IncrKingSubsetHomhom :2495-2741).
// Second-stage relationship screening on cohorts whose full N x N result does not fit: the pair list
// comes from a first pass on fewer variants.
//
// The staged block is transposed to sample-major 2-bit rows; one warp per pair streams the two rows
// (coalesced 256-byte loads) and keeps the five counts in registers: HBM/L2-bound, 2 * M/4 bytes per pair.
// Count semantics follow the reference's subset path: "1" = the FIRST sample of the listed pair.
#pragma once
#include "common.cuh"

namespace pl2 {

// raw[variant][pitch] (2-bit, variant-major) -> raw_t[sample][pitch_t], pitch_t = variants / 4 bytes.
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

// counts[pair][5] += {IBS0, HETHET, HET2HOM1, HET1HOM2, HOMHOM} over this block's variants.
static __global__ void __launch_bounds__(256) king_pairs_kernel(const uint8_t* __restrict__ raw_t, uint32_t pitch_t, uint32_t word_ct /* 64-bit words per row */, const uint32_t* __restrict__ pairs, uint64_t pair_ct, uint32_t* __restrict__ counts) {
  const uint64_t pair = static_cast<uint64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint32_t lane = threadIdx.x & 31;
  if (pair >= pair_ct) return;
  const uint64_t* a = reinterpret_cast<const uint64_t*>(raw_t + static_cast<uint64_t>(pairs[2 * pair]) * pitch_t);
  const uint64_t* b = reinterpret_cast<const uint64_t*>(raw_t + static_cast<uint64_t>(pairs[2 * pair + 1]) * pitch_t);
  constexpr uint64_t kLo = 0x5555555555555555ull;
  uint32_t ibs0 = 0, hethet = 0, het2hom1 = 0, het1hom2 = 0, homhom = 0;
  for (uint32_t w = lane; w < word_ct; w += 32) {
    const uint64_t x = __ldg(a + w), y = __ldg(b + w);
    const uint64_t lox = x & kLo, hix = (x >> 1) & kLo, loy = y & kLo, hiy = (y >> 1) & kLo;
    const uint64_t het1 = lox & ~hix, het2 = loy & ~hiy;  // code 1
    const uint64_t hom1 = ~lox & kLo, hom2 = ~loy & kLo;  // codes 0 and 2
    const uint64_t hh = hom1 & hom2;
    ibs0 += __popcll(hh & (hix ^ hiy));                   // opposite homozygotes
    hethet += __popcll(het1 & het2);
    het2hom1 += __popcll(hom1 & het2);
    het1hom2 += __popcll(hom2 & het1);
    homhom += __popcll(hh);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    ibs0 += __shfl_xor_sync(0xFFFFFFFFu, ibs0, o);
    hethet += __shfl_xor_sync(0xFFFFFFFFu, hethet, o);
    het2hom1 += __shfl_xor_sync(0xFFFFFFFFu, het2hom1, o);
    het1hom2 += __shfl_xor_sync(0xFFFFFFFFu, het1hom2, o);
    homhom += __shfl_xor_sync(0xFFFFFFFFu, homhom, o);
  }
  if (lane == 0) {
    uint32_t* c = counts + 5 * pair;
    c[0] += ibs0;
    c[1] += hethet;
    c[2] += het2hom1;
    c[3] += het1hom2;
    c[4] += homhom;
  }
}

}  // namespace pl2
