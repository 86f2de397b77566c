// ld_kernels.cuh - device kernels of the --indep-pairwise path.
//
// Replaces, for every (second, first) variant pair that can share a window, the reference's
// ComputeIndepPairwiseR2Components -> DotprodWords / SumSsqWords / SumSsqNmWords
// (2.0/plink2_ld.cc:699-723, :235, :317, :578) and the r^2 test at :1085-1090.  The sequential
// greedy window walk (IndepPairwiseThread, :862-1109) stays on the host and only looks up the
// per-pair decision bits produced here.
//
// Per variant three bit planes over founders (word = 32 samples):
//   nm = non-missing, hom = genotype in {0,2}, hp = genotype 0 ("+1"; hom & ref2het of the reference)
// so with x in {+1,0,-1}:  dot = pc(hom_a&hom_b) - 2*pc(hom_a&hom_b&(hp_a^hp_b)),
//   ssq_b|a = pc(nm_a&hom_b), sum_b|a = 2*pc(nm_a&hp_b) - ssq_b|a, nm = pc(nm_a&nm_b)   (7 popcounts/word).
#pragma once
#include "common.cuh"
#include "cp_async.cuh"

namespace pl2 {

// ---- per-variant genotype counts {hom-REF, het, hom-ALT, missing}: GenoarrCountFreqsUnsafe
// (2.0/include/pgenlib_misc.cc:702); one warp per variant over the padded raw block.
static __global__ void __launch_bounds__(256) geno_counts_kernel(const uint8_t* __restrict__ raw, uint32_t pitch, uint32_t sample_ct, uint32_t sample_ct_padded, uint32_t variant_ct, uint32_t* __restrict__ counts) {
  const uint32_t v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint32_t lane = threadIdx.x & 31;
  if (v >= variant_ct) return;
  const uint64_t* row = reinterpret_cast<const uint64_t*>(raw + static_cast<uint64_t>(v) * pitch);
  const uint32_t words = pitch / 8;
  uint32_t n1 = 0, n2 = 0, n3 = 0;
  for (uint32_t w = lane; w < words; w += 32) {
    const uint64_t x = row[w];
    const uint64_t lo = x & 0x5555555555555555ull;
    const uint64_t hi = (x >> 1) & 0x5555555555555555ull;
    n1 += __popcll(lo & ~hi);
    n2 += __popcll(hi & ~lo);
    n3 += __popcll(lo & hi);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    n1 += __shfl_xor_sync(0xFFFFFFFFu, n1, o);
    n2 += __shfl_xor_sync(0xFFFFFFFFu, n2, o);
    n3 += __shfl_xor_sync(0xFFFFFFFFu, n3, o);
  }
  if (lane == 0) {
    n3 -= (sample_ct_padded - sample_ct);  // padding is coded "missing"
    counts[4ull * v + 0] = sample_ct - n1 - n2 - n3;
    counts[4ull * v + 1] = n1;
    counts[4ull * v + 2] = n2;
    counts[4ull * v + 3] = n3;
  }
}

// ---- sample gather for the sex-chromosome / haploid forms of the LD block (2.0/plink2_ld.cc:1356-1389):
// output sample t of every variant row = input sample (map[t] & 0x7FFFFFFF), with a het call turned into
// "missing" when bit 31 of map[t] is set (SetHetMissing).  A sample listed twice carries weight 2 in every sum
// of the pair sextuple, which is exactly how chrX counts nonmales (:982-998).  One thread per output byte.
static __global__ void __launch_bounds__(256) geno_gather_kernel(const uint8_t* __restrict__ in, uint64_t in_pitch, uint8_t* __restrict__ out, uint32_t out_pitch, const uint32_t* __restrict__ map, uint32_t out_sample_ct) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= out_pitch) return;
  const uint8_t* row = in + static_cast<uint64_t>(blockIdx.y) * in_pitch;
  uint32_t byte = 0;
#pragma unroll
  for (uint32_t q = 0; q < 4; ++q) {
    const uint32_t t = 4 * b + q;
    uint32_t code = 3;
    if (t < out_sample_ct) {
      const uint32_t mv = map[t], s = mv & 0x7FFFFFFFu;
      code = (row[s >> 2] >> (2 * (s & 3))) & 3;
      if ((mv >> 31) && code == 1) code = 3;
    }
    byte |= code << (2 * q);
  }
  out[static_cast<uint64_t>(blockIdx.y) * out_pitch + b] = static_cast<uint8_t>(byte);
}

__device__ __forceinline__ uint32_t compact_even_bits(uint64_t x) {
  x &= 0x5555555555555555ull;
  x = (x | (x >> 1)) & 0x3333333333333333ull;
  x = (x | (x >> 2)) & 0x0F0F0F0F0F0F0F0Full;
  x = (x | (x >> 4)) & 0x00FF00FF00FF00FFull;
  x = (x | (x >> 8)) & 0x0000FFFF0000FFFFull;
  x = (x | (x >> 16)) & 0x00000000FFFFFFFFull;
  return static_cast<uint32_t>(x);
}

// ---- raw rows -> planes[p][kw][v] (p in {nm, hom, hp}); 32 variants x 32 sample-words per CTA via a
// shared-memory transpose so both the row reads and the plane writes are coalesced.
// variants_in_chunk is a multiple of 64; word_ct = sample_ct_padded / 32.
static __global__ void __launch_bounds__(1024) ld_split_kernel(const uint8_t* __restrict__ raw, uint32_t pitch, uint32_t word_ct, uint32_t variants_in_chunk, uint32_t* __restrict__ planes) {
  __shared__ uint32_t s[3][32][33];
  const uint32_t v0 = blockIdx.x * 32, kw0 = blockIdx.y * 32;
  const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  {
    const uint32_t v = v0 + ty, kw = kw0 + tx;
    uint32_t nm = 0, hom = 0, hp = 0;
    if (kw < word_ct) {
      const uint64_t w = *reinterpret_cast<const uint64_t*>(raw + static_cast<uint64_t>(v) * pitch + static_cast<uint64_t>(kw) * 8);
      const uint32_t lo = compact_even_bits(w);
      const uint32_t hi = compact_even_bits(w >> 1);
      nm = ~(lo & hi);
      hom = ~lo;
      hp = ~(lo | hi);
    }
    s[0][tx][ty] = nm;
    s[1][tx][ty] = hom;
    s[2][tx][ty] = hp;
  }
  __syncthreads();
  const uint32_t kw = kw0 + ty;
  if (kw < word_ct) {
    const uint64_t plane_words = static_cast<uint64_t>(word_ct) * variants_in_chunk;
    const uint64_t off = static_cast<uint64_t>(kw) * variants_in_chunk + v0 + tx;
    planes[off] = s[0][ty][tx];
    planes[plane_words + off] = s[1][ty][tx];
    planes[2 * plane_words + off] = s[2][ty][tx];
  }
}

// ---- banded pair kernel: CTA = 64 "second" variants (a) x 64 "first" variants (b); thread = 4 x 4
// pairs x 7 popcount accumulators; sample words staged by cp.async double buffering.
// flags[(a - a_out0) * band + (a - b - 1)] = (cov12^2 > thresh * var1 * var2) for 0 < a - b <= band.
constexpr uint32_t kLdKw = 12;  // sample words per smem chunk (word_ct is a multiple of 12)

static __global__ void __launch_bounds__(256, 1)
ld_band_kernel(const uint32_t* __restrict__ planes, uint32_t word_ct, uint32_t variants_in_chunk, uint32_t chunk_lo /* global index of plane column 0 */, uint32_t a_out0, uint32_t a_out1, uint32_t band, double thresh, uint8_t* __restrict__ flags) {
  __shared__ __align__(16) uint32_t s_a[2][3][kLdKw][64];
  __shared__ __align__(16) uint32_t s_b[2][3][kLdKw][64];
  const uint32_t a_start = a_out0 + blockIdx.x * 64;  // global variant index, multiple of 64
  const int64_t b_start_s = static_cast<int64_t>(a_start) - 64ll * blockIdx.y;
  if (b_start_s < static_cast<int64_t>(chunk_lo)) return;
  const uint32_t b_start = static_cast<uint32_t>(b_start_s);
  if (a_start - b_start > band + 63) return;  // no pair of this tile is within the band
  const uint32_t tid = threadIdx.x;
  const uint32_t ay = tid >> 4, bx = tid & 15;
  const uint64_t plane_words = static_cast<uint64_t>(word_ct) * variants_in_chunk;
  const uint32_t a_col = a_start - chunk_lo, b_col = b_start - chunk_lo;

  uint32_t c_nm[4][4], c_hh[4][4], c_x[4][4], c_qb[4][4], c_pb[4][4], c_qa[4][4], c_pa[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c_nm[i][j] = c_hh[i][j] = c_x[i][j] = c_qb[i][j] = c_pb[i][j] = c_qa[i][j] = c_pa[i][j] = 0;

  const uint32_t chunk_ct = word_ct / kLdKw;
  auto issue = [&](uint32_t chunk, uint32_t buf) {
    for (uint32_t i = tid; i < 2 * 3 * kLdKw * 16; i += 256) {
      const uint32_t side = i / (3 * kLdKw * 16), r = i % (3 * kLdKw * 16);
      const uint32_t p = r / (kLdKw * 16), r2 = r % (kLdKw * 16), kk = r2 / 16, seg = r2 % 16;
      const uint32_t* src = planes + p * plane_words + static_cast<uint64_t>(chunk * kLdKw + kk) * variants_in_chunk + (side ? b_col : a_col) + seg * 4;
      cp_async16(side ? &s_b[buf][p][kk][seg * 4] : &s_a[buf][p][kk][seg * 4], src);
    }
    cp_async_commit();
  };
  if (chunk_ct) issue(0, 0);
  for (uint32_t chunk = 0; chunk < chunk_ct; ++chunk) {
    const uint32_t buf = chunk & 1;
    if (chunk + 1 < chunk_ct) {
      issue(chunk + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
#pragma unroll 2
    for (uint32_t kk = 0; kk < kLdKw; ++kk) {
      const uint4 a_nm4 = *reinterpret_cast<const uint4*>(&s_a[buf][0][kk][4 * ay]);
      const uint4 a_hom4 = *reinterpret_cast<const uint4*>(&s_a[buf][1][kk][4 * ay]);
      const uint4 a_hp4 = *reinterpret_cast<const uint4*>(&s_a[buf][2][kk][4 * ay]);
      const uint4 b_nm4 = *reinterpret_cast<const uint4*>(&s_b[buf][0][kk][4 * bx]);
      const uint4 b_hom4 = *reinterpret_cast<const uint4*>(&s_b[buf][1][kk][4 * bx]);
      const uint4 b_hp4 = *reinterpret_cast<const uint4*>(&s_b[buf][2][kk][4 * bx]);
      const uint32_t a_nm[4] = {a_nm4.x, a_nm4.y, a_nm4.z, a_nm4.w}, a_hom[4] = {a_hom4.x, a_hom4.y, a_hom4.z, a_hom4.w}, a_hp[4] = {a_hp4.x, a_hp4.y, a_hp4.z, a_hp4.w};
      const uint32_t b_nm[4] = {b_nm4.x, b_nm4.y, b_nm4.z, b_nm4.w}, b_hom[4] = {b_hom4.x, b_hom4.y, b_hom4.z, b_hom4.w}, b_hp[4] = {b_hp4.x, b_hp4.y, b_hp4.z, b_hp4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t hh = a_hom[i] & b_hom[j];
          c_nm[i][j] += __popc(a_nm[i] & b_nm[j]);
          c_hh[i][j] += __popc(hh);
          c_x[i][j] += __popc(hh & (a_hp[i] ^ b_hp[j]));
          c_qb[i][j] += __popc(a_nm[i] & b_hom[j]);
          c_pb[i][j] += __popc(a_nm[i] & b_hp[j]);
          c_qa[i][j] += __popc(b_nm[j] & a_hom[i]);
          c_pa[i][j] += __popc(b_nm[j] & a_hp[i]);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t a = a_start + 4 * ay + i;
    if (a >= a_out1) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t b = b_start + 4 * bx + j;
      if (b >= a || a - b > band) continue;
      const int64_t nm = c_nm[i][j];
      const int64_t dot = static_cast<int64_t>(c_hh[i][j]) - 2 * static_cast<int64_t>(c_x[i][j]);
      const int64_t q_b = c_qb[i][j], s_b = 2 * static_cast<int64_t>(c_pb[i][j]) - q_b;  // first
      const int64_t q_a = c_qa[i][j], s_a = 2 * static_cast<int64_t>(c_pa[i][j]) - q_a;  // second
      // plink2_ld.cc:1085-1090; int64 -> double casts, unfused left-to-right multiplies
      const double cov12 = static_cast<double>(dot * nm - s_b * s_a);
      const double var1 = static_cast<double>(q_b * nm - s_b * s_b);
      const double var2 = static_cast<double>(q_a * nm - s_a * s_a);
      const bool over = __dmul_rn(cov12, cov12) > __dmul_rn(__dmul_rn(thresh, var1), var2);
      flags[static_cast<uint64_t>(a - a_out0) * band + (a - b - 1)] = over ? 1 : 0;
    }
  }
}

}  // namespace pl2
