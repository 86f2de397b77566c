// sfmt.h - SFMT-19937 (Saito & Matsumoto, "SIMD-oriented Fast Mersenne Twister", MCQMC 2006), the
// generator plink2 seeds with --seed (2.0/plink2.cc:13091-13096; vendored third-party copy at
// 2.0/include/SFMT.{c,h}, v1.4, MEXP 19937).  Restated from the published recurrence so that
// `--pca approx` starts from the same Gaussian matrix as the reference (FillGaussianDArr,
// 2.0/plink2_random.cc:29-99).  Portable 32-bit-lane formulation (no SIMD).
#pragma once
#include <cstdint>
#include <vector>

namespace pl2host {

class Sfmt19937 {
 public:
  void InitGenRand(uint32_t seed);
  void InitByArray(const uint32_t* key, int key_length);
  uint32_t GenRandU32();

 private:
  static constexpr int kN = 156, kN32 = 624, kPos1 = 122, kSl1 = 18, kSl2 = 1, kSr1 = 11, kSr2 = 1;
  void GenRandAll();
  void PeriodCertification();
  uint32_t s_[kN32];
  int idx_ = kN32;
};

// FillGaussianDArr (2.0/plink2_random.cc:68-99): entry_pair_ct (sin, cos) Box-Muller pairs,
// sliced over min(thread_ct, ceil(pairs / 262144)) generator streams (stream 0 = *main).
void FillGaussian(uint64_t entry_pair_ct, uint32_t thread_ct, Sfmt19937* main_rng, double* dst);

}  // namespace pl2host
