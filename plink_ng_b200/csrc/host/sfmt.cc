#include "sfmt.h"

#include <cmath>
#include <cstring>

namespace pl2host {

namespace {
constexpr uint32_t kMsk[4] = {0xdfffffefU, 0xddfecb7fU, 0xbffaffffU, 0xbffffff6U};
constexpr uint32_t kParity[4] = {0x00000001U, 0x00000000U, 0x00000000U, 0x13c9e684U};

inline void Lshift128(uint32_t out[4], const uint32_t in[4], int bytes) {
  const uint64_t th = (static_cast<uint64_t>(in[3]) << 32) | in[2];
  const uint64_t tl = (static_cast<uint64_t>(in[1]) << 32) | in[0];
  const int s = bytes * 8;
  const uint64_t oh = (th << s) | (tl >> (64 - s));
  const uint64_t ol = tl << s;
  out[1] = static_cast<uint32_t>(ol >> 32);
  out[0] = static_cast<uint32_t>(ol);
  out[3] = static_cast<uint32_t>(oh >> 32);
  out[2] = static_cast<uint32_t>(oh);
}
inline void Rshift128(uint32_t out[4], const uint32_t in[4], int bytes) {
  const uint64_t th = (static_cast<uint64_t>(in[3]) << 32) | in[2];
  const uint64_t tl = (static_cast<uint64_t>(in[1]) << 32) | in[0];
  const int s = bytes * 8;
  const uint64_t oh = th >> s;
  const uint64_t ol = (tl >> s) | (th << (64 - s));
  out[1] = static_cast<uint32_t>(ol >> 32);
  out[0] = static_cast<uint32_t>(ol);
  out[3] = static_cast<uint32_t>(oh >> 32);
  out[2] = static_cast<uint32_t>(oh);
}
inline uint32_t Func1(uint32_t x) { return (x ^ (x >> 27)) * 1664525U; }
inline uint32_t Func2(uint32_t x) { return (x ^ (x >> 27)) * 1566083941U; }
}  // namespace

void Sfmt19937::GenRandAll() {
  auto w = [&](int i) { return &s_[4 * i]; };
  auto recur = [&](uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d) {
    uint32_t x[4], y[4];
    Lshift128(x, a, kSl2);
    Rshift128(y, c, kSr2);
    for (int k = 0; k < 4; ++k) r[k] = a[k] ^ x[k] ^ ((b[k] >> kSr1) & kMsk[k]) ^ y[k] ^ (d[k] << kSl1);
  };
  uint32_t* r1 = w(kN - 2);
  uint32_t* r2 = w(kN - 1);
  int i = 0;
  for (; i < kN - kPos1; ++i) {
    recur(w(i), w(i), w(i + kPos1), r1, r2);
    r1 = r2;
    r2 = w(i);
  }
  for (; i < kN; ++i) {
    recur(w(i), w(i), w(i + kPos1 - kN), r1, r2);
    r1 = r2;
    r2 = w(i);
  }
}

void Sfmt19937::PeriodCertification() {
  uint32_t inner = 0;
  for (int i = 0; i < 4; ++i) inner ^= s_[i] & kParity[i];
  for (int i = 16; i > 0; i >>= 1) inner ^= inner >> i;
  if (inner & 1) return;
  for (int i = 0; i < 4; ++i) {
    uint32_t work = 1;
    for (int j = 0; j < 32; ++j) {
      if (work & kParity[i]) {
        s_[i] ^= work;
        return;
      }
      work <<= 1;
    }
  }
}

void Sfmt19937::InitGenRand(uint32_t seed) {
  s_[0] = seed;
  for (int i = 1; i < kN32; ++i) s_[i] = 1812433253U * (s_[i - 1] ^ (s_[i - 1] >> 30)) + static_cast<uint32_t>(i);
  idx_ = kN32;
  PeriodCertification();
}

void Sfmt19937::InitByArray(const uint32_t* key, int key_length) {
  constexpr int lag = 11, mid = (kN32 - lag) / 2;
  memset(s_, 0x8b, sizeof(s_));
  int count = (key_length + 1 > kN32) ? key_length + 1 : kN32;
  uint32_t r = Func1(s_[0] ^ s_[mid] ^ s_[kN32 - 1]);
  s_[mid] += r;
  r += static_cast<uint32_t>(key_length);
  s_[mid + lag] += r;
  s_[0] = r;
  --count;
  int i = 1, j = 0;
  for (; j < count && j < key_length; ++j) {
    r = Func1(s_[i] ^ s_[(i + mid) % kN32] ^ s_[(i + kN32 - 1) % kN32]);
    s_[(i + mid) % kN32] += r;
    r += key[j] + static_cast<uint32_t>(i);
    s_[(i + mid + lag) % kN32] += r;
    s_[i] = r;
    i = (i + 1) % kN32;
  }
  for (; j < count; ++j) {
    r = Func1(s_[i] ^ s_[(i + mid) % kN32] ^ s_[(i + kN32 - 1) % kN32]);
    s_[(i + mid) % kN32] += r;
    r += static_cast<uint32_t>(i);
    s_[(i + mid + lag) % kN32] += r;
    s_[i] = r;
    i = (i + 1) % kN32;
  }
  for (j = 0; j < kN32; ++j) {
    r = Func2(s_[i] + s_[(i + mid) % kN32] + s_[(i + kN32 - 1) % kN32]);
    s_[(i + mid) % kN32] ^= r;
    r -= static_cast<uint32_t>(i);
    s_[(i + mid + lag) % kN32] ^= r;
    s_[i] = r;
    i = (i + 1) % kN32;
  }
  idx_ = kN32;
  PeriodCertification();
}

uint32_t Sfmt19937::GenRandU32() {
  if (idx_ >= kN32) {
    GenRandAll();
    idx_ = 0;
  }
  return s_[idx_++];
}

void FillGaussian(uint64_t entry_pair_ct, uint32_t thread_ct, Sfmt19937* main_rng, double* dst) {
  const uint64_t max_useful = (entry_pair_ct + 262143) / 262144;
  if (thread_ct > max_useful) thread_ct = static_cast<uint32_t>(max_useful);
  if (!thread_ct) thread_ct = 1;
  // InitAllocSfmtpArr(thread_ct, use_main_sfmt_as_element_zero = 1, ...): streams 1.. are seeded by four
  // draws each from the main generator (plink2_random.cc:41-57)
  std::vector<Sfmt19937> extra(thread_ct > 1 ? thread_ct - 1 : 0);
  for (uint32_t t = 1; t < thread_ct; ++t) {
    uint32_t key[4];
    for (int k = 0; k < 4; ++k) key[k] = main_rng->GenRandU32();
    extra[t - 1].InitByArray(key, 4);
  }
  const double k2m32 = 1.0 / 4294967296.0;
  const double kPi = 3.1415926535897932;
  for (uint32_t t = 0; t < thread_ct; ++t) {
    Sfmt19937* rng = t ? &extra[t - 1] : main_rng;
    const uint64_t start = (static_cast<uint64_t>(t) * entry_pair_ct) / thread_ct;
    const uint64_t end = (static_cast<uint64_t>(t + 1) * entry_pair_ct) / thread_ct;
    double* out = dst + 2 * start;
    for (uint64_t p = start; p < end; ++p) {
      // RandNormal (plink2_random.cc:29-36): Box-Muller, returns the sin branch first
      const double u1 = (rng->GenRandU32() + 0.5) * k2m32;
      const double dxx = sqrt(-2 * log(u1));
      const double u2 = (rng->GenRandU32() + 0.5) * k2m32;
      const double dyy = (2 * kPi) * u2;
      *out++ = dxx * sin(dyy);
      *out++ = dxx * cos(dyy);
    }
  }
}

}  // namespace pl2host
