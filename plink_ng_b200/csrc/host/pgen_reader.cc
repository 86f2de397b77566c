#include "pgen_reader.h"

#include <fcntl.h>
#include <immintrin.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <thread>

namespace pl2host {

namespace {

inline bool GetVarint(const uint8_t** pp, const uint8_t* end, uint32_t* out) {
  uint32_t v = 0;
  for (uint32_t shift = 0; shift < 35; shift += 7) {
    if (*pp >= end) return false;
    const uint32_t b = *(*pp)++;
    v |= (b & 127) << shift;
    if (!(b & 128)) {
      *out = v;
      return true;
    }
  }
  return false;
}

inline uint32_t LoadLe(const uint8_t* p, uint32_t nbytes) {
  uint32_t v = 0;
  for (uint32_t i = 0; i < nbytes; ++i) v |= static_cast<uint32_t>(p[i]) << (8 * i);
  return v;
}

inline void SetCode(uint64_t* genovec, uint32_t s, uint32_t code) {
  const uint32_t sh = 2 * (s & 31);
  genovec[s >> 5] = (genovec[s >> 5] & ~(3ull << sh)) | (static_cast<uint64_t>(code) << sh);
}

inline void ZeroTrailing(uint64_t* genovec, uint32_t n) {
  if (n & 31) genovec[n >> 5] &= (1ull << (2 * (n & 31))) - 1;
}

}  // namespace

void PgenReader::Close() {
  if (map_) munmap(const_cast<uint8_t*>(map_), map_len_);
  map_ = nullptr;
  if (fd_ >= 0) close(fd_);
  fd_ = -1;
}

bool PgenReader::Open(const std::string& path, uint32_t raw_sample_ct, uint32_t raw_variant_ct, std::string* err) {
  Close();
  fd_ = open(path.c_str(), O_RDONLY);
  if (fd_ < 0) {
    *err = "Failed to open " + path + ".";
    return false;
  }
  struct stat st;
  if (fstat(fd_, &st) || st.st_size < 3) {
    *err = path + " is too small to be a .bed/.pgen file.";
    return false;
  }
  map_len_ = static_cast<uint64_t>(st.st_size);
  void* m = mmap(nullptr, map_len_, PROT_READ, MAP_PRIVATE, fd_, 0);
  if (m == MAP_FAILED) {
    *err = "mmap failed on " + path + ".";
    return false;
  }
  map_ = static_cast<const uint8_t*>(m);
  madvise(m, map_len_, MADV_SEQUENTIAL);
  if (map_[0] != 0x6c || map_[1] != 0x1b) {
    *err = path + " is not a .bed/.pgen file (bad magic number).";
    return false;
  }
  mode_ = map_[2];
  if (mode_ == 0x01) {
    if (!raw_sample_ct) {
      *err = "PLINK 1 .bed needs the sample count from the .fam file.";
      return false;
    }
    raw_sample_ct_ = raw_sample_ct;
    fixed_bpv_ = (raw_sample_ct + 3) / 4;
    fixed_start_ = 3;
    const uint64_t m_ct = (map_len_ - 3) / fixed_bpv_;
    if ((map_len_ - 3) % fixed_bpv_ || (raw_variant_ct && m_ct != raw_variant_ct)) {
      *err = "Unexpected .bed file size (wrong .fam/.bim, or sample-major .bed?).";
      return false;
    }
    raw_variant_ct_ = static_cast<uint32_t>(m_ct);
  } else if (mode_ == 0x02 || mode_ == 0x10) {
    if (map_len_ < 12) {
      *err = path + ": truncated .pgen header.";
      return false;
    }
    raw_variant_ct_ = LoadLe(map_ + 3, 4);
    raw_sample_ct_ = LoadLe(map_ + 7, 4);
    if ((raw_sample_ct && raw_sample_ct != raw_sample_ct_) || (raw_variant_ct && raw_variant_ct != raw_variant_ct_)) {
      *err = ".pgen header dimensions do not match the .psam/.pvar files.";
      return false;
    }
    const uint8_t fb = map_[11];
    nonref_mode_ = fb >> 6;
    if (mode_ == 0x02) {
      fixed_bpv_ = (raw_sample_ct_ + 3) / 4;
      fixed_start_ = 12;
      // optional provisional-REF bitarray follows the header when bits 6-7 == 3
      if ((fb >> 6) == 3) fixed_start_ += (raw_variant_ct_ + 7) / 8;
      if (fixed_start_ + static_cast<uint64_t>(raw_variant_ct_) * fixed_bpv_ > map_len_) {
        *err = path + ": truncated fixed-width .pgen.";
        return false;
      }
    } else {
      const uint32_t fmt = fb & 15;
      if (fmt > 7) {
        *err = path + ": unsupported .pgen header format byte.";
        return false;
      }
      const uint32_t type_bits = (fmt >= 4) ? 8 : 4;
      const uint32_t len_bytes = (fmt & 3) + 1;
      const uint32_t ac_bytes = (fb >> 4) & 3;
      const uint32_t nonref_mode = fb >> 6;
      const uint32_t block_ct = (raw_variant_ct_ + 65535) / 65536;
      uint64_t pos = 12;
      if (pos + 8ull * block_ct > map_len_) {
        *err = path + ": truncated .pgen block index.";
        return false;
      }
      std::vector<uint64_t> block_off(block_ct);
      memcpy(block_off.data(), map_ + pos, 8ull * block_ct);
      pos += 8ull * block_ct;
      vrtypes_.resize(raw_variant_ct_);
      rec_off_.resize(static_cast<uint64_t>(raw_variant_ct_) + 1);
      for (uint32_t b = 0; b < block_ct; ++b) {
        const uint32_t v0 = b * 65536;
        const uint32_t nb = std::min<uint32_t>(65536, raw_variant_ct_ - v0);
        const uint64_t type_bytes = (static_cast<uint64_t>(nb) * type_bits + 7) / 8;
        if (pos + type_bytes + static_cast<uint64_t>(nb) * len_bytes > map_len_) {
          *err = path + ": truncated .pgen header body.";
          return false;
        }
        for (uint32_t i = 0; i < nb; ++i) {
          vrtypes_[v0 + i] = (type_bits == 8) ? map_[pos + i] : ((map_[pos + (i >> 1)] >> (4 * (i & 1))) & 15);
        }
        pos += type_bytes;
        uint64_t off = block_off[b];
        for (uint32_t i = 0; i < nb; ++i) {
          rec_off_[v0 + i] = off;
          off += LoadLe(map_ + pos + static_cast<uint64_t>(i) * len_bytes, len_bytes);
        }
        if (b + 1 == block_ct) rec_off_[raw_variant_ct_] = off;
        pos += static_cast<uint64_t>(nb) * len_bytes;
        if (ac_bytes) pos += static_cast<uint64_t>(nb) * ac_bytes;
        if (nonref_mode == 3) pos += (nb + 7) / 8;
      }
      // blocks are contiguous, so the end of the last record of block b is the start of block b+1
      if (rec_off_[raw_variant_ct_] > map_len_) {
        *err = path + ": variant records extend past the end of the file.";
        return false;
      }
    }
  } else {
    char buf[96];
    snprintf(buf, sizeof(buf), ": storage mode 0x%02x is not supported (hard-call modes 0x01, 0x02, 0x10 only).", mode_);
    *err = path + buf;
    return false;
  }
  const uint32_t words = WordsFor(raw_sample_ct_);
  state_.ldbase.assign(words, 0);
  state_.scratch.assign(words, 0);
  state_.ldbase_vidx = 0xFFFFFFFFu;
  return true;
}

bool PgenReader::ReadRecordBytes(uint32_t vidx, const uint8_t** rec, uint32_t* len, std::string* err) const {
  if (vidx >= raw_variant_ct_) {
    *err = "variant index out of range";
    return false;
  }
  if (mode_ == 0x10) {
    // within a block the next record starts where this one ends; across blocks use the next block offset
    const uint64_t a = rec_off_[vidx];
    uint64_t b = rec_off_[vidx + 1];
    if (b < a) b = map_len_;  // (never for well-formed files)
    *rec = map_ + a;
    *len = static_cast<uint32_t>(b - a);
  } else {
    *rec = map_ + fixed_start_ + static_cast<uint64_t>(vidx) * fixed_bpv_;
    *len = fixed_bpv_;
  }
  return true;
}

// Difflist (pgen_spec.tex:354-400).  with_values: patch (sample, 2-bit value) pairs; otherwise
// set every listed sample to fixed_value.
bool PgenReader::ParseDifflistAndApply(const uint8_t* p, const uint8_t* end, bool with_values, uint64_t* genovec, uint32_t fixed_value, std::string* err, const uint8_t** after) const {
  uint32_t len;
  if (!GetVarint(&p, end, &len)) goto malformed;
  if (!len) {
    *after = p;
    return true;
  }
  {
    const uint32_t n = raw_sample_ct_;
    if (len > n) goto malformed;
    const uint32_t id_bytes = (n <= 256) ? 1 : ((n <= 65536) ? 2 : ((n <= 16777216) ? 3 : 4));
    const uint32_t group_ct = (len + 63) / 64;
    const uint8_t* first_ids = p;
    p += static_cast<uint64_t>(group_ct) * id_bytes;
    p += group_ct - 1;  // group byte sizes: only needed for random access
    const uint8_t* values = p;
    if (with_values) p += (len + 3) / 4;
    if (p > end) goto malformed;
    uint32_t sample = 0;
    for (uint32_t k = 0; k < len; ++k) {
      if (!(k & 63)) {
        sample = LoadLe(first_ids + static_cast<uint64_t>(k >> 6) * id_bytes, id_bytes);
      } else {
        uint32_t delta;
        if (!GetVarint(&p, end, &delta)) goto malformed;
        sample += delta;
      }
      if (sample >= n) goto malformed;
      const uint32_t val = with_values ? ((values[k >> 2] >> (2 * (k & 3))) & 3) : fixed_value;
      SetCode(genovec, sample, val);
    }
    *after = p;
    return true;
  }
malformed:
  *err = "malformed difflist in .pgen variant record";
  return false;
}

bool PgenReader::DecodeRecord(DecodeState* st, uint32_t vidx, uint64_t* dst, std::string* err, bool as_ld_base) const {
  const uint8_t* rec;
  uint32_t len;
  if (!ReadRecordBytes(vidx, &rec, &len, err)) return false;
  const uint32_t n = raw_sample_ct_;
  const uint32_t words = WordsFor(n);
  const uint32_t bpv = (n + 3) / 4;
  if (mode_ == 0x01) {
    // PLINK 1 coding 0 hom-ALT, 1 missing, 2 het, 3 hom-REF -> 2, 3, 1, 0 (pgen_spec.tex:436-441)
    dst[words - 1] = 0;
    memcpy(dst, rec, bpv);
    for (uint32_t w = 0; w < words; ++w) {
      const uint64_t x = dst[w];
      const uint64_t lo = x & 0x5555555555555555ull, hi = (x >> 1) & 0x5555555555555555ull;
      // new_hi = ~hi_old ... derive: 00->10, 01->11, 10->01, 11->00
      const uint64_t nhi = (~hi) & 0x5555555555555555ull;          // high bit set for old 00, 01
      const uint64_t nlo = (lo ^ hi) & 0x5555555555555555ull;      // low bit set for old 01, 10
      dst[w] = (nhi << 1) | nlo;
    }
    ZeroTrailing(dst, n);
    return true;
  }
  const uint32_t vrtype = (mode_ == 0x02) ? 0 : vrtypes_[vidx];
  if ((vrtype & 8) && !as_ld_base) {
    *err = "multiallelic variant records are not supported by the pairwise-genotype commands (drop them with --max-alleles 2, or split them first)";
    return false;
  }
  const uint8_t* end = rec + len;
  const uint32_t t = vrtype & 7;
  const uint8_t* after = rec;
  switch (t) {
    case 0:
      if (len < bpv) {
        *err = "truncated .pgen variant record";
        return false;
      }
      dst[words - 1] = 0;
      memcpy(dst, rec, bpv);
      break;
    case 1: {
      const uint32_t head = 1 + (n + 7) / 8;
      if (len < head) {
        *err = "truncated .pgen variant record";
        return false;
      }
      const uint32_t low = rec[0] >> 2, delta = rec[0] & 3;
      const uint64_t base = static_cast<uint64_t>(low) * 0x5555555555555555ull;
      const uint8_t* bits = rec + 1;
      for (uint32_t w = 0; w < words; ++w) {
        uint32_t b32 = 0;
        const uint32_t byte0 = w * 4;
        for (uint32_t i = 0; i < 4 && byte0 + i < (n + 7) / 8; ++i) b32 |= static_cast<uint32_t>(bits[byte0 + i]) << (8 * i);
        // spread 32 bits to the even positions of 64 bits
        const uint64_t spread = _pdep_u64(b32, 0x5555555555555555ull);
        dst[w] = base + static_cast<uint64_t>(delta) * spread;
      }
      if (!ParseDifflistAndApply(rec + head, end, true, dst, 0, err, &after)) return false;
      break;
    }
    case 2:
    case 3: {
      // base = previous variant whose main track is not LD-compressed
      uint32_t b = vidx;
      do {
        if (!b) {
          *err = "LD-compressed .pgen record without a base variant";
          return false;
        }
        --b;
      } while ((vrtypes_[b] & 6) == 2);
      if (st->ldbase_vidx != b) {
        if (!DecodeRecord(st, b, st->ldbase.data(), err, true)) return false;
        st->ldbase_vidx = b;
      }
      memcpy(dst, st->ldbase.data(), words * 8ull);
      if (!ParseDifflistAndApply(rec, end, true, dst, 0, err, &after)) return false;
      if (t == 3) {
        for (uint32_t w = 0; w < words; ++w) dst[w] ^= ((~dst[w]) & 0x5555555555555555ull) << 1;  // swap 0 <-> 2
      }
      break;
    }
    case 4:
    case 6:
    case 7: {
      const uint64_t fill = (t == 4) ? 0 : ((t == 6) ? 0xAAAAAAAAAAAAAAAAull : ~0ull);
      for (uint32_t w = 0; w < words; ++w) dst[w] = fill;
      if (!ParseDifflistAndApply(rec, end, true, dst, 0, err, &after)) return false;
      break;
    }
    case 5:
      // all hom-REF, zero-length record (2.0/include/pgenlib_read.cc:2741-2743, :2996-3000)
      for (uint32_t w = 0; w < words; ++w) dst[w] = 0;
      break;
    default:
      *err = "invalid .pgen variant record type";
      return false;
  }
  ZeroTrailing(dst, n);
  if (t != 2 && t != 3 && dst != st->ldbase.data()) {
    memcpy(st->ldbase.data(), dst, words * 8ull);
    st->ldbase_vidx = vidx;
  }
  return true;
}

void PgenReader::SetView(std::vector<uint32_t> variant_map, std::vector<uint64_t> sample_keep, uint32_t kept_sample_ct) {
  vmap_ = std::move(variant_map);
  sample_keep_ = std::move(sample_keep);
  view_sample_ct_ = sample_keep_.empty() ? raw_sample_ct_ : kept_sample_ct;
}

const uint64_t* PgenReader::RawInclude(const uint64_t* view_include, uint32_t* sample_ct, std::vector<uint64_t>* scratch) const {
  if (sample_keep_.empty()) return view_include;
  if (!view_include) {
    *sample_ct = view_sample_ct_;
    return sample_keep_.data();
  }
  // bit k of view_include belongs to the k-th set bit of sample_keep_: deposit 64 raw samples at a time
  scratch->assign(sample_keep_.size(), 0);
  uint64_t pos = 0;  // view bits consumed
  for (size_t w = 0; w < sample_keep_.size(); ++w) {
    const uint64_t keep = sample_keep_[w];
    const uint32_t cnt = static_cast<uint32_t>(__builtin_popcountll(keep));
    if (!cnt) continue;
    const uint32_t sh = static_cast<uint32_t>(pos & 63);
    uint64_t bits = view_include[pos >> 6] >> sh;
    if (sh && sh + cnt > 64) bits |= view_include[(pos >> 6) + 1] << (64 - sh);
    (*scratch)[w] = _pdep_u64(bits, keep);
    pos += cnt;
  }
  return scratch->data();
}

bool PgenReader::Get(uint32_t vidx, uint64_t* genovec, std::string* err) {
  if (sample_keep_.empty()) return DecodeRecord(&state_, RawV(vidx), genovec, err);
  return GetSubsetWith(&state_, RawV(vidx), sample_keep_.data(), view_sample_ct_, genovec, err);
}

bool PgenReader::GetSubset(uint32_t vidx, const uint64_t* sample_include, uint32_t sample_ct, uint64_t* genovec, std::string* err) {
  std::vector<uint64_t> scratch;
  const uint64_t* raw_include = RawInclude(sample_include, &sample_ct, &scratch);
  return GetSubsetWith(&state_, RawV(vidx), raw_include, sample_ct, genovec, err);
}

bool PgenReader::GetBlock(const uint32_t* view_vidx, uint32_t count, const uint64_t* view_include, uint32_t sample_ct, uint64_t* dst, uint64_t stride_words, uint32_t thread_ct, std::string* err) {
  if (!count) return true;
  std::vector<uint64_t> include_scratch;
  std::vector<uint32_t> raw_vidx;
  const uint64_t* sample_include = RawInclude(view_include, &sample_ct, &include_scratch);
  const uint32_t* vidx = view_vidx;
  if (!vmap_.empty()) {
    raw_vidx.resize(count);
    for (uint32_t k = 0; k < count; ++k) raw_vidx[k] = vmap_[view_vidx[k]];
    vidx = raw_vidx.data();
  }
  thread_ct = std::max(1u, std::min(thread_ct, (count + 255) / 256));  // at least 256 variants per worker
  if (thread_ct == 1) {
    // own decode state (not state_): GetBlock may be called from several host threads at once (one per device in the
    // multi-GPU LD prune), and a short block must not share the LD-base cache of Get / GetSubset
    DecodeState st;
    st.ldbase.assign(WordsFor(raw_sample_ct_), 0);
    st.scratch.assign(WordsFor(raw_sample_ct_), 0);
    for (uint32_t k = 0; k < count; ++k) {
      uint64_t* row = dst + static_cast<uint64_t>(k) * stride_words;
      if (!(sample_include ? GetSubsetWith(&st, vidx[k], sample_include, sample_ct, row, err) : DecodeRecord(&st, vidx[k], row, err))) return false;
    }
    return true;
  }
  std::vector<std::string> errs(thread_ct);
  std::vector<uint8_t> failed(thread_ct, 0);
  std::vector<std::thread> workers;
  const uint32_t words = WordsFor(raw_sample_ct_);
  for (uint32_t t = 0; t < thread_ct; ++t) {
    const uint32_t k0 = static_cast<uint32_t>(static_cast<uint64_t>(count) * t / thread_ct), k1 = static_cast<uint32_t>(static_cast<uint64_t>(count) * (t + 1) / thread_ct);
    workers.emplace_back([=, &errs, &failed]() {
      DecodeState st;
      st.ldbase.assign(words, 0);
      st.scratch.assign(words, 0);
      for (uint32_t k = k0; k < k1; ++k) {
        uint64_t* row = dst + static_cast<uint64_t>(k) * stride_words;
        if (!(sample_include ? GetSubsetWith(&st, vidx[k], sample_include, sample_ct, row, &errs[t]) : DecodeRecord(&st, vidx[k], row, &errs[t]))) {
          failed[t] = 1;
          return;
        }
      }
    });
  }
  for (auto& w : workers) w.join();
  for (uint32_t t = 0; t < thread_ct; ++t) {
    if (failed[t]) {
      *err = errs[t];
      return false;
    }
  }
  return true;
}

bool PgenReader::GetSubsetWith(DecodeState* st, uint32_t vidx, const uint64_t* sample_include, uint32_t sample_ct, uint64_t* genovec, std::string* err) const {
  if (sample_ct == raw_sample_ct_) return DecodeRecord(st, vidx, genovec, err);
  if (!DecodeRecord(st, vidx, st->scratch.data(), err)) return false;
  // CopyNyparrNonemptySubset: gather the 2-bit entries of included samples
  const uint32_t out_words = WordsFor(sample_ct);
  for (uint32_t w = 0; w < out_words; ++w) genovec[w] = 0;
  uint32_t out_pos = 0;
  const uint32_t raw_words32 = (raw_sample_ct_ + 31) / 32;
  for (uint32_t w = 0; w < raw_words32; ++w) {
    // sample_include word w/2 holds 64 samples; take the 32 that belong to genovec word w
    const uint32_t inc = static_cast<uint32_t>(sample_include[w >> 1] >> (32 * (w & 1)));
    if (!inc) continue;
    const uint64_t mask2 = _pdep_u64(inc, 0x5555555555555555ull) * 3;
    const uint64_t packed = _pext_u64(st->scratch[w], mask2);
    const uint32_t cnt = static_cast<uint32_t>(__builtin_popcount(inc));
    const uint32_t sh = 2 * (out_pos & 31);
    genovec[out_pos >> 5] |= packed << sh;
    if (sh && sh + 2 * cnt > 64) genovec[(out_pos >> 5) + 1] |= packed >> (64 - sh);
    out_pos += cnt;
  }
  return true;
}

}  // namespace pl2host
