// pgen_reader.h - the slice of the pgenlib reader surface the pairwise-genotype commands use
// (2.0/include/pgenlib_read.h:442-747): open a .bed / .pgen, then PgrGet-style sequential or
// random access to variant-major packed 2-bit hard calls, optionally restricted to a sample
// subset.  Format per pgen_spec/pgen_spec.tex: storage modes 0x01 (.bed, :132-136), 0x02 (fixed
// width, :137-139) and 0x10 (variable width, :144-145) with main-track record types 0 (raw), 1
// (1-bit + difflist), 2/3 (LD-compressed), 4/6/7 (difflist against a constant) (:443-468).
// Auxiliary tracks (multiallelic, phase, dosage; record-type bits 3-7) are not hard calls of a
// biallelic variant: multiallelic records are rejected, phase/dosage tracks are ignored exactly as
// PgrGet ignores them.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace pl2host {

class PgenReader {
 public:
  // raw_sample_ct is required for .bed (the file does not store it) and cross-checked otherwise.
  // 0 = "unknown" for .pgen.  Returns false and sets *err on failure.
  bool Open(const std::string& path, uint32_t raw_sample_ct, uint32_t raw_variant_ct, std::string* err);
  void Close();
  ~PgenReader() { Close(); }

  uint32_t raw_sample_ct() const { return raw_sample_ct_; }
  uint32_t raw_variant_ct() const { return raw_variant_ct_; }
  // header bits 6-7 (pgenlib_read.cc:874): 1 = every REF allele trusted, 2 = all provisional (always for .bed,
  // :790), 3 = per-variant flags stored, 0 = not recorded
  uint32_t nonref_flags_storage() const { return mode_ == 0x01 ? 2u : nonref_mode_; }
  // uint64 words per variant for n samples: ceil(n / 32)
  static uint32_t WordsFor(uint32_t n) { return (n + 31) / 32; }

  // Filtered view (the role of the variant_include / sample_include bitsets every reference command receives after
  // the --extract / --chr / --keep / ... filters ran, 2.0/plink2.cc:1423-1665): once set, variant indices passed to
  // Get / GetSubset / GetBlock are positions in `variant_map` (empty = identity) and sample counts / sample_include
  // bitsets are over the kept samples (`sample_keep` = bitset over raw samples, empty = all).  Callers compact their
  // sample / variant tables the same way and never see raw indices again.
  void SetView(std::vector<uint32_t> variant_map, std::vector<uint64_t> sample_keep, uint32_t kept_sample_ct);
  uint32_t view_sample_ct() const { return view_sample_ct_; }

  // PgrGet without subsetting: all raw samples of variant `vidx` into genovec[WordsFor(raw_sample_ct)].
  // Codes 0 hom-REF, 1 het, 2 hom-ALT, 3 missing; trailing entries of the last word are zero.
  bool Get(uint32_t vidx, uint64_t* genovec, std::string* err);
  // PgrGet with a sample subset (sample_include bitset over raw samples, sample_ct set bits):
  // subsetted genovec[WordsFor(sample_ct)] (CopyNyparrNonemptySubset semantics).
  bool GetSubset(uint32_t vidx, const uint64_t* sample_include, uint32_t sample_ct, uint64_t* genovec, std::string* err);
  // Multi-threaded block read (the role of PgenMtLoadInit + the per-thread PgenReaders of the reference's block
  // readers, 2.0/plink2_common.cc:3926): variants vidx[0..count) -> dst[k * stride_words], decoded by
  // `thread_ct` workers over contiguous sub-ranges (each with its own LD-base cache).  sample_include = nullptr:
  // all raw samples.  Returns false (first error) if any record fails.
  bool GetBlock(const uint32_t* vidx, uint32_t count, const uint64_t* sample_include, uint32_t sample_ct, uint64_t* dst, uint64_t stride_words, uint32_t thread_ct, std::string* err);

 private:
  // per-decoder mutable state: the last non-LD-compressed genovec (LD base) and a subsetting scratch row
  struct DecodeState {
    std::vector<uint64_t> ldbase, scratch;
    uint32_t ldbase_vidx = 0xFFFFFFFFu;
  };
  // as_ld_base: only the main track of `vidx` is wanted, as the base of an LD-compressed neighbour - a multiallelic
  // record is acceptable then (its main track has the biallelic layout; the allele patches that follow are not read)
  bool DecodeRecord(DecodeState* st, uint32_t vidx, uint64_t* dst, std::string* err, bool as_ld_base = false) const;
  bool GetSubsetWith(DecodeState* st, uint32_t vidx, const uint64_t* sample_include, uint32_t sample_ct, uint64_t* genovec, std::string* err) const;
  bool ReadRecordBytes(uint32_t vidx, const uint8_t** rec, uint32_t* len, std::string* err) const;
  bool ParseDifflistAndApply(const uint8_t* p, const uint8_t* end, bool with_values, uint64_t* genovec, uint32_t fixed_value, std::string* err, const uint8_t** after) const;

  int fd_ = -1;
  const uint8_t* map_ = nullptr;
  uint64_t map_len_ = 0;
  uint8_t mode_ = 0;
  uint32_t nonref_mode_ = 0;
  uint32_t raw_sample_ct_ = 0;
  uint32_t raw_variant_ct_ = 0;
  uint64_t fixed_start_ = 0;      // modes 0x01/0x02: offset of record 0
  uint32_t fixed_bpv_ = 0;
  std::vector<uint8_t> vrtypes_;  // mode 0x10
  std::vector<uint64_t> rec_off_; // mode 0x10: [raw_variant_ct + 1]
  DecodeState state_;             // used by Get / GetSubset (single-threaded callers)
  // filtered view
  std::vector<uint32_t> vmap_;         // view variant index -> raw variant index (empty: identity)
  std::vector<uint64_t> sample_keep_;  // raw-sample bitset of the view (empty: all samples)
  uint32_t view_sample_ct_ = 0;
  uint32_t RawV(uint32_t v) const { return vmap_.empty() ? v : vmap_[v]; }
  // view-space include (nullptr = every kept sample) -> raw-space include, or nullptr when that is "all raw samples"
  const uint64_t* RawInclude(const uint64_t* view_include, uint32_t* sample_ct, std::vector<uint64_t>* scratch) const;
};

}  // namespace pl2host
