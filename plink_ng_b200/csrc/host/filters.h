// filters.h - the sample / variant filters in front of the pairwise-genotype commands, and the .bed fileset writer
// that pins them.  In the reference every command receives sample_include / variant_include bitsets prepared by
// Plink2Core (2.0/plink2.cc:1423-1665: --chr / --not-chr / --autosome[-xy] at load, --extract / --exclude by ID,
// then --keep-fam, --keep, --remove-fam, --remove); here the same filters compact the Dataset tables and install the
// matching view in the PgenReader, so every command driver simply sees a smaller dataset.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "dataset.h"

namespace pl2host {

struct FilterSpec {
  std::vector<std::string> keep, remove, keep_fam, remove_fam;  // ID files (LoadSampleIds, plink2_common.cc:1707)
  std::vector<std::string> extract, exclude;                    // variant-ID token files (TokenExtractExclude)
  std::vector<uint8_t> chr_mask, not_chr_mask;                  // 27 flags each when the flag was given
  bool autosome = false, autosome_xy = false;
  bool any() const { return !(keep.empty() && remove.empty() && keep_fam.empty() && remove_fam.empty() && extract.empty() && exclude.empty() && chr_mask.empty() && not_chr_mask.empty()) || autosome || autosome_xy; }
};

// "1-4,22,X" style arguments (ParseChrRanges, plink2_common.cc:3695) -> 27 flags.  False + *err on a bad token.
bool ParseChrList(const std::vector<std::string>& args, const char* flag, std::vector<uint8_t>* mask, std::string* err);

// Applies the filters in the reference's order.  Log lines ("--keep: 61 samples remaining.") are appended to *log.
// Returns 0, or a PglErr-valued code with *err set (3 open failure, 6 malformed file, 7 nothing left).
int ApplyFilters(const FilterSpec& spec, Dataset* ds, std::vector<std::string>* log, std::string* err);

// keep[k] != 0: sample / variant k (current numbering) stays.  Compacts the tables and updates the reader view.
void KeepSamples(Dataset* ds, const std::vector<uint8_t>& keep);
void KeepVariants(Dataset* ds, const std::vector<uint8_t>& keep);

// --make-bed (MakePlink2NoVsort -> .bed/.bim/.fam writers, 2.0/plink2_data.cc): the current view as a PLINK 1 binary
// fileset.  Host-only.  Returns 0 or a code with *err set.  With `sample_include` (bitset over the view's samples,
// include_ct set bits) only the .bed of that subset is written: the test hook for the subset-of-a-view decode that
// the founder-only commands (LD prune, allele frequencies) use.
int WriteBedFileset(Dataset* ds, const std::string& out_prefix, uint32_t thread_ct, std::string* err, const uint64_t* sample_include = nullptr, uint32_t include_ct = 0);

}  // namespace pl2host
