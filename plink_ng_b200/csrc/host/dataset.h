// dataset.h - minimal .psam/.fam + .pvar/.bim loaders and the genotype-block streamer used by the
// command drivers.  Only what the pairwise-genotype commands read is kept: sample IDs (FID/IID/SID),
// founder status, variant chromosome / bp / ID (2.0/plink2_psam.cc LoadPsam, 2.0/plink2_pvar.cc
// LoadPvar; file rules pgen_spec/pgen_spec.tex:695-833).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "pgen_reader.h"

namespace pl2host {

struct SampleInfo {
  std::vector<std::string> fid, iid, sid;
  std::vector<std::string> pat, mat;  // parental IDs as written ("0" when the column is absent)
  std::vector<uint8_t> is_founder;
  std::vector<uint8_t> sex;  // 0 unknown, 1 male, 2 female (.fam column 5 / .psam SEX)
  // phenotype columns as read (.fam column 6 = PHENO1; .psam: every column that is not an ID / parent / SEX column);
  // typed when written (binary / quantitative / categorical, LoadPsam plink2_psam.cc:58)
  std::vector<std::string> pheno_names;
  std::vector<std::vector<std::string>> pheno_tokens;  // [phenotype][sample]
  std::vector<std::string> fam_pheno;  // .fam column 6 as --make-bed writes it (typed over all loaded samples); empty: all -9
  bool fid_present = false;  // kfSampleIdFidPresent (plink2_psam.cc:104-130, :279, :823)
  bool sid_present = false;
  uint32_t size() const { return static_cast<uint32_t>(iid.size()); }
};

struct VariantInfo {
  std::vector<uint32_t> chr_code;  // 0 = unplaced, 1..22 autosomes, 23 X, 24 Y, 25 XY, 26 MT
  std::vector<uint32_t> bp;
  std::vector<std::string> id;
  std::vector<std::string> chr_name, ref, alt;  // as written in the file (.bim: REF = column 6, ALT = column 5)
  std::vector<uint8_t> zero_allele;             // bit 0 / 1: REF / ALT was written as the '0' missing code (stored as '.')
  std::vector<std::string> cm;                  // centimorgan token (.bim column 3 / .pvar CM column); empty: no such column
  bool provisional_ref = false;                 // .bim input: REF alleles are provisional (PROVISIONAL_REF? = Y)
  uint32_t size() const { return static_cast<uint32_t>(id.size()); }
};

// chromosome token -> code (optional chr prefix, case-insensitive X / Y / XY / PAR1 / PAR2 / MT / M, 0..26)
bool ParseChr(const std::string& tok, uint32_t* code);
// Chromosome as the reference prints it under its default output encoding (kfChrOutputMT; chrtoa / ChrNameStd,
// 2.0/plink2_common.cc:2150-2227): bare number for autosomes, X / Y / XY / MT, PAR1 / PAR2 kept.
std::string ChrNameOut(uint32_t code, const std::string& as_read);
// --output-chr <26 | M | MT | chr26 | chrM | chrMT>: the mitochondrial spelling names the scheme (numeric 23-26 vs X / Y /
// XY / M[T], optional chr prefix); extra contigs are never touched.  Returns false for an unknown scheme.
bool SetOutputChrStyle(const std::string& mt_code);
bool LoadSamples(const std::string& path, SampleInfo* out, std::string* err);
// allow_extra_chr (--allow-extra-chr): a name outside the human set becomes its own diploid, autosome-like contig with a
// code >= 27 (one per distinct name, in order of appearance), printed as written - the reference's treatment of
// unrecognised contigs (not haploid, not in the --autosome set, kept by KING / GRM, its own LD-prune unit).
bool LoadVariants(const std::string& path, VariantInfo* out, std::string* err, bool allow_extra_chr = false);

inline bool IsAutosome(uint32_t chr_code) { return chr_code >= 1 && chr_code <= 22; }
// CountNonAutosomalVariants(..., count_x=1, count_mt=1) semantics used by CalcKing/CalcGrm
// (plink2_matrix_calc.cc:1704, :4654): X, Y, XY(PAR is kept by the reference; treated as
// autosomal-like), MT excluded; unplaced (0) kept.
inline bool KeptForRelationship(uint32_t chr_code) { return chr_code != 23 && chr_code != 24 && chr_code != 26; }

struct Dataset {
  SampleInfo samples;
  VariantInfo variants;
  PgenReader reader;
  std::vector<double> read_ref_freq;  // --read-freq: loaded REF frequency per variant, NaN = not loaded (empty: flag absent)
  // raw index of every kept sample / variant once a filter ran (empty: nothing filtered); `samples`, `variants` and
  // the reader's view are always compacted together (filters.cc)
  std::vector<uint32_t> sample_raw, variant_raw;
};

}  // namespace pl2host
