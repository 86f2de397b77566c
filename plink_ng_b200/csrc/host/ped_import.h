// ped_import.h - `--ped` + `--map` / `--pedmap` input (PedmapToPgen, 2.0/plink2_import_legacy.cc:1926-2400; ScanMap
// :52): the legacy text fileset is converted once, up front, to a temporary binary fileset that every command then
// reads like any other input - the reference writes <out>-temporary.pgen/.pvar/.psam, this program
// <out>-temporary.bed/.bim/.fam.  BASELINE.json's configs[0] (1.9/toy.ped + toy.map) is such a fileset.
#pragma once
#include <cstdint>
#include <string>

namespace pl2host {

// Allele rules of the reference: per variant the first allele code met while scanning the .ped (samples in file
// order, first then second allele of each call) is the provisional REF, the second distinct code the provisional ALT;
// a third code, or one missing and one present allele in a call, is an error.  After the scan REF and ALT are swapped
// when (ALT allele count + missing calls) exceeds the sample count, so REF ends up the major allele (:2300-2312).
// .map: 3 or 4 columns (CHR ID [CM] BP); variants with a negative BP are dropped.  Both the regular (two allele tokens
// per call) and the compound-genotypes (one 2-character token) .ped layouts are accepted, told apart by the token count
// of the first line.  Returns 0, or a PglErr-valued code with *err set.
int PedmapToBed(const std::string& ped_path, const std::string& map_path, const std::string& out_prefix, uint32_t* sample_ct, uint32_t* variant_ct, std::string* err);

}  // namespace pl2host
