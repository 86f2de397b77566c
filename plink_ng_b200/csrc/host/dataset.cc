#include "dataset.h"

#include <cstdlib>
#include <cstring>

#include "text_util.h"

namespace pl2host {

// human chromosome codes (the reference's default chr-set): 1-22, X=23, Y=24, XY=25, MT=26, 0 unplaced
bool ParseChr(const std::string& tok, uint32_t* code) {
  const char* s = tok.c_str();
  if ((s[0] == 'c' || s[0] == 'C') && (s[1] == 'h' || s[1] == 'H') && (s[2] == 'r' || s[2] == 'R')) s += 3;
  if (!*s) return false;
  char* endp = nullptr;
  const long v = strtol(s, &endp, 10);
  if (endp != s && !*endp) {
    if (v < 0 || v > 26) return false;
    *code = static_cast<uint32_t>(v);
    return true;
  }
  std::string u(s);
  for (auto& c : u) c = static_cast<char>(toupper(c));
  if (u == "X") *code = 23;
  else if (u == "Y") *code = 24;
  else if (u == "XY" || u == "PAR1" || u == "PAR2") *code = 25;
  else if (u == "MT" || u == "M") *code = 26;
  else return false;
  return true;
}

namespace {
bool g_out_chr_prefix = false;
int g_out_chr_mode = 2;  // 0 numeric (23..26), 1 "M", 2 "MT"
}  // namespace

bool SetOutputChrStyle(const std::string& mt_code) {
  std::string t = mt_code;
  g_out_chr_prefix = t.size() > 3 && t.compare(0, 3, "chr") == 0;
  if (g_out_chr_prefix) t = t.substr(3);
  if (t == "26") g_out_chr_mode = 0;
  else if (t == "M") g_out_chr_mode = 1;
  else if (t == "MT") g_out_chr_mode = 2;
  else return false;
  return true;
}

std::string ChrNameOut(uint32_t code, const std::string& as_read) {
  if (code > 26) return as_read;  // --allow-extra-chr contig: as written
  const std::string prefix = g_out_chr_prefix ? "chr" : "";
  if (code <= 22 || g_out_chr_mode == 0) {
    if (code == 25 && g_out_chr_mode == 0) {
      std::string u = as_read;
      if (u.size() > 3 && (u[0] == 'c' || u[0] == 'C') && (u[1] == 'h' || u[1] == 'H') && (u[2] == 'r' || u[2] == 'R')) u = u.substr(3);
      for (auto& ch : u) ch = static_cast<char>(toupper(ch));
      if (u == "PAR1" || u == "PAR2") return prefix + u;
    }
    return prefix + std::to_string(code);
  }
  if (code == 23) return prefix + "X";
  if (code == 24) return prefix + "Y";
  if (code == 26) return prefix + (g_out_chr_mode == 1 ? "M" : "MT");
  std::string u = as_read;
  if (u.size() > 3 && (u[0] == 'c' || u[0] == 'C') && (u[1] == 'h' || u[1] == 'H') && (u[2] == 'r' || u[2] == 'R')) u = u.substr(3);
  for (auto& ch : u) ch = static_cast<char>(toupper(ch));
  if (u == "PAR1" || u == "PAR2") return prefix + u;
  return prefix + "XY";
}

bool LoadSamples(const std::string& path, SampleInfo* out, std::string* err) {
  std::vector<std::string> lines;
  if (!ReadLines(path, &lines, err)) return false;
  size_t li = 0;
  // .psam: optional '##' comment lines then a '#FID ...' / '#IID ...' header; .fam: no header, 6 columns
  while (li < lines.size() && lines[li].size() >= 2 && lines[li][0] == '#' && lines[li][1] == '#') ++li;
  int col_fid = -1, col_iid = -1, col_sid = -1, col_pat = -1, col_mat = -1, col_sex = -1;
  std::vector<int> pheno_cols;
  if (li < lines.size() && !lines[li].empty() && lines[li][0] == '#') {
    std::vector<std::string> hdr = SplitWs(lines[li].substr(1));
    for (size_t c = 0; c < hdr.size(); ++c) {
      if (hdr[c] == "FID") col_fid = static_cast<int>(c);
      else if (hdr[c] == "IID") col_iid = static_cast<int>(c);
      else if (hdr[c] == "SID") col_sid = static_cast<int>(c);
      else if (hdr[c] == "PAT") col_pat = static_cast<int>(c);
      else if (hdr[c] == "MAT") col_mat = static_cast<int>(c);
      else if (hdr[c] == "SEX") col_sex = static_cast<int>(c);
      else {
        pheno_cols.push_back(static_cast<int>(c));
        out->pheno_names.push_back(hdr[c]);
      }
    }
    if (col_iid < 0 || (col_fid > 0)) {
      *err = "Invalid .psam header line in " + path + " (#FID or #IID must come first).";
      return false;
    }
    ++li;
  } else {
    col_fid = 0;
    col_iid = 1;
    col_pat = 2;
    col_mat = 3;
    col_sex = 4;
    pheno_cols.push_back(5);
    out->pheno_names.push_back("PHENO1");
  }
  out->pheno_tokens.assign(pheno_cols.size(), {});
  out->fid_present = col_fid >= 0;
  out->sid_present = col_sid >= 0;
  for (; li < lines.size(); ++li) {
    if (lines[li].empty()) continue;
    std::vector<std::string> t = SplitWs(lines[li]);
    if (t.empty()) continue;
    const int need = std::max(std::max(std::max(col_iid, col_sid), std::max(col_pat, col_mat)), col_sex);
    if (static_cast<int>(t.size()) <= need) {
      *err = "Line " + std::to_string(li + 1) + " of " + path + " has fewer tokens than expected.";
      return false;
    }
    out->fid.push_back(col_fid >= 0 ? t[col_fid] : "0");
    out->iid.push_back(t[col_iid]);
    out->sid.push_back(col_sid >= 0 ? t[col_sid] : "0");
    out->pat.push_back(col_pat >= 0 ? t[col_pat] : "0");
    out->mat.push_back(col_mat >= 0 ? t[col_mat] : "0");
    const bool founder = (col_pat < 0 || t[col_pat] == "0") && (col_mat < 0 || t[col_mat] == "0");
    out->is_founder.push_back(founder ? 1 : 0);
    // SEX: '1'/'M'/'m' male, '2'/'F'/'f' female, anything else unknown (plink2_psam.cc:609-623)
    uint8_t sex = 0;
    if (col_sex >= 0 && t[col_sex].size() == 1) {
      const char ch = t[col_sex][0];
      if (ch == '1' || ch == 'M' || ch == 'm') sex = 1;
      else if (ch == '2' || ch == 'F' || ch == 'f') sex = 2;
    }
    out->sex.push_back(sex);
    for (size_t pc = 0; pc < pheno_cols.size(); ++pc) out->pheno_tokens[pc].push_back(pheno_cols[pc] < static_cast<int>(t.size()) ? t[pheno_cols[pc]] : "NA");
  }
  if (out->iid.empty()) {
    *err = "No samples in " + path + ".";
    return false;
  }
  return true;
}

bool LoadVariants(const std::string& path, VariantInfo* out, std::string* err, bool allow_extra_chr) {
  std::vector<std::string> extra_names;
  std::vector<uint8_t> seen_chr;
  std::vector<std::string> lines;
  if (!ReadLines(path, &lines, err)) return false;
  size_t li = 0;
  while (li < lines.size() && lines[li].size() >= 2 && lines[li][0] == '#' && lines[li][1] == '#') ++li;
  int col_chr, col_pos, col_id, col_ref = -1, col_alt = -1, col_cm = -1;
  bool is_bim = false;
  if (li < lines.size() && !lines[li].empty() && lines[li][0] == '#') {
    // .pvar header: #CHROM POS ID REF ALT ...
    std::vector<std::string> hdr = SplitWs(lines[li].substr(1));
    col_chr = col_pos = col_id = -1;
    for (size_t c = 0; c < hdr.size(); ++c) {
      if (hdr[c] == "CHROM") col_chr = static_cast<int>(c);
      else if (hdr[c] == "POS") col_pos = static_cast<int>(c);
      else if (hdr[c] == "ID") col_id = static_cast<int>(c);
      else if (hdr[c] == "REF") col_ref = static_cast<int>(c);
      else if (hdr[c] == "ALT") col_alt = static_cast<int>(c);
      else if (hdr[c] == "CM") col_cm = static_cast<int>(c);
    }
    if (col_chr != 0 || col_pos < 0 || col_id < 0) {
      *err = "Invalid .pvar header line in " + path + ".";
      return false;
    }
    ++li;
  } else {
    // .bim: CHROM ID CM POS A1 A2   (5-column variant without CM also accepted)
    col_chr = 0;
    col_id = 1;
    col_pos = 3;
    col_alt = 4;
    col_ref = 5;
    col_cm = 2;
    is_bim = true;
  }
  for (; li < lines.size(); ++li) {
    if (lines[li].empty()) continue;
    std::vector<std::string> t = SplitWs(lines[li]);
    if (t.empty()) continue;
    int cpos = col_pos, cref = col_ref, calt = col_alt, ccm = col_cm;
    if (is_bim && t.size() == 5) {
      ccm = -1;
      cpos = 2;
      calt = 3;
      cref = 4;
    }
    if (static_cast<int>(t.size()) <= std::max(cpos, col_id)) {
      *err = "Line " + std::to_string(li + 1) + " of " + path + " has fewer tokens than expected.";
      return false;
    }
    uint32_t code;
    if (!ParseChr(t[col_chr], &code) && allow_extra_chr) {
      size_t k = 0;
      while (k < extra_names.size() && extra_names[k] != t[col_chr]) ++k;
      if (k == extra_names.size()) extra_names.push_back(t[col_chr]);
      code = 27 + static_cast<uint32_t>(k);
    } else if (!ParseChr(t[col_chr], &code)) {
      *err = "Invalid chromosome code '" + t[col_chr] + "' on line " + std::to_string(li + 1) + " of " + path + " (use --allow-extra-chr to keep contigs outside the human chromosome set).";
      return false;
    }
    // every chromosome must be one contiguous block (the reference: "has a split chromosome", LoadPvar) - the
    // per-chromosome drivers (LD prune, r^2 tables, frequency passes) rely on it
    if (!out->chr_code.empty() && code != out->chr_code.back()) {
      if (code < seen_chr.size() && seen_chr[code]) {
        *err = path + " has a split chromosome ('" + t[col_chr] + "' on line " + std::to_string(li + 1) + " after other chromosomes); sort the variants first (plink2 --make-pgen --sort-vars).";
        return false;
      }
    }
    if (code >= seen_chr.size()) seen_chr.resize(code + 1, 0);
    seen_chr[code] = 1;
    out->chr_code.push_back(code);
    out->bp.push_back(static_cast<uint32_t>(strtoul(t[cpos].c_str(), nullptr, 10)));
    out->id.push_back(t[col_id]);
    out->chr_name.push_back(t[col_chr]);
    // a lone '0' is the input missing-allele code and is stored as '.' (LoadPvar, plink2_pvar.cc: input_missing_geno_char)
    auto allele = [&](int col) { return col >= 0 && col < static_cast<int>(t.size()) && t[col] != "0" ? t[col] : std::string("."); };
    out->ref.push_back(allele(cref));
    out->alt.push_back(allele(calt));
    auto is_zero = [&](int col) { return col >= 0 && col < static_cast<int>(t.size()) && t[col] == "0"; };
    out->zero_allele.push_back(static_cast<uint8_t>((is_zero(cref) ? 1 : 0) | (is_zero(calt) ? 2 : 0)));
    if (col_cm >= 0) out->cm.push_back(ccm >= 0 && ccm < static_cast<int>(t.size()) ? t[ccm] : std::string("0"));
  }
  return true;
}

}  // namespace pl2host
