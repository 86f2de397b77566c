#include "ped_import.h"

#include <cstdlib>
#include <cstring>
#include <vector>

#include "text_util.h"

namespace pl2host {

namespace {
bool IsMissingCode(const char* s, size_t len) { return len == 1 && (s[0] == '.' || s[0] == '0'); }
}  // namespace

int PedmapToBed(const std::string& ped_path, const std::string& map_path, const std::string& out_prefix, uint32_t* sample_ct_out, uint32_t* variant_ct_out, std::string* err) {
  // ---- .map
  std::vector<std::string> map_lines;
  if (!ReadLines(map_path, &map_lines, err)) return 3;
  struct MapRow {
    std::string chr, id, cm;
    long bp;
  };
  std::vector<MapRow> rows;        // kept variants
  std::vector<uint8_t> raw_kept;   // per raw .map line: kept?
  uint32_t map_cols = 0;
  for (size_t li = 0; li < map_lines.size(); ++li) {
    const std::string& l = map_lines[li];
    if (l.empty() || l[0] == '#') continue;
    const std::vector<std::string> t = SplitWs(l);
    if (t.empty()) continue;
    if (!map_cols) {
      map_cols = static_cast<uint32_t>(t.size());
      if (map_cols > 4) {
        *err = map_path + " is not a .map file (too many columns).";
        return 6;
      }
      if (map_cols < 3) {
        *err = "Line " + std::to_string(li + 1) + " of " + map_path + " has fewer tokens than expected.";
        return 6;
      }
    }
    if (t.size() < map_cols) {
      *err = "Line " + std::to_string(li + 1) + " of " + map_path + " has fewer tokens than expected.";
      return 6;
    }
    char* endp = nullptr;
    const std::string& bp_tok = t[map_cols - 1];
    const long bp = strtol(bp_tok.c_str(), &endp, 10);
    if (endp == bp_tok.c_str() || *endp || bp > 2147483646L || bp < -2147483646L) {
      *err = "Invalid bp coordinate on line " + std::to_string(li + 1) + " of " + map_path + ".";
      return 6;
    }
    raw_kept.push_back(bp >= 0);
    if (bp >= 0) rows.push_back({t[0], t[1], map_cols == 4 ? t[2] : std::string("0"), bp});
  }
  const uint32_t raw_m = static_cast<uint32_t>(raw_kept.size()), m = static_cast<uint32_t>(rows.size());
  if (!m) {
    *err = "No variants in " + map_path + ".";
    return 13;
  }
  // ---- .ped
  std::vector<std::string> ped_lines;
  if (!ReadLines(ped_path, &ped_lines, err)) return 3;
  std::vector<std::string> ref(m), alt(m);      // provisional alleles ("" = not seen yet)
  std::vector<uint32_t> alt_plus_missing(m, 0);
  std::vector<std::vector<uint8_t>> codes;      // [sample][variant], .bed codes: 3 hom REF, 2 het, 0 hom ALT, 1 missing
  std::vector<std::string> fam_rows;
  int compound = -1;
  for (size_t li = 0; li < ped_lines.size(); ++li) {
    const std::string& l = ped_lines[li];
    if (l.empty() || l[0] == '#') continue;
    const std::vector<std::string> t = SplitWs(l);
    if (t.empty()) continue;
    if (compound < 0) {
      if (t.size() == 6 + 2ull * raw_m) compound = 0;
      else if (t.size() == 6ull + raw_m) compound = 1;
      else {
        *err = "Unexpected number of columns in .ped file (" + std::to_string(6ull + raw_m) + " or " + std::to_string(6 + 2ull * raw_m) + " expected).";
        return 7;
      }
    }
    if (t.size() < 6 + (compound ? 1ull : 2ull) * raw_m) {
      *err = "Line " + std::to_string(li + 1) + " of .ped file has fewer tokens than expected.";
      return 6;
    }
    fam_rows.push_back(t[0] + "\t" + t[1] + "\t" + t[2] + "\t" + t[3] + "\t" + t[4] + "\t" + t[5]);
    codes.emplace_back(m);
    std::vector<uint8_t>& row = codes.back();
    uint32_t v = 0;
    for (uint32_t rv = 0; rv < raw_m; ++rv) {
      const char *a1, *a2;
      size_t l1, l2;
      if (compound) {
        const std::string& g = t[6 + rv];
        if (g.size() != 2) {
          *err = "--pedmap: .map file and number of tokens in first line of .ped file imply that the latter is in the compound-genotypes format, but line " + std::to_string(li + 1) + " has a genotype that isn't length-2.";
          return 7;
        }
        a1 = g.data();
        a2 = g.data() + 1;
        l1 = l2 = 1;
      } else {
        a1 = t[6 + 2 * rv].data();
        l1 = t[6 + 2 * rv].size();
        a2 = t[7 + 2 * rv].data();
        l2 = t[7 + 2 * rv].size();
      }
      if (!raw_kept[rv]) continue;
      const bool het = l1 != l2 || memcmp(a1, a2, l1) != 0;
      auto same = [](const std::string& s, const char* p, size_t len) { return !s.empty() && s.size() == len && !memcmp(s.data(), p, len); };
      auto half_missing = [&]() {
        *err = "Half-missing genotype on line " + std::to_string(li + 1) + " of " + ped_path + ".";
        return 6;
      };
      auto multiallelic = [&]() {
        *err = "Multiallelic variant in " + ped_path + ". This violates the .ped specification; please reformat the file as e.g. VCF.";
        return 6;
      };
      uint8_t g;
      int as_ref = -1;  // 1: first allele is REF, 0: first allele is ALT
      if (same(ref[v], a1, l1)) as_ref = 1;
      else if (same(alt[v], a1, l1)) as_ref = 0;
      else if (IsMissingCode(a1, l1)) {
        if (het) return half_missing();
        as_ref = 2;
      } else if (ref[v].empty()) {
        ref[v].assign(a1, l1);
        as_ref = 1;
      } else if (alt[v].empty()) {
        alt[v].assign(a1, l1);
        as_ref = 0;
      } else {
        return multiallelic();
      }
      if (as_ref == 2) {
        g = 1;
      } else if (as_ref == 1) {
        g = het ? 2 : 3;
        if (het && !same(alt[v], a2, l2)) {
          if (!alt[v].empty()) return IsMissingCode(a2, l2) ? half_missing() : multiallelic();
          if (IsMissingCode(a2, l2)) return half_missing();
          alt[v].assign(a2, l2);
        }
      } else {
        g = het ? 2 : 0;
        if (het && !same(ref[v], a2, l2)) return IsMissingCode(a2, l2) ? half_missing() : multiallelic();
      }
      alt_plus_missing[v] += (4u - g) >> 1;
      row[v] = g;
      ++v;
    }
  }
  const uint32_t n = static_cast<uint32_t>(codes.size());
  if (!n) {
    *err = "No samples in " + ped_path + ".";
    return 13;
  }
  // ---- REF = major allele
  std::vector<uint8_t> flip(m, 0);
  for (uint32_t v = 0; v < m; ++v) {
    if (alt_plus_missing[v] > n) {
      flip[v] = 1;
      std::swap(ref[v], alt[v]);
    }
  }
  // ---- temporary fileset
  {
    OutFile f;
    if (!f.Open(out_prefix + ".fam")) {
      *err = "Failed to open " + out_prefix + ".fam for writing.";
      return 3;
    }
    for (const std::string& r : fam_rows) {
      f.Write(r.data(), r.size());
      f.Write("\n", 1);
    }
    if (!f.Close()) return 5;
  }
  {
    OutFile f;
    if (!f.Open(out_prefix + ".bim")) {
      *err = "Failed to open " + out_prefix + ".bim for writing.";
      return 3;
    }
    for (uint32_t v = 0; v < m; ++v) {
      const std::string line = rows[v].chr + "\t" + rows[v].id + "\t" + rows[v].cm + "\t" + std::to_string(rows[v].bp) + "\t" + (alt[v].empty() ? "." : alt[v]) + "\t" + (ref[v].empty() ? "." : ref[v]) + "\n";
      f.Write(line.data(), line.size());
    }
    if (!f.Close()) return 5;
  }
  {
    OutFile f;
    if (!f.Open(out_prefix + ".bed")) {
      *err = "Failed to open " + out_prefix + ".bed for writing.";
      return 3;
    }
    const uint8_t magic[3] = {0x6c, 0x1b, 0x01};
    f.Write(magic, 3);
    std::vector<uint8_t> row((n + 3) / 4);
    for (uint32_t v = 0; v < m; ++v) {
      std::fill(row.begin(), row.end(), 0);
      for (uint32_t k = 0; k < n; ++k) {
        uint8_t g = codes[k][v];
        if (flip[v] && (g == 0 || g == 3)) g = 3 - g;
        row[k >> 2] |= static_cast<uint8_t>(g << (2 * (k & 3)));
      }
      f.Write(row.data(), row.size());
    }
    if (!f.Close()) return 5;
  }
  *sample_ct_out = n;
  *variant_ct_out = m;
  return 0;
}

}  // namespace pl2host
