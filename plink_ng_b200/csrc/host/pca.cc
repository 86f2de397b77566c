#include "pca.h"

#include <cstdio>
#include <cstring>
#include <vector>

#include "text_util.h"

namespace pl2host {


// .eigenvec / .eigenval writers (plink2_matrix_calc.cc:6237-6290): "#[FID\t]IID[\tSID]\tPC1..PCk",
// one line per sample; eigenvalues one per line, both through dtoa_g.
bool WriteEigen(const std::string& out_prefix, const SampleInfo& S, uint32_t pc_ct, const double* eigvals, const double* eigvecs) {
  const uint32_t n = S.size();
  OutFile fv, fe;
  if (!fv.Open(out_prefix + ".eigenvec") || !fe.Open(out_prefix + ".eigenval")) return false;
  std::string h = "#";
  if (S.fid_present) h += "FID\t";
  h += "IID";
  if (S.sid_present) h += "\tSID";
  for (uint32_t k = 0; k < pc_ct; ++k) h += "\tPC" + std::to_string(k + 1);
  h += "\n";
  fv.Puts(h.c_str());
  for (uint32_t s = 0; s < n; ++s) {
    std::string id;
    if (S.fid_present) id += S.fid[s] + "\t";
    id += S.iid[s];
    if (S.sid_present) id += "\t" + S.sid[s];
    char* w = fv.Reserve(id.size() + 16 * pc_ct + 16);
    memcpy(w, id.data(), id.size());
    w += id.size();
    for (uint32_t k = 0; k < pc_ct; ++k) {
      *w++ = '\t';
      w = dtoa_g(eigvecs[static_cast<uint64_t>(k) * n + s], w);
    }
    *w++ = '\n';
    fv.Advance(w);
  }
  for (uint32_t k = 0; k < pc_ct; ++k) {
    char* w = fe.Reserve(32);
    w = dtoa_g(eigvals[k], w);
    *w++ = '\n';
    fe.Advance(w);
  }
  return fv.Close() && fe.Close();
}


}  // namespace pl2host
