// pca.h - .eigenvec / .eigenval writers of --pca (CalcPca, 2.0/plink2_matrix_calc.cc:6237-6290).
#pragma once
#include <cstdint>
#include <string>

#include "dataset.h"

namespace pl2host {

// eigvecs: [pc][sample].  Returns false on write failure.
bool WriteEigen(const std::string& out_prefix, const SampleInfo& S, uint32_t pc_ct, const double* eigvals, const double* eigvecs);

}  // namespace pl2host
