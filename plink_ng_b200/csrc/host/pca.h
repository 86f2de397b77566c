// pca.h - --pca driver (CalcPca, 2.0/plink2_matrix_calc.cc:5594).
#pragma once
#include <cstdint>
#include <string>

#include "../../../include/plink2_b200.h"
#include "dataset.h"

namespace pl2host {

// exact: top-k eigenpairs of the GRM held by `grm_job`; approx: randomized range finder.
// Returns a PglErr-style exit code.
int RunPca(const std::string& out_prefix, uint32_t pc_ct, bool approx, bool seed_given, uint64_t seed, uint32_t threads, Dataset* ds, Pl2GpuCtx* ctx, Pl2GrmJob* grm_job);

}  // namespace pl2host
