#include "pca.h"

#include <cstdio>

namespace pl2host {

int RunPca(const std::string&, uint32_t, bool, bool, uint64_t, uint32_t, Dataset*, Pl2GpuCtx*, Pl2GrmJob*) {
  fprintf(stderr, "Error: --pca is not available in this build yet.\n");
  return 63;
}

}  // namespace pl2host
