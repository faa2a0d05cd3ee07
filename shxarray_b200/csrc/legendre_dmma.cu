// FP64 tensor-core (DMMA) Legendre-stage kernels -- placeholder until the tuned kernels land.
#include "internal.h"
namespace shb {
bool launch_legendre_syn_dmma(const Plan&, cudaStream_t) { return false; }
bool launch_legendre_ana_dmma(const Plan&, cudaStream_t) { return false; }
}  // namespace shb
