// Kernel-family dispatch for the Legendre stage.
#include "internal.h"

namespace shb {
void launch_legendre_syn_dfma(const Plan& p, cudaStream_t st);
bool launch_legendre_ana_dfma(const Plan& p, cudaStream_t st);
bool launch_legendre_syn_tab(const Plan& p, cudaStream_t st);
bool launch_legendre_ana_tab(const Plan& p, cudaStream_t st);
bool launch_legendre_syn_dmma(const Plan& p, cudaStream_t st);
bool launch_legendre_ana_dmma(const Plan& p, cudaStream_t st);

int device_sm_count() {
    static int sms[64] = {};
    int d = 0;
    cudaGetDevice(&d);
    if (d < 0 || d >= 64) d = 0;
    if (!sms[d]) cudaDeviceGetAttribute(&sms[d], cudaDevAttrMultiProcessorCount, d);
    return sms[d];
}

void launch_legendre_syn(const Plan& p, cudaStream_t st) {
    if (p.dmma && launch_legendre_syn_tab(p, st)) return;
    if (p.dmma && launch_legendre_syn_dmma(p, st)) return;
    launch_legendre_syn_dfma(p, st);
}
// false: the scratch array of a very tall grid (recursion state of > 8192 rows per CTA) could not be allocated
bool launch_legendre_ana(const Plan& p, cudaStream_t st) {
    if (p.dmma && launch_legendre_ana_tab(p, st)) return true;
    if (p.dmma && launch_legendre_ana_dmma(p, st)) return true;
    return launch_legendre_ana_dfma(p, st);
}
}  // namespace shb
