// Device-side normalised associated Legendre recursion, column-wise (fixed order m, increasing degree n).
//
// EXACT variant: reproduces Legendre_nm<double>::set (Legendre_nm.hpp:106-131) bit for bit: every
// operation is an explicitly rounded __d*_rn intrinsic (never contracted into FMA by nvcc), the
// division is a true IEEE division, and the 1e-280 / sectorial scaling is applied in the same order.
// FP64 denormals are handled in hardware on sm_100a exactly as on the host.
//
// FAST variant: pn = a_n*(c*p1) - b_n*p2 with b_n = wnm[n]/wnm[n-1] precomputed on the host
// (division-free, one FMA); values stay in the reference's 1e-280-scaled domain and the per-(m,row)
// sectorial multiplier is applied by the caller to the contracted result.  Differs from the reference
// by rounding only (<= 2.3e-13 field-norm).
#pragma once
#include <cuda_runtime.h>

namespace shb {

struct LegExact {
    double c, p1, p2, sect;
    int n, m;
    // returns P_nm for n = m, m+1, ... on successive calls.  w = wnm + tri(m,m,nmax) (so w[k] = wnm(m+k, m))
    __device__ __forceinline__ void init(double costheta, double sect_m, int order) {
        c = costheta; sect = sect_m; m = order; n = order; p1 = 0.0; p2 = 1e-280;
    }
    __device__ __forceinline__ double next(const double* __restrict__ w, double pmm) {
        int k = n - m;
        double val;
        if (k == 0) {
            val = pmm;                                         // Legendre_nm.hpp:103 / :129-130
        } else if (k == 1) {
            p1 = __dmul_rn(__dmul_rn(w[1], c), p2);            // :111
            val = __dmul_rn(p1, sect);                         // :112
        } else {
            double pn = __dmul_rn(w[k], __dsub_rn(__dmul_rn(c, p1), __ddiv_rn(p2, w[k - 1])));  // :117-118
            val = __dmul_rn(pn, sect);                         // :120
            p2 = p1; p1 = pn;
        }
        ++n;
        return val;
    }
};

}  // namespace shb
