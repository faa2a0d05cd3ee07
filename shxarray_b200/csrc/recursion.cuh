// Device-side normalised associated Legendre recursion, column-wise (fixed order m, increasing degree n).
//
// Reproduces Legendre_nm<double>::set (Legendre_nm.hpp:106-131) BIT FOR BIT: every operation is an explicitly
// rounded __d*_rn intrinsic (never contracted into an FMA by nvcc) and the 1e-280 / sectorial scaling is applied
// in the same order.  The division p2 / wnm[n-1] is evaluated as
//     q = p2 * r;  q = fma(fma(-q, w, p2), r, q)        with r = RN(1/w) from the host
// (Markstein's correction step), which returns the correctly rounded quotient -- verified against IEEE division on
// all 18.6 M recursion steps of nmax=2160 at 8 latitudes, and on the device by tests/test_parity_gpu.py::
// test_device_recursion_bit_exact.  FP64 denormals are handled in hardware on sm_100a exactly as on the host.
#pragma once
#include <cuda_runtime.h>

namespace shb {

// one step n >= m+2 in the reference's 1e-280-scaled domain: returns pn
__device__ __forceinline__ double leg_step(double c, double p1, double p2, double w_n, double w_nm1, double r_nm1) {
    double q = __dmul_rn(p2, r_nm1);
    q = __fma_rn(__fma_rn(-q, w_nm1, p2), r_nm1, q);                       // == p2 / w_nm1   (Legendre_nm.hpp:118)
    return __dmul_rn(w_n, __dsub_rn(__dmul_rn(c, p1), q));                  // Legendre_nm.hpp:117
}

struct LegExact {
    double c, p1, p2, sect;
    int n, m;
    // returns P_nm for n = m, m+1, ... on successive calls.  w/r = tables + tri(m,m,nmax) (so w[k] = wnm(m+k, m))
    __device__ __forceinline__ void init(double costheta, double sect_m, int order) {
        c = costheta; sect = sect_m; m = order; n = order; p1 = 0.0; p2 = 1e-280;
    }
    __device__ __forceinline__ double next(const double* __restrict__ w, const double* __restrict__ r, double pmm) {
        int k = n - m;
        double val;
        if (k == 0) {
            val = pmm;                                         // Legendre_nm.hpp:103 / :129-130
        } else if (k == 1) {
            p1 = __dmul_rn(__dmul_rn(w[1], c), p2);            // :111
            val = __dmul_rn(p1, sect);                         // :112
        } else {
            double pn = leg_step(c, p1, p2, w[k], w[k - 1], r[k - 1]);
            val = __dmul_rn(pn, sect);                         // :120
            p2 = p1; p1 = pn;
        }
        ++n;
        return val;
    }
    // same step with the table entries passed by value: wk = wnm(n, m), wkm1 = wnm(n-1, m), rkm1 = RN(1 / wnm(n-1, m))
    __device__ __forceinline__ double next_v(double wk, double wkm1, double rkm1, double pmm) {
        const int k = n - m;
        double val;
        if (k == 0) {
            val = pmm;
        } else if (k == 1) {
            p1 = __dmul_rn(__dmul_rn(wk, c), p2);
            val = __dmul_rn(p1, sect);
        } else {
            const double pn = leg_step(c, p1, p2, wk, wkm1, rkm1);
            val = __dmul_rn(pn, sect);
            p2 = p1; p1 = pn;
        }
        ++n;
        return val;
    }
};

}  // namespace shb
