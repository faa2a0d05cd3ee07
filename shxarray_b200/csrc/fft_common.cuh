// Register-level pieces shared by the longitude FFT kernels (lon_fft.cu: in-place mixed-radix passes for any 2,3,5-smooth
// length; lon_fft3.cu: the three-pass composite-radix kernels of the usual grid lengths).
#pragma once
#include "internal.h"

namespace shb {
namespace fft {

// barrier among the `count` threads that work on one transform (ids 1..15; id 0 is __syncthreads)
__device__ __forceinline__ void group_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
// multiply by s*i
__device__ __forceinline__ double2 mul_si(double2 a, double s) { return make_double2(-s * a.y, s * a.x); }

// length-R DFT in registers, X[k] = sum_n v[n] exp(s*2*pi*i*n*k/R); s = -1 forward, +1 inverse
template <int R>
__device__ __forceinline__ void dft(double2* v, double s);

template <>
__device__ __forceinline__ void dft<2>(double2* v, double) {
    double2 a = v[0], b = v[1];
    v[0] = cadd(a, b); v[1] = csub(a, b);
}
__device__ __forceinline__ void dft4(double2& b0, double2& b1, double2& b2, double2& b3, double s) {
    double2 p = cadd(b0, b2), q = csub(b0, b2), r = cadd(b1, b3), t = mul_si(csub(b1, b3), s);
    b0 = cadd(p, r); b2 = csub(p, r); b1 = cadd(q, t); b3 = csub(q, t);
}
template <>
__device__ __forceinline__ void dft<4>(double2* v, double s) { dft4(v[0], v[1], v[2], v[3], s); }
template <>
__device__ __forceinline__ void dft<8>(double2* v, double s) {
    const double h = 0.70710678118654752440;
    double2 a0 = cadd(v[0], v[4]), c0 = csub(v[0], v[4]);
    double2 a1 = cadd(v[1], v[5]), c1 = csub(v[1], v[5]);
    double2 a2 = cadd(v[2], v[6]), c2 = csub(v[2], v[6]);
    double2 a3 = cadd(v[3], v[7]), c3 = csub(v[3], v[7]);
    // c_n *= w8^n, w8 = exp(s*i*pi/4)
    c1 = make_double2(h * (c1.x - s * c1.y), h * (c1.y + s * c1.x));
    c2 = mul_si(c2, s);
    c3 = make_double2(h * (-c3.x - s * c3.y), h * (-c3.y + s * c3.x));
    dft4(a0, a1, a2, a3, s);
    dft4(c0, c1, c2, c3, s);
    v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
    v[1] = c0; v[3] = c1; v[5] = c2; v[7] = c3;
}
template <>
__device__ __forceinline__ void dft<3>(double2* v, double s) {
    const double c3 = -0.5, s3 = 0.86602540378443864676;
    double2 t1 = cadd(v[1], v[2]);
    double2 t2 = make_double2(v[0].x + c3 * t1.x, v[0].y + c3 * t1.y);
    double2 u = mul_si(csub(v[1], v[2]), s * s3);
    v[0] = cadd(v[0], t1); v[1] = cadd(t2, u); v[2] = csub(t2, u);
}
template <>
__device__ __forceinline__ void dft<5>(double2* v, double s) {
    const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;
    const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;
    double2 a1 = cadd(v[1], v[4]), b1 = csub(v[1], v[4]);
    double2 a2 = cadd(v[2], v[3]), b2 = csub(v[2], v[3]);
    double2 m1 = make_double2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
    double2 m2 = make_double2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
    double2 u1 = mul_si(make_double2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y), s);
    double2 u2 = mul_si(make_double2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y), s);
    v[0] = cadd(v[0], cadd(a1, a2));
    v[1] = cadd(m1, u1); v[4] = csub(m1, u1);
    v[2] = cadd(m2, u2); v[3] = csub(m2, u2);
}

// powers w^1..w^(R-1) with multiplication depth <= 3
template <int R>
__device__ __forceinline__ void twiddle_powers(double2 w1, double2* w) {
    w[1] = w1;
    if (R > 2) w[2] = cmul(w1, w1);
    if (R > 3) w[3] = cmul(w[2], w1);
    if (R > 4) w[4] = cmul(w[2], w[2]);
    if (R > 5) w[5] = cmul(w[4], w1);
    if (R > 6) w[6] = cmul(w[3], w[3]);
    if (R > 7) w[7] = cmul(w[4], w[3]);
}

// the two grid rows one complex transform reads or writes (real part <-> primary latitude, imaginary part <-> mirror)
struct RowIO {
    double* o1; double* o2;            // rows of the primary / mirror latitude (o2 unused without a mirror)
    int64_t gs_lon;
    bool live, mirror;
};

}  // namespace fft
}  // namespace shb
