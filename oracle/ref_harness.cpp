// TEST INFRASTRUCTURE ONLY -- never imported, linked or executed by the product path
// (shxarray_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load the library built from this file.
//
// oracle/_ref/libshlib_ref.so: the reference's OWN arithmetic.  This translation unit
// #includes the reference headers where they lie (-I/root/reference/src/builtin_backend:
// Ynm.hpp, Legendre_nm.hpp -- nothing is copied into this repo) and drives them with the two
// hot loops of the shlib engine:
//   synthesis : src/builtin_backend/synthesis.pyx:175-189  (per point Ynm.set + dgemv)
//   analysis  : src/builtin_backend/analysis.pyx:125-139   (per point Ynm.set + dger under an OMP lock)
// BLAS is the one shlib itself calls (scipy.linalg.cython_blas -> scipy's bundled OpenBLAS,
// symbols scipy_dgemv_/scipy_dger_), resolved at run time with dlopen so no link-time path is baked in.
// The Cython module itself cannot be imported in this image (xarray is absent).
#include <dlfcn.h>
#include <omp.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "Ynm.hpp"  // pulls Legendre_nm.hpp

typedef void (*dgemv_t)(const char*, const int*, const int*, const double*, const double*, const int*,
                        const double*, const int*, const double*, double*, const int*);
typedef void (*dger_t)(const int*, const int*, const double*, const double*, const int*, const double*,
                       const int*, double*, const int*);
static dgemv_t g_dgemv = nullptr;
static dger_t g_dger = nullptr;

extern "C" {

// returns 0 on success
int ref_load_blas(const char* path) {
    void* h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) { fprintf(stderr, "ref_load_blas: %s\n", dlerror()); return 1; }
    g_dgemv = (dgemv_t)dlsym(h, "scipy_dgemv_");
    g_dger = (dger_t)dlsym(h, "scipy_dger_");
    typedef void (*setthr_t)(int);
    setthr_t st = (setthr_t)dlsym(h, "scipy_openblas_set_num_threads");
    if (st) st(1);  // BLAS is called from inside the OpenMP region
    return (g_dgemv && g_dger) ? 0 : 2;
}

int ref_max_threads() { return omp_get_max_threads(); }

// Legendre_nm<double>(nmax).set(costheta, pnm)  -- Legendre_nm.hpp:62-132
void ref_legendre(int nmax, double costheta, double* pnm) {
    Legendre_nm<double> leg(nmax);
    leg.set(costheta, pnm);
}

// Ynm_cpp<double>(size,n,m).set(lon,lat) for a list of points -- Ynm.hpp:84-134
void ref_ynm(int nsh, const int* n, const int* m, int npts, const double* lon, const double* lat, double* out) {
    Ynm_cpp<double> ynm(nsh, n, m);
    for (int p = 0; p < npts; ++p) {
        ynm.set(lon[p], lat[p]);
        std::memcpy(out + (size_t)p * nsh, ynm.data(), sizeof(double) * nsh);
    }
}

// synthesis.pyx:142-189.  coef is the 2-D view `inval` [auxsize, shsize]:
//   nm_fast != 0 : element (b,k) at coef[b*ld + k]  -> trans='T', lda=shsize(=ld)   (synthesis.pyx:142-147)
//   nm_fast == 0 : element (b,k) at coef[k*ld + b]  -> trans='N', lda=ld            (synthesis.pyx:148-153)
// out: [npoints, auxsize] C-contiguous, points = ilat*nlon+ilon (grid) or idx (points).
int ref_synthesis(int nsh, const int* nv, const int* mv, int auxsize, const double* coef, int nm_fast, int ld,
                  int grid, int nlat, const double* latv, int nlon, const double* lonv, double* outv, int nthreads) {
    if (!g_dgemv) return 1;
    char trans; int m, n, lda = ld, incx = 1, incy = 1;
    double alpha = 1.0, beta = 0.0;
    if (nm_fast) { trans = 'T'; m = nsh; n = auxsize; }
    else { trans = 'N'; m = auxsize; n = nsh; }
    if (nthreads < 1) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
    {
        Ynm_cpp<double> ynm(nsh, nv, mv);
        if (grid) {
#pragma omp for
            for (int ilat = 0; ilat < nlat; ++ilat) {
                for (int ilon = 0; ilon < nlon; ++ilon) {
                    ynm.set(lonv[ilon], latv[ilat]);
                    g_dgemv(&trans, &m, &n, &alpha, coef, &lda, ynm.data(), &incx, &beta,
                            outv + ((size_t)ilat * nlon + ilon) * auxsize, &incy);
                }
            }
        } else {
#pragma omp for
            for (int idx = 0; idx < nlat; ++idx) {
                ynm.set(lonv[idx], latv[idx]);
                g_dgemv(&trans, &m, &n, &alpha, coef, &lda, ynm.data(), &incx, &beta, outv + (size_t)idx * auxsize, &incy);
            }
        }
    }
    return 0;
}

// analysis.pyx:64-139.  grid element (ilat,ilon,b) at gridv[ilat*s_lat + ilon*s_lon + b*s_aux] (element strides),
// which covers the four layouts of analysis.pyx:100-120.  out: [auxsize, nsh] C-contiguous, must be zeroed by the caller
// (analysis.pyx:54 allocates zeros).  weight = |stepx*stepy|/(4 pi) is computed by the caller (analysis.pyx:57-60).
int ref_analysis(int nmax, int auxsize, const double* gridv, long s_lat, long s_lon, long s_aux, int nlat,
                 const double* latv, int nlon, const double* lonv, double weight, double* outv, int nthreads) {
    if (!g_dger) return 1;
    int nsh = (nmax + 1) * (nmax + 1);
    std::vector<int> nv(nsh), mv(nsh);
    {   // sh_indexing.py:66
        int k = 0;
        for (int nn = 0; nn <= nmax; ++nn)
            for (int mm = -nn; mm <= nn; ++mm) { nv[k] = nn; mv[k] = mm; ++k; }
    }
    int m = nsh, n = auxsize, lda = nsh, incx = 1, incy = (int)s_aux;
    const double d2r = M_PI / 180;  // analysis.pyx:121 (np.pi/180)
    if (nthreads < 1) nthreads = omp_get_max_threads();
    omp_lock_t lock;
    omp_init_lock(&lock);
#pragma omp parallel num_threads(nthreads)
    {
        Ynm_cpp<double> ynm(nsh, nv.data(), mv.data());
#pragma omp for
        for (int ilat = 0; ilat < nlat; ++ilat) {
            double alpha = weight * cos(latv[ilat] * d2r);
            for (int ilon = 0; ilon < nlon; ++ilon) {
                ynm.set(lonv[ilon], latv[ilat]);
                omp_set_lock(&lock);
                g_dger(&m, &n, &alpha, ynm.data(), &incx, gridv + (size_t)ilat * s_lat + (size_t)ilon * s_lon, &incy, outv, &lda);
                omp_unset_lock(&lock);
            }
        }
    }
    omp_destroy_lock(&lock);
    return 0;
}

}  // extern "C"
