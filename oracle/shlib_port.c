/* TEST INFRASTRUCTURE ONLY -- never imported, linked or executed by the product path
 * (shxarray_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load the library built from this file.
 *
 * Plain-C restatement ("port") of the shlib engine's arithmetic for the synthesis / analysis
 * hot path of ITC-Water-Resources/shxarray v1.3.2.  Each function cites the reference lines it
 * follows.  It is pinned (tests/test_oracle.py) against
 *   - the reference's own known answers (tests/fixtures.py:66,69-75; the paracap analysis grid), and
 *   - oracle/_ref/libshlib_ref.so, i.e. the reference's unmodified headers compiled in place,
 *     for which P_nm tables must agree BIT FOR BIT and transforms to ~1e-15 (BLAS summation order).
 * Compile with -ffp-contract=off: the reference is built without FMA (setup.py:21-32, plain -O3 on x86-64).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* packed position of (n,m), m-major: Legendre_nm.hpp:25-29 */
static size_t tri(int n, int m, int nmax) { return (size_t)m * (nmax + 1) - ((size_t)m * (m + 1)) / 2 + n; }

size_t port_tri_size(int nmax) { return tri(nmax, nmax, nmax) + 1; }

/* square-root tables: Legendre_nm.hpp:68-78.  wnn has nmax+1 entries, wnm port_tri_size(nmax). */
void port_tables(int nmax, double* wnn, double* wnm) {
    wnn[0] = 0.0;
    if (nmax >= 1) wnn[1] = sqrt(3.0);
    for (int n = 2; n <= nmax; ++n) wnn[n] = sqrt((2 * n + 1) / (2.0 * n));
    memset(wnm, 0, sizeof(double) * port_tri_size(nmax));
    for (int m = 0; m <= nmax; ++m)
        for (int n = m + 1; n <= nmax; ++n)
            wnm[tri(n, m, nmax)] = sqrt((2 * n + 1.0) / (n + m) * (2 * n - 1.0) / (n - m));
}

/* All normalised P_nm(costheta): Legendre_nm.hpp:89-132.
 * Column recursion per order with the 1e-280 / 1e+280 rescaling; sectorial carried across orders. */
void port_legendre_t(int nmax, const double* wnn, const double* wnm, double costheta, double* pnm) {
    const double tiny = 1e-280;
    double sint = sqrt(1 - costheta * costheta); /* :93  (pow(x,2) == x*x exactly) */
    double sect = 1.0 / tiny;                    /* :101 */
    pnm[0] = 1.0;                                /* :103 */
    for (int m = 0; m < nmax; ++m) {
        size_t d = tri(m, m, nmax);
        double back2 = tiny;                     /* :108 */
        double back1 = wnm[d + 1] * costheta * back2; /* :111 */
        pnm[d + 1] = back1 * sect;               /* :112 */
        for (int n = m + 2; n <= nmax; ++n) {
            size_t k = d + (size_t)(n - m);
            double cur = wnm[k] * (costheta * back1 - back2 / wnm[k - 1]); /* :117-118 */
            pnm[k] = cur * sect;                 /* :120 */
            back2 = back1;
            back1 = cur;
        }
        sect *= wnn[m + 1] * sint;               /* :127 */
        pnm[tri(m + 1, m + 1, nmax)] = sect * tiny; /* :129-130 */
    }
}

void port_legendre(int nmax, double costheta, double* pnm) {
    double* wnn = (double*)malloc(sizeof(double) * (nmax + 2));
    double* wnm = (double*)malloc(sizeof(double) * port_tri_size(nmax));
    port_tables(nmax, wnn, wnm);
    port_legendre_t(nmax, wnn, wnm, costheta, pnm);
    free(wnn);
    free(wnm);
}

/* Surface harmonics for an arbitrary (n,m) list at one point: Ynm.hpp:106-134.
 * The reference sorts by (m,n) only to evaluate each trig factor once (:99,:126-129); the values
 * do not depend on that order, so here every entry evaluates its own factor. */
typedef struct {
    int nmax, nsh;
    const int *n, *m;
    double *wnn, *wnm, *pnm, *cosm, *sinm;
    double latprev;
} port_ynm_t;

static void ynm_init(port_ynm_t* y, int nsh, const int* n, const int* m) {
    y->nsh = nsh; y->n = n; y->m = m;
    y->nmax = -1;
    for (int i = 0; i < nsh; ++i) if (n[i] > y->nmax) y->nmax = n[i]; /* Ynm.hpp:89-93 */
    y->wnn = (double*)malloc(sizeof(double) * (y->nmax + 2));
    y->wnm = (double*)malloc(sizeof(double) * port_tri_size(y->nmax));
    y->pnm = (double*)calloc(port_tri_size(y->nmax), sizeof(double));
    y->cosm = (double*)malloc(sizeof(double) * (y->nmax + 1));
    y->sinm = (double*)malloc(sizeof(double) * (y->nmax + 1));
    port_tables(y->nmax, y->wnn, y->wnm);
    y->latprev = -1000; /* Ynm.hpp:66 */
}
static void ynm_free(port_ynm_t* y) { free(y->wnn); free(y->wnm); free(y->pnm); free(y->cosm); free(y->sinm); }

/* DIAGNOSTIC SWITCH, not reference behaviour.  The reference evaluates trig(k * lonr) with lonr = lon*pi/180 rounded to
 * double and the product k*lonr rounded again (Ynm.hpp:114,127): at k ~ 2000 that argument is off by up to ~2e-12 rad,
 * which is the largest systematic difference between shlib and any transform that uses exact DFT twiddles (SURVEY.md A.2:
 * 2.3e-12 ... 3.8e-12 per basis function at nmax 2160).  With the switch on, the argument is formed in long double from
 * the longitude in degrees, i.e. the basis functions shlib INTENDS; tests use it to show that the whole residual of the
 * nmax-2160 white-noise analysis case is this argument rounding and nothing else. */
static int g_exact_trig = 0;
void port_set_exact_trig(int on) { g_exact_trig = on; }

static void ynm_set(port_ynm_t* y, double lon, double lat, double* out) {
    if (lat != y->latprev) { /* Ynm.hpp:107-113 */
        double costheta = sin(lat * M_PI / 180.0);
        port_legendre_t(y->nmax, y->wnn, y->wnm, costheta, y->pnm);
        y->latprev = lat;
    }
    double lonr = lon * M_PI / 180.0; /* :114 */
    if (g_exact_trig) {
        const long double d2r = 3.141592653589793238462643383279503L / 180.0L;
        for (int k = 0; k <= y->nmax; ++k) {
            /* k*lon is exact in long double for the grids in use; reduce in degrees before converting */
            long double a = fmodl((long double)k * (long double)lon, 360.0L) * d2r;
            y->cosm[k] = (double)cosl(a); y->sinm[k] = (double)sinl(a);
        }
    } else
    for (int k = 0; k <= y->nmax; ++k) { y->cosm[k] = cos(k * lonr); y->sinm[k] = sin(k * lonr); } /* :127 */
    for (int i = 0; i < y->nsh; ++i) {
        int mm = y->m[i], am = mm < 0 ? -mm : mm;
        double trig = mm < 0 ? y->sinm[am] : y->cosm[am];
        out[i] = trig * y->pnm[tri(y->n[i], am, y->nmax)]; /* :130 */
    }
}

void port_ynm(int nsh, const int* n, const int* m, int npts, const double* lon, const double* lat, double* out) {
    port_ynm_t y; ynm_init(&y, nsh, n, m);
    for (int p = 0; p < npts; ++p) ynm_set(&y, lon[p], lat[p], out + (size_t)p * nsh);
    ynm_free(&y);
}

/* Synthesis: synthesis.pyx:175-189.  out[point, b] = sum_k coef(b,k) * Y_k(point)  (the dgemv).
 * coef(b,k) = coef[b*ld + k] if nm_fast else coef[k*ld + b]  (synthesis.pyx:104-118,142-153).
 * grid!=0: points = (ilat, ilon) row-major; grid==0: nlat points (lonv[i], latv[i]). */
int port_synthesis(int nsh, const int* nv, const int* mv, int auxsize, const double* coef, int nm_fast, int ld,
                   int grid, int nlat, const double* latv, int nlon, const double* lonv, double* outv, int nthreads) {
#ifdef _OPENMP
    if (nthreads < 1) nthreads = omp_get_max_threads();
#endif
#pragma omp parallel num_threads(nthreads)
    {
        port_ynm_t y; ynm_init(&y, nsh, nv, mv);
        double* yv = (double*)malloc(sizeof(double) * nsh);
#pragma omp for
        for (int ilat = 0; ilat < nlat; ++ilat) {
            int npt = grid ? nlon : 1;
            for (int j = 0; j < npt; ++j) {
                double lon = grid ? lonv[j] : lonv[ilat];
                ynm_set(&y, lon, latv[ilat], yv);
                double* o = outv + ((size_t)ilat * npt + j) * auxsize;
                for (int b = 0; b < auxsize; ++b) {
                    double acc = 0.0;
                    if (nm_fast) { const double* c = coef + (size_t)b * ld; for (int k = 0; k < nsh; ++k) acc += c[k] * yv[k]; }
                    else { for (int k = 0; k < nsh; ++k) acc += coef[(size_t)k * ld + b] * yv[k]; }
                    o[b] = acc;
                }
            }
        }
        free(yv);
        ynm_free(&y);
    }
    return 0;
}

/* Analysis: analysis.pyx:125-139.  out[b, k] += alpha(lat) * Y_k(point) * f(point, b)  (the dger),
 * alpha = weight * cos(lat * (pi/180)) (:121,:129), full triangle n-major m=-n..n (analysis.pyx:29, sh_indexing.py:66).
 * grid element (ilat,ilon,b) = gridv[ilat*s_lat + ilon*s_lon + b*s_aux].  out [auxsize, nsh] must be zeroed by the caller. */
int port_analysis(int nmax, int auxsize, const double* gridv, long s_lat, long s_lon, long s_aux, int nlat,
                  const double* latv, int nlon, const double* lonv, double weight, double* outv, int nthreads) {
    int nsh = (nmax + 1) * (nmax + 1);
    int* nv = (int*)malloc(sizeof(int) * nsh);
    int* mv = (int*)malloc(sizeof(int) * nsh);
    for (int n = 0, k = 0; n <= nmax; ++n)
        for (int m = -n; m <= n; ++m, ++k) { nv[k] = n; mv[k] = m; }
    const double d2r = M_PI / 180;
#ifdef _OPENMP
    if (nthreads < 1) nthreads = omp_get_max_threads();
#endif
#pragma omp parallel num_threads(nthreads)
    {
        port_ynm_t y; ynm_init(&y, nsh, nv, mv);
        double* yv = (double*)malloc(sizeof(double) * nsh);
        double* part = (double*)calloc((size_t)auxsize * nsh, sizeof(double)); /* per-thread partial instead of the lock */
#pragma omp for
        for (int ilat = 0; ilat < nlat; ++ilat) {
            double alpha = weight * cos(latv[ilat] * d2r);
            for (int ilon = 0; ilon < nlon; ++ilon) {
                ynm_set(&y, lonv[ilon], latv[ilat], yv);
                const double* f = gridv + (size_t)ilat * s_lat + (size_t)ilon * s_lon;
                for (int b = 0; b < auxsize; ++b) {
                    double t = alpha * f[(size_t)b * s_aux];
                    double* o = part + (size_t)b * nsh;
                    for (int k = 0; k < nsh; ++k) o[k] += yv[k] * t;
                }
            }
        }
#pragma omp critical
        for (size_t i = 0; i < (size_t)auxsize * nsh; ++i) outv[i] += part[i];
        free(part);
        free(yv);
        ynm_free(&y);
    }
    free(nv);
    free(mv);
    return 0;
}
