/* shb200.h -- C ABI of the B200-native spherical-harmonic synthesis/analysis engine.
 *
 * Drop-in boundary for ONE path of ITC-Water-Resources/shxarray: grid/point synthesis and grid
 * analysis of real, 4pi-normalised spherical-harmonic coefficients over a batch of fields.  The
 * reference has no C ABI for this path (it is a Cython class inside one extension module); the
 * entry points below are what a reference-side binding would call instead of
 *
 *   Synthesis._apply_dgemv   src/builtin_backend/synthesis.pyx:76-189   -> shb200_synthesis[_host]
 *   Analysis._apply_ana      src/builtin_backend/analysis.pyx:64-139    -> shb200_analysis[_host]
 *   Ynm_cpp / Legendre_nm    src/builtin_backend/Ynm.hpp:84-134, Legendre_nm.hpp:62-132
 *                                                                       -> (inside the plan: tables + device recursion)
 *   shtns grid generation    src/shtns_backend/shtns.py:345-363 (Gauss nodes) -> shb200_gauss_nodes
 *
 * Conventions (identical to the reference): degrees, geocentric latitude, real harmonics with
 * m<0 = sine terms (sh_indexing.py:52), no Condon-Shortley phase, 4pi normalisation.
 * All floating-point data are FP64.  Plain pointers and sizes only; no torch/xarray types.
 * Every function returns 0 on success or a negative SHB200_E* code; shb200_last_error() gives the
 * message for the calling thread.  There is NO CPU fallback: without a CUDA device every compute
 * call fails with SHB200_ECUDA.
 */
#ifndef SHB200_H
#define SHB200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SHB200_VERSION 100

enum { SHB200_OK = 0, SHB200_EINVAL = -1, SHB200_ECUDA = -2, SHB200_ENOMEM = -3, SHB200_EUNSUPPORTED = -4 };

/* shb200_plan_desc.mode */
enum { SHB200_MODE_GRID = 0,   /* lat x lon tensor grid (synthesis.pyx:179-183) */
       SHB200_MODE_POINTS = 1  /* nlat == nlon point pairs   (synthesis.pyx:186-189) */ };

/* shb200_plan_desc.quadrature (analysis only) */
enum { SHB200_QUAD_NONE = 0,      /* plan is synthesis-only */
       SHB200_QUAD_SHLIB = 1,     /* row weight = weight * cos(lat*pi/180)   (analysis.pyx:57-60,129) */
       SHB200_QUAD_LATWEIGHTS = 2 /* row weight = weight * lat_weights[i]    (e.g. Gauss-Legendre) */ };

/* shb200_plan_desc.flags */
enum { SHB200_FLAG_NO_SYMMETRY = 1, /* do not pair mirror latitudes */
       SHB200_FLAG_DENSE_LON = 2,   /* force the dense trig longitude stage (no FFT) */
       SHB200_FLAG_NO_DMMA = 4,     /* force the CUDA-core (DFMA) Legendre kernels */
       SHB200_FLAG_TIMING = 8,      /* record CUDA events around every stage (see shb200_plan_stage_times) */
       SHB200_FLAG_NO_PTABLE = 16   /* never keep a device-resident P_nm table: fused on-the-fly recursion kernels */ };

typedef struct shb200_plan shb200_plan;

typedef struct shb200_plan_desc {
    int32_t struct_size;  /* = sizeof(shb200_plan_desc) */
    int32_t device;       /* CUDA device ordinal */
    int32_t nmax;         /* maximum degree (analysis output / upper bound of the nm list) */
    int32_t max_batch;    /* largest number of fields per call (product of the non-nm dims, synthesis.pyx:89) */
    int64_t nsh;          /* length of the nm list; 0 => full triangle in nm order n-major m=-n..n (sh_indexing.py:66) */
    const int32_t* n;     /* host, nsh entries: degree of each input coefficient (any subset/order, Ynm.hpp:84-101) */
    const int32_t* m;     /* host, nsh entries: order, m<0 = sine */
    int32_t mode;         /* SHB200_MODE_* */
    int32_t nlat;         /* number of latitudes (or points) */
    const double* lat_deg;/* host */
    int32_t nlon;         /* number of longitudes (== nlat in point mode) */
    int32_t quadrature;   /* SHB200_QUAD_* */
    const double* lon_deg;/* host */
    double weight;        /* quadrature prefactor, e.g. |dlon*dlat|/(4 pi) in radians (analysis.pyx:57-60) */
    const double* lat_weights; /* host, nlat entries, SHB200_QUAD_LATWEIGHTS only */
    int32_t flags;        /* SHB200_FLAG_* */
    int32_t ptable_budget_mib; /* largest P_nm table the plan may keep in HBM, MiB; 0 = default: a quarter of the device memory, at most what is free less 4 GiB */
} shb200_plan_desc;

typedef struct shb200_plan_info {
    int32_t nmax, max_batch, padded_batch;
    int32_t nlat, nlon, nrows;   /* nrows = mirror-paired latitude rows actually iterated */
    int32_t lon_fft;             /* 1 = FFT longitude stage, 0 = dense trig sums */
    int32_t dmma;                /* 0 = CUDA-core kernels, 1 = tensor-core kernels with fused recursion, 2 = tensor-core kernels + P_nm table */
    int64_t nsh;                 /* length of the synthesis nm list */
    int64_t nsh_full;            /* (nmax+1)^2 = analysis output length */
    int64_t workspace_bytes;     /* device memory owned by the plan */
    int32_t launches_synthesis;  /* kernels launched per synthesis call */
    int32_t launches_analysis;   /* kernels launched per analysis call */
} shb200_plan_info;

int shb200_version(void);
const char* shb200_last_error(void);

int shb200_plan_create(const shb200_plan_desc* desc, shb200_plan** out);
void shb200_plan_destroy(shb200_plan* plan);
int shb200_plan_get_info(const shb200_plan* plan, shb200_plan_info* info);

/* Per-stage device times (ms) of the most recent synthesis (ms[0..2] = pack, legendre, longitude) and analysis
 * (ms[3..5] = longitude, legendre, unpack) call on a plan created with SHB200_FLAG_TIMING.  CUDA events recorded on
 * the launching stream; the call synchronises that stream's last event.  Used by bench.py for the roofline. */
int shb200_plan_stage_times(shb200_plan* plan, double* ms, int32_t n);

/* Fused isotropic kernel: per-degree scale applied while the coefficients are packed (which=0,
 * synthesis input) or unpacked (which=1, analysis output).  This is what shxarray's isotropic operators do on the host
 * before/after a transform -- Gaussian filter, gravity functionals stokes<->tws/geoid/uplift: IsoKernelBase.__call__
 * src/shxarray/kernels/isokernelbase.py:75-92 (dain * expanddiag(nm)).  scale: host array of nmax+1 doubles indexed by degree,
 * copied by the call; NULL switches the scale off.  The product c*scale[n] is one IEEE multiplication, i.e. bit-identical to
 * scaling on the host first. */
int shb200_plan_set_degree_scale(shb200_plan* plan, int32_t which, const double* scale, int32_t n);

/* Multi-GPU m-order sharding of ANALYSIS: the plan integrates only the orders m = rank, rank+world,
 * rank+2*world, ... (interleaved: work per order is proportional to nmax-m+1) and writes exact zeros for every other
 * coefficient, so that summing (or gathering) the outputs of the `world` ranks gives the full result bit-identically to the
 * unsharded transform.  rank=0, world=1 restores the default.  Synthesis is unaffected (it shards by latitude or batch). */
int shb200_plan_set_order_shard(shb200_plan* plan, int32_t rank, int32_t world);

/* Synthesis  f(b, point) = sum_k C(b,k) * Ybar_{n_k m_k}(point)      [synthesis.pyx:175-189]
 *   coef element (b,k)          at coef[b*cs_b + k*cs_nm]             (device pointer, element strides)
 *   grid element (b,ilat,ilon)  at grid[b*gs_b + ilat*gs_lat + ilon*gs_lon]; point mode: (b,i) at grid[b*gs_b + i*gs_lat]
 * Asynchronous on `stream` (a cudaStream_t); inputs are not modified.  1 <= batch <= max_batch; a call with fewer fields
 * than the plan was made for works on the leading 2*ceil8(batch) columns of the plan's workspace only, so its cost follows
 * `batch`, not `max_batch`.  A plan serves one stream at a time (its workspace is shared by consecutive calls). */
int shb200_synthesis(shb200_plan* plan, int32_t batch, const double* coef, int64_t cs_b, int64_t cs_nm,
                     double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, void* stream);

/* Analysis  C(b, n^2+n+m) = sum_ij w_i * Ybar_nm(lon_j, lat_i) * f(b,i,j)   [analysis.pyx:125-139]
 *   output is always the full triangle nmin=0 in nm order (analysis.pyx:29), element (b,k) at coef[b*cs_b + k*cs_nm]. */
int shb200_analysis(shb200_plan* plan, int32_t batch, const double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon,
                    double* coef, int64_t cs_b, int64_t cs_nm, void* stream);

/* Same two calls on HOST buffers: host->device copy, transform, device->host copy, synchronous.
 * This is the call the reference-facing plugin makes for numpy-backed DataArrays.  When the batch axis is the slowest
 * axis of both arrays, batches of >= 16 fields are cut into up to 4 chunks that overlap copy-in, kernels and copy-out on
 * three streams (pinned host memory is needed for the copies to be asynchronous; pageable memory works, without overlap). */
int shb200_synthesis_host(shb200_plan* plan, int32_t batch, const double* coef, int64_t cs_b, int64_t cs_nm,
                          double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon);
int shb200_analysis_host(shb200_plan* plan, int32_t batch, const double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon,
                         double* coef, int64_t cs_b, int64_t cs_nm);

/* Fused round trip used by engine.multiply (shtns.py:394-420 / shlib.pyx:140-152):
 *   out = analysis_planA( synthesis_planS(a) * synthesis_planS(b) )  with the grids kept on the device. */
int shb200_multiply(shb200_plan* syn_a, shb200_plan* syn_b, shb200_plan* ana, int32_t batch,
                    const double* coef_a, int64_t as_b, int64_t as_nm, const double* coef_b, int64_t bs_b, int64_t bs_nm,
                    double* coef_out, int64_t os_b, int64_t os_nm, void* stream);

/* Gauss-Legendre nodes x (ascending in latitude: x = sin(lat), south to north) and weights (sum = 2). */
int shb200_gauss_nodes(int32_t n, double* x, double* w);

/* Debug/test hook: normalised P_nm table for one latitude computed by the DEVICE recursion,
 * packed like Legendre_nm.hpp:25-29 ((nmax+1)(nmax+2)/2 doubles, host pointer). */
int shb200_debug_legendre(int32_t device, int32_t nmax, double lat_deg, double* pnm_host);

#ifdef __cplusplus
}
#endif
#endif /* SHB200_H */
