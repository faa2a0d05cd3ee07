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

#define SHB200_VERSION 200

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
       SHB200_FLAG_NO_PTABLE = 16,  /* never keep a device-resident P_nm table: fused on-the-fly recursion kernels */
       SHB200_FLAG_NO_POLAR_SKIP = 32 /* table kernels visit every (n, m, latitude) triple, also where |P_nm| < 1e-24 (see
                                         shb200_plan_info.legendre_visited_*) */ };

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
    int32_t ptable_budget_mib; /* largest P_nm table the plan may keep in HBM, MiB; 0 = default: half of the device memory, at most what is free less 8 GiB */
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
    /* Fraction of the (degree, order, latitude-row) triples the Legendre stage of a full-width call actually visits.  The
     * table kernels skip, per order and 32-row tile, the leading 16-degree blocks in which every |P_nm| < 1e-24 (towards
     * the poles P_nm ~ sin^m(theta) stays negligible until n ~ m/sin(theta): the "polar optimisation" of SHTns); a term that
     * small changes no result above 1e-12 of the field maximum.  1.0 for the other kernel families / with
     * SHB200_FLAG_NO_POLAR_SKIP.  Roofline figures must count executed work only: F_sym x this fraction. */
    double legendre_visited_synthesis;
    double legendre_visited_analysis;
} shb200_plan_info;

/* Per-call options of the four transform entry points (NULL = defaults).  Nothing a call changes is kept on the plan:
 * plans are immutable after creation and may be shared by threads (calls on one plan are serialised by a per-plan mutex
 * for the duration of the launch sequence; the plan's workspace is handed from stream to stream through an event).
 *   degree_scale : fused isotropic kernel: per-degree factor applied while the coefficients are packed
 *                  (synthesis input) or unpacked (analysis output) -- what shxarray's isotropic operators do on the host
 *                  before/after a transform (Gaussian filter, gravity functionals stokes<->tws/geoid/uplift:
 *                  IsoKernelBase.__call__ src/shxarray/kernels/isokernelbase.py:75-92).  HOST array of nmax+1 doubles indexed
 *                  by degree, copied by the call.  c*scale[n] is one IEEE multiplication, bit-identical to scaling first.
 *   shard_rank/world : multi-GPU m-order sharding of ANALYSIS: integrate only the orders
 *                  m = rank, rank+world, ... (interleaved: work per order is proportional to nmax-m+1) and write exact
 *                  zeros for every other coefficient, so that gathering (or summing) the outputs of the `world` ranks
 *                  gives the full result bit-identically to the unsharded transform.  world <= 1: off.  Ignored by synthesis. */
typedef struct shb200_call_opts {
    int32_t struct_size;        /* = sizeof(shb200_call_opts) */
    int32_t shard_rank;
    int32_t shard_world;
    int32_t reserved;
    const double* degree_scale;
} shb200_call_opts;

int shb200_version(void);
const char* shb200_last_error(void);

int shb200_plan_create(const shb200_plan_desc* desc, shb200_plan** out);
void shb200_plan_destroy(shb200_plan* plan);
int shb200_plan_get_info(const shb200_plan* plan, shb200_plan_info* info);

/* Per-stage device times (ms) of the most recent synthesis (ms[0..2] = pack, legendre, longitude) and analysis
 * (ms[3..5] = longitude, legendre, unpack) call on a plan created with SHB200_FLAG_TIMING.  CUDA events recorded on
 * the launching stream; the call synchronises that stream's last event.  Used by bench.py for the roofline. */
int shb200_plan_stage_times(shb200_plan* plan, double* ms, int32_t n);

/* Comma-separated names of the kernels launched by the most recent transform call on this plan (tests and smoke() assert
 * the kernel family -- e.g. "k_pack_n,k_leg_syn_tab<4>,k_lon_syn_fft" -- so that a silent fallback cannot pass).  Copies at most
 * n-1 characters + NUL into buf; returns the full length. */
int shb200_plan_last_kernels(const shb200_plan* plan, char* buf, int32_t n);

/* Synthesis  f(b, point) = sum_k C(b,k) * Ybar_{n_k m_k}(point)      [synthesis.pyx:175-189]
 *   coef element (b,k)          at coef[b*cs_b + k*cs_nm]             (device pointer, element strides)
 *   grid element (b,ilat,ilon)  at grid[b*gs_b + ilat*gs_lat + ilon*gs_lon]; point mode: (b,i) at grid[b*gs_b + i*gs_lat]
 * Asynchronous on `stream` (a cudaStream_t); inputs are not modified.  1 <= batch <= max_batch; a call with fewer fields
 * than the plan was made for works on the leading 2*ceil8(batch) columns of the plan's workspace only, so its cost follows
 * `batch`, not `max_batch`.  A plan serves one stream at a time (its workspace is shared by consecutive calls). */
int shb200_synthesis(shb200_plan* plan, int32_t batch, const double* coef, int64_t cs_b, int64_t cs_nm,
                     double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, void* stream, const shb200_call_opts* opts);

/* Analysis  C(b, n^2+n+m) = sum_ij w_i * Ybar_nm(lon_j, lat_i) * f(b,i,j)   [analysis.pyx:125-139]
 *   output is always the full triangle nmin=0 in nm order (analysis.pyx:29), element (b,k) at coef[b*cs_b + k*cs_nm]. */
int shb200_analysis(shb200_plan* plan, int32_t batch, const double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon,
                    double* coef, int64_t cs_b, int64_t cs_nm, void* stream, const shb200_call_opts* opts);

/* Same two calls on HOST buffers: host->device copy, transform, device->host copy, synchronous.
 * This is the call the reference-facing plugin makes for numpy-backed DataArrays.  Any non-negative element strides.
 *   - Only the elements the call owns are read / written: arrays with holes (a strided `out=` view, one chunk of a
 *     batch-fastest array) are gathered / scattered run by run, never copied as a flat span.
 *   - Page-locked buffers (cudaHostAlloc / cudaHostRegister / shb200_host_alloc) are DMA-ed directly; pageable memory
 *     goes through a ring of pinned bounce buffers filled / drained by a small team of host threads
 *     (SHB200_HOST_THREADS, default min(8, cores); pieces of SHB200_BOUNCE_MIB = 8 MiB) so that host memcpy, H2D, kernels and D2H overlap.
 *   - When the batch axis is the slowest axis of both arrays, batches of >= 16 fields are cut into up to 4 chunks that
 *     flow through three streams (copy-in, kernels, copy-out): both PCIe directions run at once. */
int shb200_synthesis_host(shb200_plan* plan, int32_t batch, const double* coef, int64_t cs_b, int64_t cs_nm,
                          double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, const shb200_call_opts* opts);
int shb200_analysis_host(shb200_plan* plan, int32_t batch, const double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon,
                         double* coef, int64_t cs_b, int64_t cs_nm, const shb200_call_opts* opts);

/* Synthesis and analysis of the same batch in ONE call: coef -> grid (delivered to the host array `grid`) -> coef_out.
 * Equivalent to shb200_synthesis_host followed by shb200_analysis_host on `grid` -- the analysis input is read back from
 * the host array, so the bytes on the wire are those of the two calls -- but software-pipelined over chunks of the batch so
 * that the grid of chunk k+1 travels to the host while the grid of chunk k travels back (both PCIe directions busy).
 * `grid` must be page-locked (shb200_host_alloc / cudaHostRegister); the plan needs quadrature weights. */
int shb200_roundtrip_host(shb200_plan* plan, int32_t batch, const double* coef, int64_t cs_b, int64_t cs_nm,
                          double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon,
                          double* coef_out, int64_t os_b, int64_t os_nm,
                          const shb200_call_opts* syn_opts, const shb200_call_opts* ana_opts);

/* Page-locked host memory from the library's pool (cudaHostAlloc, portable).  The engine hands its numpy OUTPUT arrays out
 * of this pool so that results are DMA-ed straight into the array the caller receives (no bounce copy, no first-touch page
 * faults); blocks go back to the pool with shb200_host_free and are reused.  shb200_host_trim releases cached blocks. */
int shb200_host_alloc(int64_t bytes, void** out);
int shb200_host_free(void* ptr);
int shb200_host_trim(void);

/* Fused round trip used by engine.multiply (shtns.py:394-420 / shlib.pyx:140-152):
 *   out = analysis_planA( synthesis_planS(a) * synthesis_planS(b) )  with the grids kept on the device. */
int shb200_multiply(shb200_plan* syn_a, shb200_plan* syn_b, shb200_plan* ana, int32_t batch,
                    const double* coef_a, int64_t as_b, int64_t as_nm, const double* coef_b, int64_t bs_b, int64_t bs_nm,
                    double* coef_out, int64_t os_b, int64_t os_nm, void* stream);

/* Grid-space operator  out = analysis_planA( mask x synthesis_planS(coef) ), the grid kept on the device:  the ocean
 * function of the sea-level equation applied to a load (replaces SpectralSeaLevelSolver.oceanf, spectralsealevel.py:97-104,
 * an O(nsh^2) product-to-sum matrix limited to nmax <= 150), basin masks (geom/polygons.py), any pointwise weight.
 * mask: ONE device-resident [nlat, nlon] field, element (i,j) at mask[i*ms_lat + j*ms_lon], shared by the batch.
 * syn_opts / ana_opts: per-degree scales fused into either transform (Love-number kernels: geoid, uplift). */
int shb200_apply_mask(shb200_plan* syn, shb200_plan* ana, int32_t batch, const double* coef, int64_t cs_b, int64_t cs_nm,
                      const double* mask, int64_t ms_lat, int64_t ms_lon, double* coef_out, int64_t os_b, int64_t os_nm,
                      void* stream, const shb200_call_opts* syn_opts, const shb200_call_opts* ana_opts);

/* ---- data formats either side of the transform, on the device --------------------------------------
 * shb200_gfc_parse: the record block of an ICGEM ("gfc", D exponents allowed: io/icgem.py:94-128) or GRACE GSM ("GRCOF2",
 * io/gsmv6.py:66-90) coefficient file -> dense nm vector cnm[(nmax+1)^2] (position n^2+n+m; sine terms at -m), optional
 * sigcnm (with_sigma: columns 6/7) and `present` counts per position (a count > 1 is a duplicated record).  text: the file's
 * bytes on the device.  Records with n > nmax are skipped.  Decimal conversion is correctly rounded (identical to strtod /
 * Python float()); lines this kernel cannot decide are NOT stored: their byte offsets go to fallback_offsets (first
 * fallback_capacity of them) for the host to convert.  stats (device, uint64[4]) = records stored, records skipped (n > nmax),
 * lines handed back, reserved.  Outputs are zeroed by the call. */
int shb200_gfc_parse(const char* text, int64_t nbytes, const char* record_key, int32_t d_exponent, int32_t nmax, int32_t with_sigma,
                     double* cnm, double* sigcnm, int32_t* present, uint64_t* stats, int64_t* fallback_offsets, int32_t fallback_capacity,
                     void* stream);
/* tokens text[start[i] .. stop[i]) -> doubles (ok[i] = 0: not decided here); the converter of shb200_gfc_parse on its own */
int shb200_parse_floats(const char* text, const int64_t* start, const int64_t* stop, int64_t ntokens, double* out, uint8_t* ok, void* stream);
/* (n, m, value) records in any order -> dense nm vectors: coef[b*cs_b + (n^2+n+m)*cs_nm] = values[b*vs_b + k*vs_k] */
int shb200_scatter_nm(const int32_t* n, const int32_t* m, const double* values, int64_t vs_b, int64_t vs_k, int64_t nrecords, int32_t batch,
                      int32_t nmax, double* coef, int64_t cs_b, int64_t cs_nm, void* stream);
/* Polygons -> 0/1 grid masks (geom/polygons.py:47-93: grid points contained in each polygon, planar lon/lat), ready for
 * shb200_analysis.  All rings of all polygons in vx/vy (device); ring r = vertices ring_start[r] .. ring_start[r+1]-1 (closed
 * implicitly); polygon p = rings poly_ring[p] .. poly_ring[p+1]-1, even-odd rule (holes, multi-part).  wrap_lon: grid longitudes
 * are mapped to [-180, 180) first (:76-81).  mask element (p,i,j) at mask[p*ms_p + i*ms_lat + j*ms_lon]. */
int shb200_polygon_mask(const double* vx, const double* vy, const int32_t* ring_start, const int32_t* poly_ring, int32_t npoly,
                        const double* lon, const double* lat, int32_t nlat, int32_t nlon, int32_t wrap_lon,
                        double* mask, int64_t ms_p, int64_t ms_lat, int64_t ms_lon, void* stream);

/* ---- multi-GPU exchange formats (one process per GPU; the collective itself is the caller's: ncclAllGather) ----------
 * m-order-sharded analysis (BASELINE config 4): rank r of W integrates the orders m = r, r+W, ... and
 * emits only those, packed m-major:  packed[b*ps_b + off_j + 2(n - m_j) + cs],  m_j = r + jW,
 * off_j = 2 [ j(nmax+1-r) - W j(j-1)/2 ],  cs = 0 cosine / 1 sine (the m = 0 sine entries are zeros).
 * shb200_shard_size = elements per field of rank r's shard.  opts->shard_rank / shard_world select the shard
 * (degree_scale is applied here).  After gathering the W shards (rank r's at packed_all + r*rank_stride),
 * shb200_unpack_shards writes the nm vector of analysis.pyx:29 (position n^2+n+m) -- bit-identical to the unsharded call. */
int64_t shb200_shard_size(int32_t nmax, int32_t rank, int32_t world);
int shb200_analysis_packed(shb200_plan* plan, int32_t batch, const double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon,
                           double* packed, int64_t ps_b, void* stream, const shb200_call_opts* opts);
int shb200_unpack_shards(int32_t nmax, int32_t world, int32_t batch, const double* packed_all, int64_t rank_stride, int64_t ps_b,
                         double* coef, int64_t cs_b, int64_t cs_nm, void* stream);
/* Latitude-sharded synthesis: the gathered buffer src[r*rank_stride + b*src_b + i*nlon + j] holds row i
 * of rank r; row_of_slot[r*rmax + i] (device, int32) is its latitude index in the full grid or -1 for padding. */
int shb200_scatter_rows(int32_t batch, int32_t world, int32_t rmax, int32_t nlon, const double* src, int64_t rank_stride, int64_t src_b,
                        const int32_t* row_of_slot, double* dst, int64_t dst_b, int64_t dst_lat, void* stream);

/* Peer-memory variant of the latitude-sharded exchange (one process per GPU of one NVLink / NVSwitch box): every rank owns a
 * [batch][nlat][nlon] grid allocated with shb200_peer_alloc, publishes its 64-byte IPC handle (cudaIpcGetMemHandle) to the
 * other ranks by any means (torch.distributed.all_gather_object in distributed.py), maps theirs with shb200_peer_open, and
 * after synthesising its own rows calls shb200_push_rows: ONE kernel stores each of its rows at its final place in EVERY rank's
 * grid (peer stores over NVLink; dst[w] = base pointer of rank w's grid, a HOST array of `world` device pointers, world <= 16).
 * No gathered staging buffer, no second pass (shb200_scatter_rows), no NCCL all-gather; the ranks only need a barrier before
 * they read their grid (distributed.py: a stream-ordered one-element all-reduce).  src row i of field b at
 * src[b*src_b + i*nlon]; rows_dev[i] (device, int32) = its latitude index. */
int shb200_peer_alloc(int64_t bytes, void** ptr, void* handle64);
int shb200_peer_open(const void* handle64, void** ptr);
int shb200_peer_close(void* ptr);
int shb200_peer_free(void* ptr);
int shb200_push_rows(int32_t batch, int32_t world, int32_t nown, int32_t nlon, const double* src, int64_t src_b, const int32_t* rows_dev,
                     void* const* dst, int64_t dst_b, int64_t dst_lat, void* stream);

/* Gauss-Legendre nodes x (ascending in latitude: x = sin(lat), south to north) and weights (sum = 2). */
int shb200_gauss_nodes(int32_t n, double* x, double* w);

/* Debug/test hook: normalised P_nm table for one latitude computed by the DEVICE recursion,
 * packed like Legendre_nm.hpp:25-29 ((nmax+1)(nmax+2)/2 doubles, host pointer). */
int shb200_debug_legendre(int32_t device, int32_t nmax, double lat_deg, double* pnm_host);

/* Debug/test hook (host only, no GPU): copies a strided host array (nax <= 3 axes, extents `ext`, element strides) to its
 * dense image (to_dense != 0) or back, through the same layout analysis, run decomposition and thread team the host entry
 * points use for pageable memory.  info (optional, 10 values): merged dims, run length, elements, nested?, DMA-able?, span,
 * outermost axis, dense stride of each axis. */
int shb200_debug_host_copy(int32_t nax, const int64_t* ext, const int64_t* strides, double* host, double* dense, int32_t to_dense, int64_t* info);

/* Debug/test hook (host only, no GPU): the chunk sizes (fields per chunk, at most 8 values) the host entry points cut a call of
 * `batch` fields into -- which = 0 synthesis, 1 analysis, 2 round trip; chunkable = the batch axis is the slowest axis of every
 * array of the call.  Returns the number of chunks (> 0) or an error code (< 0). */
int shb200_debug_host_chunks(int32_t batch, int32_t which, int32_t chunkable, int32_t* sizes);

/* Debug/measurement hook: cycle counters of the three-pass longitude kernels (lon_fft3.cu).  enable != 0 switches the
 * counting on (thread 0 of every CTA adds its clock64 deltas), 0 off; the current sums are copied to out (16 values, may be
 * NULL: synthesis [CTAs, assembly, pass 0, pass 1, last pass + row stores], then analysis [CTAs, table fill, first pass + row
 * loads, pass 1, pass 0, spectrum pick-up]) and cleared. */
int shb200_debug_fft_cycles(int32_t enable, int64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* SHB200_H */
