// Transform planner + C ABI of libshb200.so (host side).
//
// The planner builds, ON THE HOST with the same libm/IEEE operations the reference uses, every
// quantity whose rounding matters for parity with shlib:
//   wnm / wnn square-root tables         Legendre_nm.hpp:68-78
//   costheta = sin(lat*pi/180)           Ynm.hpp:109
//   sinTheta = sqrt(1 - costheta^2)      Legendre_nm.hpp:93
//   sectorial products (1e280 scaling)   Legendre_nm.hpp:101,127-130
//   alpha = weight*cos(lat*(pi/180))     analysis.pyx:121,129
// and decides the execution plan: mirror-latitude pairing, FFT vs dense longitude stage, kernel family.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <unordered_map>

#include "internal.h"

namespace shb {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error '%s' in %s (%s:%d)", cudaGetErrorString(e), what, file, line);
    g_err = buf;
    return e == cudaErrorMemoryAllocation ? SHB200_ENOMEM : SHB200_ECUDA;
}

template <typename T>
static int upload(Plan& p, const std::vector<T>& h, const T** dptr) {
    void* d = nullptr;
    size_t bytes = sizeof(T) * (h.empty() ? 1 : h.size());
    SHB_CUDA(cudaMalloc(&d, bytes));
    p.owned.push_back(d);
    p.workspace_bytes += bytes;
    if (!h.empty()) SHB_CUDA(cudaMemcpy(d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
    *dptr = (const T*)d;
    return 0;
}

static int dev_alloc(Plan& p, size_t bytes, void** out) {
    void* d = nullptr;
    SHB_CUDA(cudaMalloc(&d, bytes ? bytes : 8));
    p.owned.push_back(d);
    p.workspace_bytes += bytes;
    *out = d;
    return 0;
}

void launch_ptab_gen(const Plan& p, int* first_k, cudaStream_t st);

static double ulp_scale(double x) { return std::ldexp(1.0, -52) * std::fmax(std::fabs(x), 1.0); }

// Polar skipping (legendre_tab.cu): turn the per-(order, 32-row tile) first significant degree offsets into first 16-degree
// blocks (in place, on the device array the kernels read), the heaviest-first tile lists of the persistent synthesis kernel
// and the fractions of the (n, m, row) triples each kernel shape still visits.
static int build_polar_schedule(Plan& p, int* first_k_dev) {
    DevTables& t = p.t;
    const int N = p.nmax, R = p.nrows, RT32 = t.RT32;
    std::vector<int> fb((size_t)(N + 1) * RT32);
    SHB_CUDA(cudaMemcpy(fb.data(), first_k_dev, sizeof(int) * fb.size(), cudaMemcpyDeviceToHost));
    for (int m = 0; m <= N; ++m) {
        const int nblk = (N - m + 16) / 16;
        for (int q = 0; q < RT32; ++q) {
            int& v = fb[(size_t)m * RT32 + q];
            v = std::min(v / 16, nblk);            // untouched entries (0x7f7f7f7f: pad rows, all-negligible tiles) -> nblk
        }
    }
    SHB_CUDA(cudaMemcpy(first_k_dev, fb.data(), sizeof(int) * fb.size(), cudaMemcpyHostToDevice));
    t.first_blk = first_k_dev;
    const double all = (double)tri_size(N) * R;
    auto rows_in = [&](int q) { return std::max(0, std::min(32, R - 32 * q)); };
    for (int v = 0; v < 3; ++v) {
        // synthesis tile shapes: NWC = 4, 2, 1 -> 2, 4, 8 sub-tiles of 32 rows (legendre_tab.cu: TS<NWC>)
        const int nwr = 2 << v, nrt = (R + 32 * nwr - 1) / (32 * nwr);
        std::vector<std::pair<int, int>> work((size_t)(N + 1) * nrt);
        double visited = 0.0;
        for (int m = 0; m <= N; ++m) {
            const int nblk = (N - m + 16) / 16;
            for (int rt = 0; rt < nrt; ++rt) {
                int f = nblk, rows = 0;
                for (int q = nwr * rt; q < std::min(RT32, nwr * (rt + 1)); ++q) { f = std::min(f, fb[(size_t)m * RT32 + q]); rows += rows_in(q); }
                work[(size_t)m * nrt + rt] = {nblk - f, m * nrt + rt};
                visited += (double)std::max(0, N - m + 1 - 16 * f) * rows;
            }
        }
        std::stable_sort(work.begin(), work.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.first > b.first; });
        // descriptor per list position: (tile = m * nrt + rt, first chunk to visit) -- one load per tile in the kernel
        std::vector<int2> order(work.size());
        for (size_t i = 0; i < work.size(); ++i) {
            const int tile = work[i].second, m = tile / nrt;
            order[i] = make_int2(tile, (N - m + 16) / 16 - work[i].first);
        }
        int rc = upload(p, order, &p.syn_order[v]);
        if (rc) return rc;
        p.visited_syn[v] = visited / all;
        // analysis stage widths: NCQ = 4, 2, 1 -> stages of 2, 4, 8 blocks (legendre_tab.cu: TA<NCQ>), row tiles of 32
        const int sblk = 2 << v;
        visited = 0.0;
        for (int m = 0; m <= N; ++m)
            for (int q = 0; q < RT32; ++q) visited += (double)std::max(0, N - m + 1 - 16 * sblk * (fb[(size_t)m * RT32 + q] / sblk)) * rows_in(q);
        p.visited_ana[v] = visited / all;
    }
    return 0;
}

static int build_plan(const shb200_plan_desc& d, Plan& p) {
    if (d.nmax < 0 || d.max_batch < 1 || d.nlat < 1 || d.nlon < 1 || !d.lat_deg || !d.lon_deg) {
        set_error("plan_create: invalid sizes or null coordinate arrays");
        return SHB200_EINVAL;
    }
    if (d.mode == SHB200_MODE_POINTS && d.nlat != d.nlon) {
        set_error("plan_create: point mode needs nlat == nlon");
        return SHB200_EINVAL;
    }
    if (d.mode == SHB200_MODE_POINTS && d.quadrature != SHB200_QUAD_NONE) {
        set_error("plan_create: analysis is not defined for point sets");
        return SHB200_EINVAL;
    }
    if (d.quadrature == SHB200_QUAD_LATWEIGHTS && !d.lat_weights) {
        set_error("plan_create: SHB200_QUAD_LATWEIGHTS needs lat_weights");
        return SHB200_EINVAL;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("no CUDA device available: libshb200 has no CPU fallback");
        return SHB200_ECUDA;
    }
    if (d.device < 0 || d.device >= ndev) { set_error("plan_create: bad device ordinal"); return SHB200_EINVAL; }
    SHB_CUDA(cudaSetDevice(d.device));

    p.device = d.device; p.nmax = d.nmax; p.maxB = d.max_batch;
    p.Bp = d.max_batch <= 4 ? 4 : (d.max_batch + 7) / 8 * 8;   // columns come in blocks of 2*GS = (cos,sin) x GS fields
    p.NC = 2 * p.Bp;
    p.mode = d.mode; p.nlat = d.nlat; p.nlon = d.nlon; p.quadrature = d.quadrature; p.flags = d.flags;
    const int N = d.nmax;
    const int64_t T = tri_size(N);

    // ---- nm list -> slot map -----------------------------------------------------------------
    std::vector<int64_t> slot_src((size_t)T * 2, -1);
    std::vector<int64_t> dup_slot, dup_src, fwd;
    std::vector<int32_t> degk;
    if (d.nsh < 0 || d.nsh > ((int64_t)1 << 31)) { set_error("plan_create: nsh out of range"); return SHB200_EINVAL; }
    if (d.nsh > 0) {
        if (!d.n || !d.m) { set_error("plan_create: nsh > 0 but n/m arrays are null"); return SHB200_EINVAL; }
        fwd.assign((size_t)d.nsh, -1);
        degk.assign(d.n, d.n + d.nsh);
        p.nsh = d.nsh;
        p.full_triangle = (d.nsh == (int64_t)(N + 1) * (N + 1));
        for (int64_t k = 0; k < d.nsh; ++k) {
            int n = d.n[k], m = d.m[k], am = m < 0 ? -m : m;
            if (n < 0 || n > N || am > n) { set_error("plan_create: (n,m) entry out of range"); return SHB200_EINVAL; }
            int64_t s = tri(n, am, N) * 2 + (m < 0 ? 1 : 0);
            if (slot_src[s] < 0) { slot_src[s] = k; fwd[k] = s; }
            else { dup_slot.push_back(s); dup_src.push_back(k); }
            if (p.full_triangle && k != (int64_t)n * n + n + m) p.full_triangle = false;
        }
    } else {
        p.nsh = (int64_t)(N + 1) * (N + 1);
        p.full_triangle = true;
        fwd.assign((size_t)p.nsh, -1);
        degk.resize((size_t)p.nsh);
        for (int n = 0; n <= N; ++n)
            for (int m = -n; m <= n; ++m) degk[(int64_t)n * n + n + m] = n;
        for (int n = 0; n <= N; ++n)
            for (int m = -n; m <= n; ++m) {
                int am = m < 0 ? -m : m;
                slot_src[tri(n, am, N) * 2 + (m < 0 ? 1 : 0)] = (int64_t)n * n + n + m;
                fwd[(int64_t)n * n + n + m] = tri(n, am, N) * 2 + (m < 0 ? 1 : 0);
            }
    }

    // ---- Legendre tables (bit-identical to Legendre_nm.hpp:68-78) ------------------------------
    std::vector<double> wnn(N + 2, 0.0), wnm((size_t)T, 0.0), bnm((size_t)T, 0.0);
    if (N >= 1) wnn[1] = sqrt(3.0);
    for (int n = 2; n <= N; ++n) wnn[n] = sqrt((2 * n + 1) / (2.0 * n));
    for (int m = 0; m <= N; ++m)
        for (int n = m + 1; n <= N; ++n)
            wnm[tri(n, m, N)] = sqrt((2 * n + 1.0) / (n + m) * (2 * n - 1.0) / (n - m));
    // correctly rounded reciprocals: the device divides by wnm with q=p*r; q+=fma(-q,w,p)*r (Markstein), which
    // reproduces the IEEE quotient of Legendre_nm.hpp:117-118 (checked on every recursion step at nmax=2160)
    for (int m = 0; m <= N; ++m)
        for (int n = m + 1; n <= N; ++n) bnm[tri(n, m, N)] = 1.0 / wnm[tri(n, m, N)];

    // ---- latitude rows: mirror pairing ---------------------------------------------------------
    std::vector<double> cost_lat(d.nlat);
    for (int i = 0; i < d.nlat; ++i) cost_lat[i] = sin(d.lat_deg[i] * M_PI / 180.0);  // Ynm.hpp:109
    std::vector<int> i1, i2;
    {
        std::vector<char> used(d.nlat, 0);
        std::unordered_multimap<uint64_t, int> bybits;
        const bool pair_ok = d.mode == SHB200_MODE_GRID && !(d.flags & SHB200_FLAG_NO_SYMMETRY);
        if (pair_ok)
            for (int i = 0; i < d.nlat; ++i) {
                uint64_t b; memcpy(&b, &cost_lat[i], 8);
                bybits.emplace(b, i);
            }
        for (int i = 0; i < d.nlat; ++i) {
            if (used[i]) continue;
            used[i] = 1;
            int mate = -1;
            if (pair_ok && cost_lat[i] != 0.0) {
                double neg = -cost_lat[i];
                uint64_t b; memcpy(&b, &neg, 8);
                auto range = bybits.equal_range(b);
                for (auto it = range.first; it != range.second; ++it)
                    if (!used[it->second]) { mate = it->second; break; }
            }
            if (mate >= 0) used[mate] = 1;
            i1.push_back(i); i2.push_back(mate);
        }
    }
    // rows from the poles to the equator (stable in the caller's order among equal |cos(theta)|): a 32-row tile of the table
    // kernels then spans one latitude band, which is what makes polar skipping (first_blk) effective for any latitude order
    {
        std::vector<int> ord(i1.size());
        for (size_t r = 0; r < ord.size(); ++r) ord[r] = (int)r;
        std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return fabs(cost_lat[i1[a]]) > fabs(cost_lat[i1[b]]); });
        std::vector<int> a1(i1.size()), a2(i1.size());
        for (size_t r = 0; r < ord.size(); ++r) { a1[r] = i1[ord[r]]; a2[r] = i2[ord[r]]; }
        i1.swap(a1); i2.swap(a2);
    }
    p.nrows = (int)i1.size();
    const int R = p.nrows;
    std::vector<double> cost(R), w1(R, 0.0), w2(R, 0.0);
    for (int r = 0; r < R; ++r) cost[r] = cost_lat[i1[r]];
    if (d.quadrature != SHB200_QUAD_NONE) {
        const double d2r = M_PI / 180;  // analysis.pyx:121
        auto wrow = [&](int i) {
            return d.quadrature == SHB200_QUAD_SHLIB ? d.weight * cos(d.lat_deg[i] * d2r)  // analysis.pyx:129
                                                     : d.weight * d.lat_weights[i];
        };
        for (int r = 0; r < R; ++r) { w1[r] = wrow(i1[r]); w2[r] = i2[r] >= 0 ? wrow(i2[r]) : 0.0; }
    }
    // sectorial multipliers per (order, row): Legendre_nm.hpp:93,101,127-130
    std::vector<double> sect((size_t)(N + 1) * R), pmm((size_t)(N + 1) * R);
    {
        const double tiny = 1e-280;
        for (int r = 0; r < R; ++r) {
            double c = cost[r];
            double sint = sqrt(1 - c * c);
            double s = 1.0 / tiny;
            sect[r] = s; pmm[r] = 1.0;
            for (int m = 0; m < N; ++m) {
                s *= wnn[m + 1] * sint;
                sect[(size_t)(m + 1) * R + r] = s;
                pmm[(size_t)(m + 1) * R + r] = s * tiny;
            }
        }
    }

    // ---- longitude stage decision ------------------------------------------------------------
    p.K = d.nlon;
    p.lon_fft = false;
    double lon0 = d.lon_deg[0];
    if (d.mode == SHB200_MODE_GRID && !(d.flags & SHB200_FLAG_DENSE_LON) && d.nlon >= 2 && lon_fft_supported(d.nlon, p.Bp)) {
        bool uniform = true;
        long double step = 360.0L / d.nlon;
        for (int j = 0; j < d.nlon && uniform; ++j) {
            long double ideal = (long double)lon0 + j * step;
            if (fabsl((long double)d.lon_deg[j] - ideal) > 8.0L * ulp_scale(d.lon_deg[j])) uniform = false;
        }
        if (uniform) p.lon_fft = true;
    }
    std::vector<double2> tw, phase;
    std::vector<double> lonr(d.nlon);
    for (int j = 0; j < d.nlon; ++j) lonr[j] = d.lon_deg[j] * M_PI / 180.0;  // Ynm.hpp:114
    if (p.lon_fft) {
        tw.resize(p.K);
        const long double twopi = 6.283185307179586476925286766559L;
        for (int k = 0; k < p.K; ++k) {
            // exact octant reduction keeps the table accurate to <1 ulp
            long double a = twopi * (long double)k / (long double)p.K;
            tw[k] = make_double2((double)cosl(a), (double)-sinl(a));
        }
        if (p.K % 4 == 0) { tw[p.K / 4] = make_double2(0.0, -1.0); tw[p.K / 2] = make_double2(-1.0, 0.0); tw[3 * p.K / 4] = make_double2(0.0, 1.0); }
        else if (p.K % 2 == 0) tw[p.K / 2] = make_double2(-1.0, 0.0);
        phase.resize(N + 1);
        long double l0 = (long double)lon0 * (3.141592653589793238462643383279503L / 180.0L);
        for (int m = 0; m <= N; ++m) {
            long double a = fmodl((long double)m * l0, twopi);
            phase[m] = make_double2((double)cosl(a), (double)sinl(a));
        }
    }

    // ---- kernel family -----------------------------------------------------------------------
    p.dmma = !(d.flags & SHB200_FLAG_NO_DMMA);

    // ---- upload ------------------------------------------------------------------------------
    int rc;
    DevTables& t = p.t;
    if ((rc = upload(p, wnm, &t.wnm))) return rc;
    if ((rc = upload(p, bnm, &t.rinv))) return rc;
    if ((rc = upload(p, cost, &t.cost))) return rc;
    if ((rc = upload(p, sect, &t.sect))) return rc;
    if ((rc = upload(p, pmm, &t.pmm))) return rc;
    if ((rc = upload(p, i1, &t.row_i1))) return rc;
    if ((rc = upload(p, i2, &t.row_i2))) return rc;
    if ((rc = upload(p, w1, &t.w1))) return rc;
    if ((rc = upload(p, w2, &t.w2))) return rc;
    if ((rc = upload(p, tw, &t.tw))) return rc;
    if ((rc = upload(p, phase, &t.phase))) return rc;
    if ((rc = upload(p, lonr, &t.lonr))) return rc;
    if (p.lon_fft) {
        std::vector<int> perm;
        lon_fft_perm(p.K, p.Bp, perm);
        if ((rc = upload(p, perm, &p.fft_perm))) return rc;
    }
    int fft3_group = 0;
    if (p.lon_fft && lon_fft3_plan(p.K, p.Bp, &fft3_group)) {
        std::vector<int> perm;
        lon_fft3_perm(p.K, perm);
        if ((rc = upload(p, perm, &p.fft3_perm))) return rc;
    }
    if ((rc = upload(p, slot_src, &t.slot_src))) return rc;
    p.ndup = (int64_t)dup_slot.size();
    if ((rc = upload(p, dup_slot, &p.dup_slot))) return rc;
    if ((rc = upload(p, dup_src, &p.dup_src))) return rc;
    if ((rc = upload(p, fwd, &p.fwd_slot))) return rc;
    if ((rc = upload(p, degk, &p.deg_of_k))) return rc;
    for (int w = 0; w < 2; ++w) {
        void* q;
        if ((rc = dev_alloc(p, sizeof(double) * (N + 1), &q))) return rc;
        p.dscale[w] = (double*)q;
    }
    t.nmax = N; t.nrows = R; t.nlat = d.nlat; t.nlon = d.nlon; t.Bp = p.Bp; t.NC = p.NC; t.GS = fft3_group ? fft3_group : (p.Bp >= 8 ? 8 : 4);
    t.GSH = t.GS == 8 ? 3 : 2;
    if (t.GS != 4 && t.GS != 8) { set_error("plan_create: internal error: field group size must be 4 or 8"); return SHB200_EINVAL; }
    t.Cs = p.NC + 2; t.Gs = p.NC; t.phase_unit = (p.lon_fft && lon0 == 0.0) ? 1 : 0; t.pt_base = nullptr; t.RT32 = 0; t.m0 = 0; t.mstride = 1; t.first_blk = nullptr;

    void* ptr;
    // 16 / 64 slack rows: the tensor-core kernels fetch whole 16-degree / 32-row blocks, also past the last order
    p.points_fused = d.mode == SHB200_MODE_POINTS && points_fused_wanted(R);
    const size_t cm_elems = ((size_t)T + 16) * t.Cs, g_elems = p.points_fused ? (size_t)64 * t.Gs : ((size_t)(N + 1) * R * 2 + 64) * t.Gs;
    if ((rc = dev_alloc(p, sizeof(double) * cm_elems, &ptr))) return rc;
    p.Cm = (double*)ptr;
    if ((rc = dev_alloc(p, sizeof(double) * g_elems, &ptr))) return rc;
    p.G = (double*)ptr;
    SHB_CUDA(cudaMemset(p.Cm, 0, sizeof(double) * cm_elems));
    SHB_CUDA(cudaMemset(p.G, 0, sizeof(double) * g_elems));
    SHB_CUDA(cudaEventCreateWithFlags(&p.ev_done, cudaEventDisableTiming));
    // device-resident P_nm table for the tensor-core kernels (legendre_tab.cu) when it fits the budget
    if (p.dmma && !(d.flags & SHB200_FLAG_NO_PTABLE) && p.NC >= 32 && d.mode == SHB200_MODE_GRID) {
        p.Rpad = (R + 63) / 64 * 64;
        // tiled layout: [order][16-degree block][32-row tile][16][34] doubles (legendre_tab.cu)
        std::vector<int64_t> ptb(N + 2, 0);
        for (int m = 0; m <= N; ++m) ptb[m + 1] = ptb[m] + (N - m + 16) / 16;
        // + 8 sub-tiles of slack: the tall tile shapes of the narrow-column kernels fetch up to 8 sub-tiles at once, also at the very end
        const size_t pt_elems = (size_t)ptb[N + 1] * (p.Rpad / 32) * (16 * 34) + 8 * (16 * 34);
        // default budget: half of the device's memory (90 GiB on a 180 GB B200), and never more than what is free now less
        // 8 GiB of headroom for the caller's grids.  nmax 2160 on shlib's own grid (shlib.pyx:87-92: 5760 x 11520, 2880 mirror
        // pairs) needs 57 GB -- the table holds the northern rows only -- and 20 GB on the 5' grid; streaming it beats
        // recomputing P_nm per call (the fused kernels) by 5-8 x for multi-field batches.
        size_t free_b = 0, total_b = 0;
        SHB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const double mib = 1024.0 * 1024.0;
        const double auto_mib = std::min((double)total_b / mib / 2.0, (double)free_b / mib - 8192.0);
        const double budget_mib = d.ptable_budget_mib > 0 ? (double)d.ptable_budget_mib : auto_mib;
        const double need_mib = (double)pt_elems * 8.0 / (1024.0 * 1024.0);
        if (need_mib <= budget_mib) {
            if ((rc = upload(p, ptb, &t.pt_base))) return rc;
            t.RT32 = p.Rpad / 32;
            // the table is an accelerator, not a requirement: if the allocation fails after all (fragmentation, another
            // process took the memory) the plan falls back to the fused-recursion kernels
            void* tab = nullptr;
            if (cudaMalloc(&tab, sizeof(double) * pt_elems) == cudaSuccess) {
                p.owned.push_back(tab);
                p.workspace_bytes += sizeof(double) * pt_elems;
                p.Ptab = (double*)tab;
                SHB_CUDA(cudaMemset(p.Ptab, 0, sizeof(double) * pt_elems));
                int* first_k = nullptr;      // per (order, 32-row tile): first degree offset n - m with |P_nm| >= POLAR_EPS
                const size_t nfk = (size_t)(N + 1) * t.RT32;
                if (!(d.flags & SHB200_FLAG_NO_POLAR_SKIP)) {
                    if ((rc = dev_alloc(p, sizeof(int) * nfk, &ptr))) return rc;
                    first_k = (int*)ptr;
                    SHB_CUDA(cudaMemset(first_k, 0x7f, sizeof(int) * nfk));
                }
                launch_ptab_gen(p, first_k, 0);
                SHB_CUDA(cudaDeviceSynchronize());
                if (first_k && (rc = build_polar_schedule(p, first_k))) return rc;
            } else {
                cudaGetLastError();      // clear the allocation error
                t.RT32 = 0;
            }
        }
    }
    // tiny-batch plans (recursion kernels): polar-skipping seeds for the synthesis kernel
    if (p.NC < 32 && d.mode == SHB200_MODE_GRID && !(d.flags & SHB200_FLAG_NO_POLAR_SKIP) && N >= 256) {
        if ((rc = build_syn_seeds(p, &p.seed_visited_syn))) return rc;
        if (d.quadrature != SHB200_QUAD_NONE && p.NC <= 24 && (rc = build_ana_seeds(p, &p.seed_visited_ana))) return rc;     // single-field analysis kernel
    }
    if (d.flags & SHB200_FLAG_TIMING)
        for (int i = 0; i < 8; ++i) SHB_CUDA(cudaEventCreate(&p.ev[i]));
    p.launches_syn = 3 + (p.ndup ? 1 : 0);
    p.launches_ana = 3;
    return 0;
}

static void free_plan(Plan& p) {
    drop_graphs(p);
    if (p.cap_stream) { cudaStreamDestroy(p.cap_stream); p.cap_stream = nullptr; }
    cudaSetDevice(p.device);
    free_host_io(p);
    for (void* d : p.owned) cudaFree(d);
    p.owned.clear();
    if (p.ana_scratch) { cudaFree(p.ana_scratch); p.ana_scratch = nullptr; p.ana_scratch_bytes = 0; }
    if (p.mul_grid) cudaFree(p.mul_grid);
    if (p.ev_done) cudaEventDestroy(p.ev_done);
    for (int i = 0; i < 8; ++i) if (p.ev[i]) cudaEventDestroy(p.ev[i]);
}

int parse_opts(const Plan& p, const shb200_call_opts* o, CallOpts& out) {
    out = CallOpts();
    if (!o) return 0;
    if (o->struct_size != (int32_t)sizeof(shb200_call_opts)) { set_error("call opts: struct_size mismatch"); return SHB200_EINVAL; }
    if (o->shard_world > 1) {
        if (o->shard_rank < 0 || o->shard_rank >= o->shard_world) { set_error("call opts: need 0 <= shard_rank < shard_world"); return SHB200_EINVAL; }
        out.shard_rank = o->shard_rank; out.shard_world = o->shard_world;
    }
    out.degree_scale = o->degree_scale;
    (void)p;
    return 0;
}

// Everything one call changes on the plan, applied for the duration of the host-side launch sequence (the caller holds
// p.mu; kernel arguments are copied when a kernel is launched) and restored on exit:
//  * effective width: a call with fewer fields than the plan was made for runs on the leading columns only -- fields come
//    in groups of 8 with their cosine and sine columns adjacent, so `batch` fields occupy the first 2*ceil8(batch) columns
//    of every Cm / G row; row strides (Cs, Gs) stay those of the plan;
//  * the analysis order shard and the per-degree scale of this call (shb200_call_opts);
//  * workspace hand-over: a call on another stream than the previous one waits for that call's last kernel.
struct CallScope {
    Plan& p;
    int Bp0, NC0;
    cudaStream_t st;
    bool events = true;
    CallScope(Plan& plan, int batch, cudaStream_t stream, const CallOpts& o, int which) : p(plan), Bp0(plan.Bp), NC0(plan.NC), st(stream) {
        if (p.Bp > 8) {
            const int bp = std::min(p.Bp, std::max((batch + 7) / 8 * 8, o.min_fields));
            p.Bp = bp; p.NC = 2 * bp; p.t.Bp = bp; p.t.NC = 2 * bp;
        }
        p.cur_batch = batch;
        p.t.m0 = which == 1 ? o.shard_rank : 0;
        p.t.mstride = which == 1 ? o.shard_world : 1;
        events = !o.no_events;
        if (events && p.has_last && p.last_stream != st) cudaStreamWaitEvent(st, p.ev_done, 0);
        p.dscale_on[0] = p.dscale_on[1] = false;
        if (o.degree_scale) {
            // pageable source: the runtime stages the nmax+1 values before returning, so the caller's array is free again
            cudaMemcpyAsync(p.dscale[which], o.degree_scale, sizeof(double) * (p.nmax + 1), cudaMemcpyHostToDevice, st);
            p.dscale_on[which] = true;
        }
        p.trace.clear();
    }
    ~CallScope() {
        if (events) {
            cudaEventRecord(p.ev_done, st);
            p.last_stream = st; p.has_last = true;
        }
        p.Bp = Bp0; p.NC = NC0; p.t.Bp = Bp0; p.t.NC = NC0;
        p.t.m0 = 0; p.t.mstride = 1;
        p.dscale_on[0] = p.dscale_on[1] = false;
    }
};

// ---- CUDA graphs for launch-latency-bound transforms ------------------------------------------------------------------
// BASELINE configs 1 and 5 (nmax 60 / 256, one to four fields) and the spectral products of the sea-level iteration are three to
// ten kernels of 7-40 us each: their cost is launch latency, not work.  The second call with the same arguments (same plan(s),
// batch, pointers, strides, order shard) on a small plan captures its launch sequence on a private stream
// (cudaStreamCaptureModeThreadLocal) and instantiates it; from then on the call is ONE cudaGraphLaunch on the caller's stream.
// The first call runs directly: it performs every first-use initialisation (function attributes, scratch allocations) that
// must not happen under capture.  Calls with a fused degree scale (a host array read at call time), timed plans
// (SHB200_FLAG_TIMING) and large plans are never captured; SHB200_GRAPHS=0 switches the mechanism off.
static bool graphs_on() {
    static const bool v = [] { const char* e = getenv("SHB200_GRAPHS"); return !e || atoi(e) != 0; }();
    return v;
}

bool graph_eligible(const Plan& p, const CallOpts& o) {
    if (!graphs_on() || (p.flags & SHB200_FLAG_TIMING) || o.degree_scale || o.no_events) return false;
    const int64_t nfull = (int64_t)(p.nmax + 1) * (p.nmax + 1);
    return nfull * p.Bp <= ((int64_t)1 << 21) && (int64_t)p.nlat * (p.mode == SHB200_MODE_POINTS ? 1 : p.nlon) * p.Bp <= ((int64_t)1 << 23);
}

void drop_graphs(Plan& p) {
    for (GraphEntry& e : p.graphs)
        if (e.exec) cudaGraphExecDestroy(e.exec);
    p.graphs.clear();
}

int run_graphed(Plan& owner, const GraphKey& key, cudaStream_t st, Plan* const* touched, int ntouched,
                const std::function<int(cudaStream_t, bool)>& body) {
    GraphEntry* e = nullptr;
    for (GraphEntry& g : owner.graphs)
        if (g.key == key) { e = &g; break; }
    if (!e) {
        if (owner.graphs.size() >= 16) {      // least recently used entry makes room
            size_t old = 0;
            for (size_t i = 1; i < owner.graphs.size(); ++i)
                if (owner.graphs[i].used < owner.graphs[old].used) old = i;
            if (owner.graphs[old].exec) cudaGraphExecDestroy(owner.graphs[old].exec);
            owner.graphs.erase(owner.graphs.begin() + old);
        }
        GraphEntry fresh;
        fresh.key = key;
        fresh.used = ++owner.graph_clock;
        owner.graphs.push_back(fresh);
        return body(st, false);               // first call with these arguments: direct
    }
    e->used = ++owner.graph_clock;
    if (e->bad) return body(st, false);
    if (!e->exec) {
        if (!owner.cap_stream && cudaStreamCreateWithFlags(&owner.cap_stream, cudaStreamNonBlocking) != cudaSuccess) {
            cudaGetLastError(); e->bad = true; return body(st, false);
        }
        if (cudaStreamBeginCapture(owner.cap_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError(); e->bad = true; return body(st, false);
        }
        const int rc = body(owner.cap_stream, true);
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(owner.cap_stream, &graph);
        cudaGraphExec_t exec = nullptr;
        e = nullptr;                          // the launch sequence may have dropped the plan's graphs (a scratch buffer grew)
        for (GraphEntry& g : owner.graphs)
            if (g.key == key) { e = &g; break; }
        if (!e || rc || ce != cudaSuccess || !graph || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            if (e) e->bad = true;
            return body(st, false);
        }
        cudaGraphDestroy(graph);
        e->exec = exec;
        e->trace = owner.trace;
    }
    for (int i = 0; i < ntouched; ++i) {
        Plan& q = *touched[i];
        if (q.has_last && q.last_stream != st) cudaStreamWaitEvent(st, q.ev_done, 0);
    }
    SHB_CUDA(cudaGraphLaunch(e->exec, st));
    for (int i = 0; i < ntouched; ++i) {
        Plan& q = *touched[i];
        cudaEventRecord(q.ev_done, st);
        q.last_stream = st; q.has_last = true;
    }
    owner.trace = e->trace;
    return 0;
}

static int check_batch(const Plan& p, int batch) {
    if (batch < 1 || batch > p.maxB) { set_error("batch must be in [1, max_batch] of the plan"); return SHB200_EINVAL; }
    return 0;
}

int do_synthesis(Plan& p, int batch, const double* coef, int64_t cs_b, int64_t cs_nm, double* grid, int64_t gs_b,
                 int64_t gs_lat, int64_t gs_lon, cudaStream_t st, const CallOpts& o) {
    int rc = check_batch(p, batch);
    if (rc) return rc;
    SHB_CUDA(cudaSetDevice(p.device));
    const bool tm = p.flags & SHB200_FLAG_TIMING;
    CallScope scope(p, batch, st, o, 0);
    if (tm) cudaEventRecord(p.ev[0], st);
    launch_pack(p, batch, coef, cs_b, cs_nm, st);
    if (tm) cudaEventRecord(p.ev[1], st);
    if (p.points_fused) {
        launch_syn_points(p, batch, grid, gs_b, gs_lat, st);     // Legendre and longitude stage of a point set in one kernel
        if (tm) cudaEventRecord(p.ev[2], st);
    } else {
        launch_legendre_syn(p, st);
        if (tm) cudaEventRecord(p.ev[2], st);
        rc = launch_lon_syn(p, batch, grid, gs_b, gs_lat, gs_lon, st);
        if (rc) return rc;
    }
    if (tm) { cudaEventRecord(p.ev[3], st); p.timed_syn = true; }
    SHB_CUDA(cudaGetLastError());
    return 0;
}

int do_analysis(Plan& p, int batch, const double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, double* coef,
                int64_t cs_b, int64_t cs_nm, cudaStream_t st, const CallOpts& o) {
    if (p.quadrature == SHB200_QUAD_NONE) { set_error("analysis: plan was created without quadrature weights"); return SHB200_EINVAL; }
    int rc = check_batch(p, batch);
    if (rc) return rc;
    SHB_CUDA(cudaSetDevice(p.device));
    const bool tm = p.flags & SHB200_FLAG_TIMING;
    CallScope scope(p, batch, st, o, 1);
    if (tm) cudaEventRecord(p.ev[4], st);
    rc = launch_lon_ana(p, batch, grid, gs_b, gs_lat, gs_lon, st);
    if (rc) return rc;
    if (tm) cudaEventRecord(p.ev[5], st);
    if (!launch_legendre_ana(p, st)) { set_error("analysis: out of device memory for the recursion state of a grid with more than 8192 latitude rows"); return SHB200_ENOMEM; }
    if (tm) cudaEventRecord(p.ev[6], st);
    launch_unpack(p, batch, coef, cs_b, cs_nm, st);
    if (tm) { cudaEventRecord(p.ev[7], st); p.timed_ana = true; }
    SHB_CUDA(cudaGetLastError());
    return 0;
}

// analysis whose output is the rank's compact m-major shard (shard.cu) instead of the zero-padded nm vector
int do_analysis_packed(Plan& p, int batch, const double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, double* packed, int64_t ps_b,
                       cudaStream_t st, const CallOpts& o) {
    if (p.quadrature == SHB200_QUAD_NONE) { set_error("analysis: plan was created without quadrature weights"); return SHB200_EINVAL; }
    int rc = check_batch(p, batch);
    if (rc) return rc;
    SHB_CUDA(cudaSetDevice(p.device));
    const bool tm = p.flags & SHB200_FLAG_TIMING;
    CallScope scope(p, batch, st, o, 1);
    if (tm) cudaEventRecord(p.ev[4], st);
    rc = launch_lon_ana(p, batch, grid, gs_b, gs_lat, gs_lon, st);
    if (rc) return rc;
    if (tm) cudaEventRecord(p.ev[5], st);
    if (!launch_legendre_ana(p, st)) { set_error("analysis: out of device memory for the recursion state of a grid with more than 8192 latitude rows"); return SHB200_ENOMEM; }
    if (tm) cudaEventRecord(p.ev[6], st);
    launch_pack_shard(p, batch, packed, ps_b, st);
    if (tm) { cudaEventRecord(p.ev[7], st); p.timed_ana = true; }
    SHB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace shb

using namespace shb;

extern "C" {

int shb200_version(void) { return SHB200_VERSION; }
const char* shb200_last_error(void) { return g_err.c_str(); }

int shb200_plan_create(const shb200_plan_desc* desc, shb200_plan** out) {
    if (!desc || !out || desc->struct_size != (int32_t)sizeof(shb200_plan_desc)) {
        set_error("plan_create: null argument or struct_size mismatch");
        return SHB200_EINVAL;
    }
    shb200_plan* h = new shb200_plan();
    static std::atomic<uint64_t> next_serial{1};
    h->p.serial = next_serial.fetch_add(1);
    int rc = build_plan(*desc, h->p);
    if (rc) { free_plan(h->p); delete h; *out = nullptr; return rc; }
    *out = h;
    return 0;
}

void shb200_plan_destroy(shb200_plan* plan) {
    if (!plan) return;
    free_plan(plan->p);
    delete plan;
}

int shb200_plan_get_info(const shb200_plan* plan, shb200_plan_info* info) {
    if (!plan || !info) { set_error("plan_get_info: null argument"); return SHB200_EINVAL; }
    const Plan& p = plan->p;
    info->nmax = p.nmax; info->max_batch = p.maxB; info->padded_batch = p.Bp;
    info->nlat = p.nlat; info->nlon = p.nlon; info->nrows = p.nrows;
    info->lon_fft = p.lon_fft ? 1 : 0; info->dmma = p.dmma ? (p.Ptab ? 2 : 1) : 0;
    info->nsh = p.nsh; info->nsh_full = (int64_t)(p.nmax + 1) * (p.nmax + 1);
    info->workspace_bytes = p.workspace_bytes;
    info->launches_synthesis = p.launches_syn; info->launches_analysis = p.launches_ana;
    const int shape = p.NC > 64 ? 0 : (p.NC > 32 ? 1 : 2);      // tile shape of a full-width call (legendre_tab.cu)
    info->legendre_visited_synthesis = p.Ptab && p.t.first_blk ? p.visited_syn[shape] : p.seed_visited_syn;
    info->legendre_visited_analysis = p.Ptab && p.t.first_blk ? p.visited_ana[shape] : p.seed_visited_ana;
    return 0;
}

int shb200_synthesis(shb200_plan* plan, int32_t batch, const double* coef, int64_t cs_b, int64_t cs_nm,
                     double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, void* stream, const shb200_call_opts* opts) {
    if (!plan || !coef || !grid) { set_error("synthesis: null argument"); return SHB200_EINVAL; }
    Plan& p = plan->p;
    CallOpts o;
    int rc = parse_opts(p, opts, o);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(p.mu);
    if (graph_eligible(p, o) && batch >= 1 && batch <= p.maxB) {
        GraphKey k;
        k.kind = 1; k.batch = batch; k.serial[0] = p.serial;
        k.ptr[0] = coef; k.ptr[1] = grid;
        k.stride[0] = cs_b; k.stride[1] = cs_nm; k.stride[2] = gs_b; k.stride[3] = gs_lat; k.stride[4] = gs_lon;
        Plan* touched[1] = {&p};
        SHB_CUDA(cudaSetDevice(p.device));
        return run_graphed(p, k, (cudaStream_t)stream, touched, 1, [&](cudaStream_t s, bool capturing) {
            CallOpts oo = o;
            oo.no_events = capturing;
            return do_synthesis(p, batch, coef, cs_b, cs_nm, grid, gs_b, gs_lat, gs_lon, s, oo);
        });
    }
    return do_synthesis(p, batch, coef, cs_b, cs_nm, grid, gs_b, gs_lat, gs_lon, (cudaStream_t)stream, o);
}

int shb200_analysis(shb200_plan* plan, int32_t batch, const double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon,
                    double* coef, int64_t cs_b, int64_t cs_nm, void* stream, const shb200_call_opts* opts) {
    if (!plan || !coef || !grid) { set_error("analysis: null argument"); return SHB200_EINVAL; }
    Plan& p = plan->p;
    CallOpts o;
    int rc = parse_opts(p, opts, o);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(p.mu);
    if (graph_eligible(p, o) && p.quadrature != SHB200_QUAD_NONE && batch >= 1 && batch <= p.maxB) {
        GraphKey k;
        k.kind = 2; k.batch = batch; k.serial[0] = p.serial; k.shard_rank = o.shard_rank; k.shard_world = o.shard_world;
        k.ptr[0] = grid; k.ptr[1] = coef;
        k.stride[0] = gs_b; k.stride[1] = gs_lat; k.stride[2] = gs_lon; k.stride[3] = cs_b; k.stride[4] = cs_nm;
        Plan* touched[1] = {&p};
        SHB_CUDA(cudaSetDevice(p.device));
        return run_graphed(p, k, (cudaStream_t)stream, touched, 1, [&](cudaStream_t s, bool capturing) {
            CallOpts oo = o;
            oo.no_events = capturing;
            return do_analysis(p, batch, grid, gs_b, gs_lat, gs_lon, coef, cs_b, cs_nm, s, oo);
        });
    }
    return do_analysis(p, batch, grid, gs_b, gs_lat, gs_lon, coef, cs_b, cs_nm, (cudaStream_t)stream, o);
}

int shb200_plan_last_kernels(const shb200_plan* plan, char* buf, int32_t n) {
    if (!plan || !buf || n < 1) { set_error("plan_last_kernels: invalid argument"); return SHB200_EINVAL; }
    Plan& p = const_cast<Plan&>(plan->p);
    std::lock_guard<std::mutex> lock(p.mu);
    const size_t len = std::min(p.trace.size(), (size_t)n - 1);
    memcpy(buf, p.trace.data(), len);
    buf[len] = 0;
    return (int)p.trace.size();
}

int shb200_plan_stage_times(shb200_plan* plan, double* ms, int32_t n) {
    if (!plan || !ms || n < 6) { set_error("plan_stage_times: need a plan and room for 6 values"); return SHB200_EINVAL; }
    Plan& p = plan->p;
    if (!(p.flags & SHB200_FLAG_TIMING)) { set_error("plan_stage_times: plan was created without SHB200_FLAG_TIMING"); return SHB200_EINVAL; }
    std::lock_guard<std::mutex> lock(p.mu);
    SHB_CUDA(cudaSetDevice(p.device));
    for (int i = 0; i < 6; ++i) ms[i] = 0.0;
    float f;
    if (p.timed_syn) {
        SHB_CUDA(cudaEventSynchronize(p.ev[3]));
        for (int i = 0; i < 3; ++i) { SHB_CUDA(cudaEventElapsedTime(&f, p.ev[i], p.ev[i + 1])); ms[i] = f; }
    }
    if (p.timed_ana) {
        SHB_CUDA(cudaEventSynchronize(p.ev[7]));
        for (int i = 0; i < 3; ++i) { SHB_CUDA(cudaEventElapsedTime(&f, p.ev[4 + i], p.ev[5 + i])); ms[3 + i] = f; }
    }
    return 0;
}

// Gauss-Legendre nodes/weights by Newton iteration on P_n in long double.
int shb200_gauss_nodes(int32_t n, double* x, double* w) {
    if (n < 1 || !x || !w) { set_error("gauss_nodes: invalid argument"); return SHB200_EINVAL; }
    const long double pi = 3.141592653589793238462643383279503L;
    for (int i = 0; i < (n + 1) / 2; ++i) {
        long double z = cosl(pi * (i + 0.75L) / (n + 0.5L));  // i-th root from the north
        long double pp = 1;
        for (int it = 0; it < 100; ++it) {
            long double p1 = 1, p2 = 0;
            for (int j = 1; j <= n; ++j) {
                long double p3 = p2; p2 = p1;
                p1 = ((2 * j - 1) * z * p2 - (j - 1) * p3) / j;
            }
            pp = n * (z * p1 - p2) / (z * z - 1);
            long double dz = p1 / pp;
            z -= dz;
            if (fabsl(dz) < 1e-19L) break;
        }
        {   // derivative at the converged root
            long double p1 = 1, p2 = 0;
            for (int j = 1; j <= n; ++j) { long double p3 = p2; p2 = p1; p1 = ((2 * j - 1) * z * p2 - (j - 1) * p3) / j; }
            pp = n * (z * p1 - p2) / (z * z - 1);
        }
        long double wi = 2 / ((1 - z * z) * pp * pp);
        // ascending latitude: south first
        x[i] = (double)-z; x[n - 1 - i] = (double)z;
        w[i] = (double)wi; w[n - 1 - i] = (double)wi;
    }
    if (n % 2) x[n / 2] = 0.0;
    return 0;
}

}  // extern "C"
