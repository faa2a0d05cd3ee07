// Host-buffer entry points of libshb200.so: shb200_synthesis_host / shb200_analysis_host / shb200_roundtrip_host and the
// page-locked host memory pool (shb200_host_alloc / _free / _trim).
//
// These are the calls the reference-facing plugin makes for numpy-backed DataArrays (what shlib's Synthesis / Analysis
// operators receive: synthesis.pyx:104-140, analysis.pyx:100-120 -- plain host arrays of any layout).  At the BASELINE
// configs they are PCIe-bound (cfg3: 0.27 + 0.57 GB per direction against ~1.3 ms of kernels), so the work here is moving
// bytes:
//   * the device staging arrays are DENSE copies of the caller's arrays in the caller's axis order; only the elements a
//     call owns are read from / written to host memory (a strided `out=` view or one chunk of a batch-fastest array has
//     holes: they are skipped run by run, never overwritten);
//   * page-locked host memory (cudaHostAlloc / cudaHostRegister / shb200_host_alloc) is DMA-ed directly
//     (cudaMemcpyAsync / cudaMemcpy2DAsync); pageable memory goes through a ring of pinned bounce buffers that a small team
//     of host threads fills (H2D side) or drains (D2H side) piece by piece while the previous piece is on the wire;
//   * when the batch axis is the slowest axis of both arrays the batch is cut into up to 4 chunks of whole 8-field groups
//     that flow through three streams -- copy-in, kernels, copy-out -- so both PCIe directions run at once and the
//     kernels hide behind the copies.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <thread>
#include <unordered_map>
#include <unistd.h>

#include "internal.h"

namespace shb {

// ---- a small persistent team of host threads for the bounce copies ---------------------------------------------------
class Team {
  public:
    explicit Team(int nworkers) : pid_(getpid()) {
        for (int i = 0; i < nworkers; ++i) workers_.emplace_back([this, i] { loop(i); });
        for (auto& w : workers_) w.detach();          // the team lives as long as the process
        nworkers_ = nworkers;
    }
    bool usable() const { return pid_ == getpid(); }  // a fork()ed child has no worker threads
    // fn(lo, hi) over [0, n) in nworkers + 1 slices (the caller takes one); returns when all slices are done
    void run(int64_t n, const std::function<void(int64_t, int64_t)>& fn) {
        std::lock_guard<std::mutex> job(job_mu_);
        const int parts = nworkers_ + 1;
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn; n_ = n; parts_ = parts; pending_ = nworkers_; ++gen_;
        }
        cv_.notify_all();
        slice(nworkers_, parts, n, fn);
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }

  private:
    static void slice(int i, int parts, int64_t n, const std::function<void(int64_t, int64_t)>& fn) {
        // slices aligned to 8 elements (one 64-byte line)
        const int64_t per = ((n + parts - 1) / parts + 7) / 8 * 8;
        const int64_t lo = std::min(n, per * i), hi = std::min(n, per * (i + 1));
        if (hi > lo) fn(lo, hi);
    }
    void loop(int i) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int64_t, int64_t)>* fn;
            int64_t n; int parts;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_; fn = fn_; n = n_; parts = parts_;
            }
            slice(i, parts, n, *fn);
            {
                std::lock_guard<std::mutex> lk(mu_);
                --pending_;
            }
            done_.notify_all();
        }
    }
    std::vector<std::thread> workers_;
    int nworkers_ = 0;
    pid_t pid_;
    std::mutex job_mu_, mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int64_t, int64_t)>* fn_ = nullptr;
    int64_t n_ = 0;
    int parts_ = 1, pending_ = 0;
    uint64_t gen_ = 0;
};

static int host_threads() {
    static const int n = [] {
        const char* e = getenv("SHB200_HOST_THREADS");
        // measured on the B200 box (tools/e2e_sweep.py, profiles/r02_e2e_sweep.jsonl): 266 MB of pageable coefficients in,
        // 570 MB out: 23 / 17 / 16.6 / 15.3 ms with 2 / 4 / 8 / 16 threads -- beyond 4 the PCIe link (both directions busy) is
        // the limit; a pageable 570 MB grid in: 36 / 23 / 20 / 17 ms
        int v = e ? atoi(e) : std::min(8, (int)std::max(1u, std::thread::hardware_concurrency()));
        return std::max(1, std::min(v, 64));
    }();
    return n;
}

// one team per copy direction: the copy-in pump (the calling thread) and the copy-out pump (a helper thread) run at once
static Team* team(int dir) {
    static std::mutex mu;
    static Team* t[2] = {nullptr, nullptr};
    std::lock_guard<std::mutex> lk(mu);
    if (!t[dir] || !t[dir]->usable()) t[dir] = new Team(std::max(0, host_threads() - 1));
    return t[dir];
}

// ---- strided host arrays -----------------------------------------------------------------------------------------
// Up to three logical axes (extent, host stride in elements), sorted by stride and merged where adjacent axes are
// contiguous with each other.  The dense (device staging) copy has the same axis order without the holes.
struct Layout {
    int nd = 0;
    int64_t n[3] = {1, 1, 1}, hs[3] = {0, 0, 0};
    int64_t total = 1;     // elements
    int64_t run = 1;       // contiguous elements per run on the host side
    int nouter = 0;        // dims above the run
    int64_t span = 1;      // host elements between the first and one past the last owned element
    bool nested = true;    // distinct indices never address the same host element
};

// ext/hs: logical axes.  ds_axis: dense stride of every logical axis (axes of extent 1 get `total`, like an outermost axis);
// outer_axis: the logical axis with the largest stride (-1 if every extent is 1).
static void make_layout(int nax, const int64_t* ext, const int64_t* hs, Layout& L, int64_t* ds_axis, int* outer_axis) {
    int order[3], cnt = 0;
    for (int a = 0; a < nax; ++a) if (ext[a] > 1) order[cnt++] = a;
    std::stable_sort(order, order + cnt, [&](int a, int b) { return hs[a] > hs[b]; });
    L = Layout();
    int64_t total = 1;
    for (int a = 0; a < nax; ++a) total *= ext[a];
    L.total = total;
    int64_t ds = 1;
    for (int k = cnt - 1; k >= 0; --k) { ds_axis[order[k]] = ds; ds *= ext[order[k]]; }
    for (int a = 0; a < nax; ++a) if (ext[a] <= 1) ds_axis[a] = total;
    *outer_axis = cnt ? order[0] : -1;
    L.span = 1;
    for (int k = 0; k < cnt; ++k) L.span += (ext[order[k]] - 1) * hs[order[k]];
    for (int k = 0; k + 1 < cnt; ++k)
        if (hs[order[k]] < hs[order[k + 1]] * ext[order[k + 1]]) L.nested = false;
    if (cnt && hs[order[cnt - 1]] == 0) L.nested = false;
    // merge adjacent axes that are contiguous with each other
    for (int k = 0; k < cnt; ++k) {
        const int a = order[k];
        if (L.nd > 0 && L.hs[L.nd - 1] == hs[a] * ext[a]) { L.n[L.nd - 1] *= ext[a]; L.hs[L.nd - 1] = hs[a]; }
        else { L.n[L.nd] = ext[a]; L.hs[L.nd] = hs[a]; ++L.nd; }
    }
    if (L.nd > 0 && L.hs[L.nd - 1] == 1) { L.run = L.n[L.nd - 1]; L.nouter = L.nd - 1; }
    else { L.run = 1; L.nouter = L.nd; }
}

// copy the dense elements [e0, e1) between the host array (element 0 at `host`) and its dense image (element `ebase` at `dense`)
static void copy_elems(const Layout& L, char* host, char* dense, int64_t ebase, int64_t e0, int64_t e1, bool to_dense) {
    const int64_t w = L.run;
    int64_t r = e0 / w, within = e0 - r * w;
    int64_t idx[3] = {0, 0, 0}, hoff = 0;
    for (int k = L.nouter - 1; k >= 0; --k) { idx[k] = r % L.n[k]; r /= L.n[k]; hoff += idx[k] * L.hs[k]; }
    int64_t e = e0;
    while (e < e1) {
        const int64_t len = std::min(w - within, e1 - e);
        char* h = host + (hoff + within) * 8;
        char* d = dense + (e - ebase) * 8;
        if (len == 1) { if (to_dense) *reinterpret_cast<double*>(d) = *reinterpret_cast<const double*>(h); else *reinterpret_cast<double*>(h) = *reinterpret_cast<const double*>(d); }
        else if (to_dense) memcpy(d, h, (size_t)len * 8);
        else memcpy(h, d, (size_t)len * 8);
        e += len; within = 0;
        for (int k = L.nouter - 1; k >= 0; --k) {
            ++idx[k]; hoff += L.hs[k];
            if (idx[k] < L.n[k]) break;
            hoff -= L.hs[k] * L.n[k]; idx[k] = 0;
        }
    }
}

static bool is_pinned(const void* ptr, int64_t span_elems) {
    auto one = [](const void* q) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, q) != cudaSuccess) { cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeHost;
    };
    return one(ptr) && one(reinterpret_cast<const char*>(ptr) + (span_elems - 1) * 8);
}

// can the DMA engines take this layout straight from / to page-locked host memory?
static bool direct_ok(const Layout& L) {
    if (L.nd <= 1 && (L.nd == 0 || L.hs[0] == 1)) return true;                     // one contiguous block
    if (L.run * 8 < 32 || L.hs[L.nd - 1] != 1) return false;                       // element gathers: the host team does them
    if (L.nd == 3 && L.n[0] > 4096) return false;
    for (int k = 0; k < L.nouter; ++k) if (L.hs[k] * 8 > ((int64_t)1 << 30)) return false;   // pitch limit of 2-D copies
    return true;
}

static int direct_copy(const Layout& L, char* host, char* dev, bool h2d, cudaStream_t st) {
    const cudaMemcpyKind kind = h2d ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    if (L.nd <= 1) {
        if (h2d) SHB_CUDA(cudaMemcpyAsync(dev, host, (size_t)L.total * 8, kind, st));
        else SHB_CUDA(cudaMemcpyAsync(host, dev, (size_t)L.total * 8, kind, st));
        return 0;
    }
    const size_t width = (size_t)L.run * 8;
    const int64_t n0 = L.nd == 3 ? L.n[0] : 1, h0 = L.nd == 3 ? L.hs[0] : 0;
    const int64_t rows = L.n[L.nd - 2], hpitch = L.hs[L.nd - 2] * 8;
    for (int64_t i = 0; i < n0; ++i) {
        char* h = host + i * h0 * 8;
        char* d = dev + i * rows * (int64_t)width;
        if (h2d) SHB_CUDA(cudaMemcpy2DAsync(d, width, h, (size_t)hpitch, width, (size_t)rows, kind, st));
        else SHB_CUDA(cudaMemcpy2DAsync(h, (size_t)hpitch, d, width, width, (size_t)rows, kind, st));
    }
    return 0;
}

// ---- per-plan state of the host entry points -----------------------------------------------------------------------
struct HostIO {
    static constexpr int MAXCH = 8, SLOTS = 3;
    cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[MAXCH] = {}, ev_cmp[MAXCH] = {}, ev_syn[MAXCH] = {}, ev_d2h[MAXCH] = {}, ev_mid[MAXCH] = {};
    cudaEvent_t ev_slot[2][SLOTS] = {};
    char* bounce[2][SLOTS] = {};          // pinned; [0] = copy-in ring, [1] = copy-out ring
    size_t bounce_bytes = 0;
    double* stage[2] = {nullptr, nullptr};    // dense device images: [0] coefficients, [1] grid
    int64_t stage_elems[2] = {0, 0};
};

static size_t bounce_size() {
    static const size_t n = [] {
        const char* e = getenv("SHB200_BOUNCE_MIB");
        const int v = e ? atoi(e) : 8;      // 4 MiB pieces: 14.2 ms, 16 MiB: 16.6 ms, 64 MiB: 16.1 ms for the call above (8 threads)
        return (size_t)std::max(1, std::min(v, 1024)) << 20;
    }();
    return n;
}

static int ensure_hio(Plan& p) {
    if (p.hio) return 0;
    HostIO* h = new HostIO();
    p.hio = h;
    SHB_CUDA(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    SHB_CUDA(cudaStreamCreateWithFlags(&h->s_cmp, cudaStreamNonBlocking));
    SHB_CUDA(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < HostIO::MAXCH; ++i) {
        SHB_CUDA(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
        SHB_CUDA(cudaEventCreateWithFlags(&h->ev_cmp[i], cudaEventDisableTiming));
        SHB_CUDA(cudaEventCreateWithFlags(&h->ev_syn[i], cudaEventDisableTiming));
        SHB_CUDA(cudaEventCreateWithFlags(&h->ev_d2h[i], cudaEventDisableTiming));
        SHB_CUDA(cudaEventCreateWithFlags(&h->ev_mid[i], cudaEventDisableTiming));
    }
    for (int d = 0; d < 2; ++d)
        for (int s = 0; s < HostIO::SLOTS; ++s) SHB_CUDA(cudaEventCreateWithFlags(&h->ev_slot[d][s], cudaEventDisableTiming));
    return 0;
}

static int ensure_bounce(HostIO& h, int dir) {
    if (h.bounce[dir][0]) return 0;
    h.bounce_bytes = bounce_size();
    for (int s = 0; s < HostIO::SLOTS; ++s) SHB_CUDA(cudaHostAlloc((void**)&h.bounce[dir][s], h.bounce_bytes, cudaHostAllocPortable));
    return 0;
}

static int ensure_stage(HostIO& h, int which, int64_t need) {
    if (h.stage_elems[which] >= need) return 0;
    if (h.stage[which]) { SHB_CUDA(cudaDeviceSynchronize()); cudaFree(h.stage[which]); }
    h.stage[which] = nullptr; h.stage_elems[which] = 0;
    SHB_CUDA(cudaMalloc((void**)&h.stage[which], sizeof(double) * (size_t)need));
    h.stage_elems[which] = need;
    return 0;
}

void free_host_io(Plan& p) {
    HostIO* h = p.hio;
    if (!h) return;
    for (int d = 0; d < 2; ++d)
        for (int s = 0; s < HostIO::SLOTS; ++s) {
            if (h->bounce[d][s]) cudaFreeHost(h->bounce[d][s]);
            if (h->ev_slot[d][s]) cudaEventDestroy(h->ev_slot[d][s]);
        }
    for (int i = 0; i < HostIO::MAXCH; ++i) {
        if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
        if (h->ev_cmp[i]) cudaEventDestroy(h->ev_cmp[i]);
        if (h->ev_syn[i]) cudaEventDestroy(h->ev_syn[i]);
        if (h->ev_d2h[i]) cudaEventDestroy(h->ev_d2h[i]);
        if (h->ev_mid[i]) cudaEventDestroy(h->ev_mid[i]);
    }
    for (int w = 0; w < 2; ++w) if (h->stage[w]) cudaFree(h->stage[w]);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_cmp) cudaStreamDestroy(h->s_cmp);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    delete h;
    p.hio = nullptr;
}

// One host array of a call: its logical axes (axis 0 = batch), the pinned / direct decision and the dense device image.
struct HostArray {
    int nax = 0;
    int64_t ext[3] = {1, 1, 1}, hs[3] = {0, 0, 0};
    char* host = nullptr;
    double* dev = nullptr;            // dense image of the WHOLE call
    bool batch_outer = false;         // the batch axis has the largest stride (chunks are then dense sub-blocks)
    bool pinned = false;
    int64_t ds_b = 0;                 // dense stride of the batch axis in the whole-call image
    Layout full;
};

static void describe(HostArray& a) {
    int64_t ds[3]; int outer;
    make_layout(a.nax, a.ext, a.hs, a.full, ds, &outer);
    a.batch_outer = outer == 0 || outer == -1 || a.ext[0] == 1;
    a.ds_b = 1;
    for (int k = 1; k < a.nax; ++k) a.ds_b *= a.ext[k];
    a.pinned = is_pinned(a.host, a.full.span);
}

// layout and dense strides of the fields [b0, b0 + nb) of an array (whole call when nb == ext[0])
struct ChunkView {
    Layout L;
    int64_t ds[3];
    char* host;
    double* dev;
};
static ChunkView chunk_view(const HostArray& a, int b0, int nb) {
    ChunkView v;
    int64_t ext[3] = {nb, a.ext[1], a.ext[2]};
    int outer;
    make_layout(a.nax, ext, a.hs, v.L, v.ds, &outer);
    v.host = a.host + (int64_t)b0 * a.hs[0] * 8;
    v.dev = a.dev + (int64_t)b0 * a.ds_b;
    return v;
}

// copy-in of one chunk on stream s_in: direct DMA or gather -> bounce -> H2D, piece by piece
static int copy_in(HostIO& h, const HostArray& a, const ChunkView& v, int& slot_ctr) {
    if (a.pinned && direct_ok(v.L)) return direct_copy(v.L, v.host, reinterpret_cast<char*>(v.dev), true, h.s_in);
    int rc = ensure_bounce(h, 0);
    if (rc) return rc;
    const int64_t piece = (int64_t)(h.bounce_bytes / 8);
    Team* tm = team(0);
    for (int64_t e0 = 0; e0 < v.L.total; e0 += piece, ++slot_ctr) {
        const int64_t e1 = std::min(v.L.total, e0 + piece);
        const int s = slot_ctr % HostIO::SLOTS;
        if (slot_ctr >= HostIO::SLOTS) SHB_CUDA(cudaEventSynchronize(h.ev_slot[0][s]));      // the slot's previous H2D has left
        char* slot = h.bounce[0][s];
        const Layout& L = v.L;
        char* host = v.host;
        tm->run(e1 - e0, [&](int64_t lo, int64_t hi) { copy_elems(L, host, slot, e0, e0 + lo, e0 + hi, true); });
        SHB_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(v.dev) + e0 * 8, slot, (size_t)(e1 - e0) * 8, cudaMemcpyHostToDevice, h.s_in));
        SHB_CUDA(cudaEventRecord(h.ev_slot[0][s], h.s_in));
    }
    return 0;
}

// copy-out of one chunk through the bounce ring (runs on the copy-out pump thread); s_out already waits for the chunk's kernels
struct OutRing {
    struct Pending { int64_t e0, e1; const ChunkView* v; };
    Pending pend[HostIO::SLOTS];
    bool used[HostIO::SLOTS] = {false, false, false};
    int ctr = 0;
};
static int drain_slot(HostIO& h, OutRing& r, int s) {
    if (!r.used[s]) return 0;
    SHB_CUDA(cudaEventSynchronize(h.ev_slot[1][s]));
    const OutRing::Pending& q = r.pend[s];
    char* slot = h.bounce[1][s];
    const Layout& L = q.v->L;
    char* host = q.v->host;
    const int64_t e0 = q.e0;
    team(1)->run(q.e1 - q.e0, [&](int64_t lo, int64_t hi) { copy_elems(L, host, slot, e0, e0 + lo, e0 + hi, false); });
    r.used[s] = false;
    return 0;
}
static int copy_out_bounced(HostIO& h, const ChunkView& v, OutRing& r) {
    const int64_t piece = (int64_t)(h.bounce_bytes / 8);
    for (int64_t e0 = 0; e0 < v.L.total; e0 += piece, ++r.ctr) {
        const int64_t e1 = std::min(v.L.total, e0 + piece);
        const int s = r.ctr % HostIO::SLOTS;
        int rc = drain_slot(h, r, s);
        if (rc) return rc;
        SHB_CUDA(cudaMemcpyAsync(h.bounce[1][s], reinterpret_cast<const char*>(v.dev) + e0 * 8, (size_t)(e1 - e0) * 8, cudaMemcpyDeviceToHost, h.s_out));
        SHB_CUDA(cudaEventRecord(h.ev_slot[1][s], h.s_out));
        r.pend[s] = {e0, e1, &v};
        r.used[s] = true;
    }
    return 0;
}

// chunks of at least 16 fields: below 32 columns the Legendre stage leaves the tensor-core kernels
static int host_chunks(int batch) { return std::max(1, std::min(4, batch / 16)); }

// Chunk sizes of one host call: up to 4 equal chunks of whole 8-field groups; the chunk at the exposed end of the pipeline --
// the FIRST of a synthesis (nothing travels back before its grid exists), the LAST of an analysis (its coefficients travel back
// after everything else), both for the round trip -- is cut into 8 + 8 + rest fields, which shortens the un-overlapped head /
// tail of the call (SHB200_HOST_EDGE_SPLIT=0: equal chunks only).  which = 0 synthesis, 1 analysis, 2 round trip; chunkable:
// the batch axis is the slowest axis of every array of the call.  At most 4 + 2 + 2 = HostIO::MAXCH chunks.
static std::vector<int> host_chunk_plan(int batch, int which, bool chunkable) {
    std::vector<int> nbs;
    int nchunk = chunkable ? host_chunks(batch) : 1;
    const int per = nchunk == 1 ? batch : ((batch + nchunk - 1) / nchunk + 7) / 8 * 8;
    for (int b0 = 0; b0 < batch; b0 += per) nbs.push_back(std::min(per, batch - b0));
    static const bool edge_split = [] { const char* e = getenv("SHB200_HOST_EDGE_SPLIT"); return !e || atoi(e) != 0; }();
    auto split = [&](size_t at, bool small_first) {
        const int nb = nbs[at];
        if (nb <= 8) return;
        std::vector<int> parts;
        if (small_first) {
            if (nb >= 24) parts = {8, 8, nb - 16}; else parts = {8, nb - 8};
        } else {                                         // chunk boundaries stay multiples of 8 fields: the tail takes the remainder
            const int last = nb % 8 ? nb % 8 : 8, rest = nb - last;
            if (rest >= 16) parts = {rest - 8, 8, last}; else parts = {rest, last};
        }
        nbs.erase(nbs.begin() + at);
        nbs.insert(nbs.begin() + at, parts.begin(), parts.end());
    };
    if (edge_split && nbs.size() > 1) {
        if (which == 1 || which == 2) split(nbs.size() - 1, false);
        if (which == 0 || which == 2) split(0, true);
    }
    return nbs;
}

// The pipeline.  which = 0: synthesis (in = coefficients, out = grid); 1: analysis (in = grid, out = coefficients);
// 2: round trip (in = coefficients, mid = grid written to the host AND read back from it, out = coefficients).
struct HostCall {
    Plan* p;
    int which, batch;
    HostArray in, mid, out;           // mid only for the round trip
    CallOpts o_syn, o_ana;
};

static int launch_chunk(HostCall& c, int stage, const ChunkView& src, const ChunkView& dst, int nb, cudaStream_t st) {
    Plan& p = *c.p;
    if (stage == 0)      // coefficients (b, nm) -> grid (b, lat, lon)
        return do_synthesis(p, nb, src.dev, src.ds[0], src.ds[1], dst.dev, dst.ds[0], dst.ds[1], p.mode == SHB200_MODE_POINTS ? 0 : dst.ds[2], st, c.o_syn);
    return do_analysis(p, nb, src.dev, src.ds[0], src.ds[1], src.ds[2], dst.dev, dst.ds[0], dst.ds[1], st, c.o_ana);
}

static int run_host_call(HostCall& c) {
    Plan& p = *c.p;
    SHB_CUDA(cudaSetDevice(p.device));
    int rc = ensure_hio(p);
    if (rc) return rc;
    HostIO& h = *p.hio;
    const bool rt = c.which == 2;
    describe(c.in);
    describe(c.out);
    if (rt) describe(c.mid);
    if (!c.out.full.nested || (rt && !c.mid.full.nested)) { set_error("host entry points: the output array overlaps itself (zero or aliasing strides)"); return SHB200_EINVAL; }
    // dense device images
    const int in_stage = (c.which == 1) ? 1 : 0, out_stage = (c.which == 0) ? 1 : 0;
    if (rt) {
        // coefficients in and out share the coefficient staging array (in: first half, out: second half)
        if ((rc = ensure_stage(h, 0, c.in.full.total + c.out.full.total))) return rc;
        if ((rc = ensure_stage(h, 1, c.mid.full.total))) return rc;
        c.in.dev = h.stage[0]; c.out.dev = h.stage[0] + c.in.full.total; c.mid.dev = h.stage[1];
    } else {
        if ((rc = ensure_stage(h, in_stage, c.in.full.total))) return rc;
        if ((rc = ensure_stage(h, out_stage, c.out.full.total))) return rc;
        c.in.dev = h.stage[in_stage]; c.out.dev = h.stage[out_stage];
    }
    std::vector<int> nbs = host_chunk_plan(c.batch, c.which, c.in.batch_outer && c.out.batch_outer && (!rt || c.mid.batch_outer));
    const int nchunk = (int)nbs.size();        // <= 4 + 2 + 2 = HostIO::MAXCH
    // an 8-field chunk of a tensor-core plan runs 32 columns wide (half of them zero): it stays on the kernels of its neighbours
    if (p.dmma && p.Bp >= 16) c.o_syn.min_fields = c.o_ana.min_fields = 16;
    std::vector<int> b0s(nchunk, 0);
    for (int k = 1; k < nchunk; ++k) b0s[k] = b0s[k - 1] + nbs[k - 1];

    std::vector<ChunkView> vin(nchunk), vmid(rt ? nchunk : 0), vout(nchunk);
    for (int k = 0; k < nchunk; ++k) {
        vin[k] = chunk_view(c.in, b0s[k], nbs[k]);
        vout[k] = chunk_view(c.out, b0s[k], nbs[k]);
        if (rt) vmid[k] = chunk_view(c.mid, b0s[k], nbs[k]);
    }
    auto is_direct = [&](const HostArray& a) {
        for (int k = 0; k < nchunk; ++k) if (!(a.pinned && direct_ok(chunk_view(a, b0s[k], nbs[k]).L))) return false;
        return true;
    };
    const bool out_direct = is_direct(c.out);
    const bool mid_direct = rt ? is_direct(c.mid) : true;
    if (rt && !mid_direct) {
        // the grid of a round trip is written and read back chunk by chunk: that ordering is expressed with stream events,
        // which needs DMA-able (page-locked) memory.  Pageable grids take two separate calls.
        set_error("roundtrip_host: the grid array must be page-locked (shb200_host_alloc / cudaHostRegister)");
        return SHB200_EUNSUPPORTED;
    }
    if (!out_direct && (rc = ensure_bounce(h, 1))) return rc;

    // copy-out pump for a bounced output: a helper thread that follows the chunks as their kernels are launched
    std::mutex mu;
    std::condition_variable cv;
    int launched = 0;
    bool aborted = false;
    std::atomic<int> out_rc{0};
    std::thread pump;
    if (!out_direct) {
        pump = std::thread([&] {
            cudaSetDevice(p.device);
            OutRing ring;
            for (int k = 0; k < nchunk; ++k) {
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return launched > k || aborted; });
                    if (aborted) break;
                }
                if (cudaStreamWaitEvent(h.s_out, h.ev_cmp[k], 0) != cudaSuccess) { out_rc = SHB200_ECUDA; break; }
                const int r = copy_out_bounced(h, vout[k], ring);
                if (r) { out_rc = r; break; }
            }
            for (int s = 0; s < HostIO::SLOTS; ++s) {
                const int r = drain_slot(h, ring, (ring.ctr + s) % HostIO::SLOTS);
                if (r && !out_rc) out_rc = r;
            }
        });
    }
    auto finish = [&](int code) {
        if (pump.joinable()) {
            if (code) { std::lock_guard<std::mutex> lk(mu); aborted = true; }
            cv.notify_all();
            pump.join();
        }
        cudaStreamSynchronize(h.s_in);
        cudaStreamSynchronize(h.s_cmp);
        cudaError_t e = cudaStreamSynchronize(h.s_out);
        if (code) return code;
        if (out_rc) return out_rc.load();
        if (e != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
        return 0;
    };

    auto ck = [](cudaError_t e) { return e == cudaSuccess ? 0 : cuda_fail(e, "stream/event call", __FILE__, __LINE__); };
    // copy-out of a finished chunk: direct DMA on s_out, or hand it to the pump
    auto emit = [&](int k) -> int {
        int r;
        if ((r = ck(cudaEventRecord(h.ev_cmp[k], h.s_cmp)))) return r;
        if (out_direct) {
            if ((r = ck(cudaStreamWaitEvent(h.s_out, h.ev_cmp[k], 0)))) return r;
            return direct_copy(vout[k].L, vout[k].host, reinterpret_cast<char*>(vout[k].dev), false, h.s_out);
        }
        { std::lock_guard<std::mutex> lk(mu); launched = k + 1; }
        cv.notify_all();
        return 0;
    };
    int slot_ctr = 0;
    if (!rt) {
        for (int k = 0; k < nchunk; ++k) {
            if ((rc = copy_in(h, c.in, vin[k], slot_ctr))) return finish(rc);
            if ((rc = ck(cudaEventRecord(h.ev_in[k], h.s_in))) || (rc = ck(cudaStreamWaitEvent(h.s_cmp, h.ev_in[k], 0)))) return finish(rc);
            if ((rc = launch_chunk(c, c.which, vin[k], vout[k], nbs[k], h.s_cmp))) return finish(rc);
            if ((rc = emit(k))) return finish(rc);
        }
        return finish(0);
    }
    // Round trip, software-pipelined over the chunks so that the grid of chunk k+1 travels to the host while the grid of
    // chunk k travels back:
    //   s_in  : coef 0 .. n-1 | grid 0 (after its copy-out) | grid 1 | ...
    //   s_cmp : syn 0 | syn 1 | ana 0 | syn 2 | ana 1 | ...
    //   s_out : grid 0 | grid 1 | coef' 0 | grid 2 | coef' 1 | ...
    // The analysis reads the grid back from the HOST array -- the bytes a caller who makes two separate calls would upload.
    for (int k = 0; k < nchunk; ++k) {
        if ((rc = copy_in(h, c.in, vin[k], slot_ctr))) return finish(rc);
        if ((rc = ck(cudaEventRecord(h.ev_in[k], h.s_in)))) return finish(rc);
    }
    for (int k = 0; k <= nchunk; ++k) {
        if (k < nchunk) {
            char* mdev = reinterpret_cast<char*>(vmid[k].dev);
            if ((rc = ck(cudaStreamWaitEvent(h.s_cmp, h.ev_in[k], 0)))) return finish(rc);
            if ((rc = launch_chunk(c, 0, vin[k], vmid[k], nbs[k], h.s_cmp))) return finish(rc);
            if ((rc = ck(cudaEventRecord(h.ev_syn[k], h.s_cmp))) || (rc = ck(cudaStreamWaitEvent(h.s_out, h.ev_syn[k], 0)))) return finish(rc);
            if ((rc = direct_copy(vmid[k].L, vmid[k].host, mdev, false, h.s_out))) return finish(rc);
            if ((rc = ck(cudaEventRecord(h.ev_d2h[k], h.s_out))) || (rc = ck(cudaStreamWaitEvent(h.s_in, h.ev_d2h[k], 0)))) return finish(rc);
            if ((rc = direct_copy(vmid[k].L, vmid[k].host, mdev, true, h.s_in))) return finish(rc);
            if ((rc = ck(cudaEventRecord(h.ev_mid[k], h.s_in)))) return finish(rc);
        }
        if (k >= 1) {
            if ((rc = ck(cudaStreamWaitEvent(h.s_cmp, h.ev_mid[k - 1], 0)))) return finish(rc);
            if ((rc = launch_chunk(c, 1, vmid[k - 1], vout[k - 1], nbs[k - 1], h.s_cmp))) return finish(rc);
            if ((rc = emit(k - 1))) return finish(rc);
        }
    }
    return finish(0);
}

static bool bad_strides(std::initializer_list<int64_t> s) {
    for (int64_t v : s) if (v < 0) return true;
    return false;
}

static void grid_array(const Plan& p, HostArray& a, int batch, const double* ptr, int64_t gs_b, int64_t gs_lat, int64_t gs_lon) {
    a.host = reinterpret_cast<char*>(const_cast<double*>(ptr));
    a.ext[0] = batch; a.hs[0] = gs_b;
    a.ext[1] = p.nlat; a.hs[1] = gs_lat;
    if (p.mode == SHB200_MODE_POINTS) { a.nax = 2; }
    else { a.nax = 3; a.ext[2] = p.nlon; a.hs[2] = gs_lon; }
}
static void coef_array(HostArray& a, int batch, int64_t nsh, const double* ptr, int64_t cs_b, int64_t cs_nm) {
    a.host = reinterpret_cast<char*>(const_cast<double*>(ptr));
    a.nax = 2;
    a.ext[0] = batch; a.hs[0] = cs_b;
    a.ext[1] = nsh; a.hs[1] = cs_nm;
}

// ---- page-locked host memory pool ----------------------------------------------------------------------------------
struct PinnedPool {
    std::mutex mu;
    std::multimap<size_t, void*> free_blocks;
    std::unordered_map<void*, size_t> size_of;
    size_t cached = 0;
    size_t cap() {
        static const size_t c = [] {
            const char* e = getenv("SHB200_PINNED_CACHE_MIB");
            const long v = e ? atol(e) : 8192;
            return (size_t)std::max(0L, v) << 20;
        }();
        return c;
    }
};
static PinnedPool& pool() { static PinnedPool* p = new PinnedPool(); return *p; }

}  // namespace shb

using namespace shb;

extern "C" {

int shb200_synthesis_host(shb200_plan* plan, int32_t batch, const double* coef, int64_t cs_b, int64_t cs_nm,
                          double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon, const shb200_call_opts* opts) {
    if (!plan || !coef || !grid) { set_error("synthesis_host: null argument"); return SHB200_EINVAL; }
    Plan& p = plan->p;
    if (batch < 1 || batch > p.maxB) { set_error("batch must be in [1, max_batch] of the plan"); return SHB200_EINVAL; }
    if (bad_strides({cs_b, cs_nm, gs_b, gs_lat, gs_lon})) { set_error("host entry points need non-negative strides"); return SHB200_EINVAL; }
    HostCall c;
    c.p = &p; c.which = 0; c.batch = batch;
    int rc = parse_opts(p, opts, c.o_syn);
    if (rc) return rc;
    coef_array(c.in, batch, p.nsh, coef, cs_b, cs_nm);
    grid_array(p, c.out, batch, grid, gs_b, gs_lat, gs_lon);
    std::lock_guard<std::mutex> lock(p.mu);
    return run_host_call(c);
}

int shb200_analysis_host(shb200_plan* plan, int32_t batch, const double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon,
                         double* coef, int64_t cs_b, int64_t cs_nm, const shb200_call_opts* opts) {
    if (!plan || !coef || !grid) { set_error("analysis_host: null argument"); return SHB200_EINVAL; }
    Plan& p = plan->p;
    if (p.quadrature == SHB200_QUAD_NONE) { set_error("analysis: plan was created without quadrature weights"); return SHB200_EINVAL; }
    if (batch < 1 || batch > p.maxB) { set_error("batch must be in [1, max_batch] of the plan"); return SHB200_EINVAL; }
    if (bad_strides({cs_b, cs_nm, gs_b, gs_lat, gs_lon})) { set_error("host entry points need non-negative strides"); return SHB200_EINVAL; }
    HostCall c;
    c.p = &p; c.which = 1; c.batch = batch;
    int rc = parse_opts(p, opts, c.o_ana);
    if (rc) return rc;
    grid_array(p, c.in, batch, grid, gs_b, gs_lat, gs_lon);
    coef_array(c.out, batch, (int64_t)(p.nmax + 1) * (p.nmax + 1), coef, cs_b, cs_nm);
    std::lock_guard<std::mutex> lock(p.mu);
    return run_host_call(c);
}

int shb200_roundtrip_host(shb200_plan* plan, int32_t batch, const double* coef, int64_t cs_b, int64_t cs_nm,
                          double* grid, int64_t gs_b, int64_t gs_lat, int64_t gs_lon,
                          double* coef_out, int64_t os_b, int64_t os_nm,
                          const shb200_call_opts* syn_opts, const shb200_call_opts* ana_opts) {
    if (!plan || !coef || !grid || !coef_out) { set_error("roundtrip_host: null argument"); return SHB200_EINVAL; }
    Plan& p = plan->p;
    if (p.quadrature == SHB200_QUAD_NONE || p.mode != SHB200_MODE_GRID) { set_error("roundtrip_host: needs a grid plan with quadrature weights"); return SHB200_EINVAL; }
    if (batch < 1 || batch > p.maxB) { set_error("batch must be in [1, max_batch] of the plan"); return SHB200_EINVAL; }
    if (bad_strides({cs_b, cs_nm, gs_b, gs_lat, gs_lon, os_b, os_nm})) { set_error("host entry points need non-negative strides"); return SHB200_EINVAL; }
    HostCall c;
    c.p = &p; c.which = 2; c.batch = batch;
    int rc = parse_opts(p, syn_opts, c.o_syn);
    if (rc) return rc;
    if ((rc = parse_opts(p, ana_opts, c.o_ana))) return rc;
    coef_array(c.in, batch, p.nsh, coef, cs_b, cs_nm);
    grid_array(p, c.mid, batch, grid, gs_b, gs_lat, gs_lon);
    coef_array(c.out, batch, (int64_t)(p.nmax + 1) * (p.nmax + 1), coef_out, os_b, os_nm);
    std::lock_guard<std::mutex> lock(p.mu);
    return run_host_call(c);
}

int shb200_debug_host_chunks(int32_t batch, int32_t which, int32_t chunkable, int32_t* sizes) {
    if (batch < 1 || which < 0 || which > 2 || !sizes) { set_error("debug_host_chunks: invalid argument"); return SHB200_EINVAL; }
    const std::vector<int> nbs = host_chunk_plan(batch, which, chunkable != 0);
    for (size_t i = 0; i < nbs.size() && i < (size_t)HostIO::MAXCH; ++i) sizes[i] = nbs[i];
    return (int)nbs.size();
}

int shb200_debug_host_copy(int32_t nax, const int64_t* ext, const int64_t* strides, double* host, double* dense, int32_t to_dense, int64_t* info) {
    if (nax < 1 || nax > 3 || !ext || !strides || !host || !dense) { set_error("debug_host_copy: invalid argument"); return SHB200_EINVAL; }
    Layout L;
    int64_t ds[3]; int outer;
    make_layout(nax, ext, strides, L, ds, &outer);
    if (info) {
        info[0] = L.nd; info[1] = L.run; info[2] = L.total; info[3] = L.nested ? 1 : 0; info[4] = direct_ok(L) ? 1 : 0;
        info[5] = L.span; info[6] = outer;
        for (int a = 0; a < nax; ++a) info[7 + a] = ds[a];
    }
    if (!to_dense && !L.nested) { set_error("debug_host_copy: scatter into an overlapping array"); return SHB200_EINVAL; }
    // in pieces, through the thread team, exactly like the bounce path
    const int64_t piece = 1000;
    for (int64_t e0 = 0; e0 < L.total; e0 += piece) {
        const int64_t e1 = std::min(L.total, e0 + piece);
        char* h = reinterpret_cast<char*>(host);
        char* d = reinterpret_cast<char*>(dense) + e0 * 8;
        team(to_dense ? 0 : 1)->run(e1 - e0, [&](int64_t lo, int64_t hi) { copy_elems(L, h, d, e0, e0 + lo, e0 + hi, to_dense != 0); });
    }
    return 0;
}

int shb200_host_alloc(int64_t bytes, void** out) {
    if (bytes < 0 || !out) { set_error("host_alloc: invalid argument"); return SHB200_EINVAL; }
    const size_t need = ((size_t)std::max<int64_t>(bytes, 1) + ((size_t)2 << 20) - 1) >> 21 << 21;      // 2 MiB granules
    PinnedPool& P = pool();
    {
        std::lock_guard<std::mutex> lk(P.mu);
        auto it = P.free_blocks.lower_bound(need);
        if (it != P.free_blocks.end() && it->first <= need + need / 4) {
            *out = it->second;
            P.cached -= it->first;
            P.free_blocks.erase(it);
            return 0;
        }
    }
    void* ptr = nullptr;
    cudaError_t e = cudaHostAlloc(&ptr, need, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        // make room: drop the cached blocks and try once more
        shb200_host_trim();
        cudaGetLastError();
        e = cudaHostAlloc(&ptr, need, cudaHostAllocPortable);
    }
    if (e != cudaSuccess) { *out = nullptr; return cuda_fail(e, "cudaHostAlloc", __FILE__, __LINE__); }
    {
        std::lock_guard<std::mutex> lk(P.mu);
        P.size_of[ptr] = need;
    }
    *out = ptr;
    return 0;
}

int shb200_host_free(void* ptr) {
    if (!ptr) return 0;
    PinnedPool& P = pool();
    std::vector<void*> drop;
    {
        std::lock_guard<std::mutex> lk(P.mu);
        auto it = P.size_of.find(ptr);
        if (it == P.size_of.end()) { set_error("host_free: pointer was not allocated by shb200_host_alloc"); return SHB200_EINVAL; }
        P.free_blocks.emplace(it->second, ptr);
        P.cached += it->second;
        // over the cap: release the largest cached blocks first
        while (P.cached > P.cap() && !P.free_blocks.empty()) {
            auto big = std::prev(P.free_blocks.end());
            P.cached -= big->first;
            P.size_of.erase(big->second);
            drop.push_back(big->second);
            P.free_blocks.erase(big);
        }
    }
    for (void* d : drop) cudaFreeHost(d);
    return 0;
}

int shb200_host_trim(void) {
    PinnedPool& P = pool();
    std::vector<void*> drop;
    {
        std::lock_guard<std::mutex> lk(P.mu);
        for (auto& kv : P.free_blocks) { drop.push_back(kv.second); P.size_of.erase(kv.second); }
        P.free_blocks.clear();
        P.cached = 0;
    }
    for (void* d : drop) cudaFreeHost(d);
    return 0;
}

}  // extern "C"
