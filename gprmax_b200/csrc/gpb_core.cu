// gpb_core.cu -- host side of libgprmax_b200.so: the C ABI declared in include/gprmax_b200.h.
//
// Replaces the reference's solve_gpu driver (gprMax/model_build_run.py:477-716) and the PyCUDA
// helpers it calls (grid.py:247-272, pml.py:298-364, receivers.py:45-88, sources.py:235-283,
// snapshots.py:170-228, utilities.py:341-413).  No CPU fallback: every entry point that computes
// needs a CUDA device and reports an error otherwise.
#include "../../include/gprmax_b200.h"
#include "gpb_kernels.cuh"
#include "gpb_kernels_v4.cuh"
#include "gpb_kernels_coop.cuh"
#include "gpb_tma.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <pthread.h>
#include <sched.h>
#include <unistd.h>
#include <vector>

using namespace gpb;

static thread_local std::string g_err;

static int fail(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// ----------------------------------------------------------------------------------------------
// Device-memory cache.  A B-scan (-n traces) creates and destroys one solver per trace with identical array sizes, and
// cudaMalloc / cudaFree of the multi-GB field and ID arrays cost more than 100 ms per trace (profiles/e2e_breakdown.py).
// Freed blocks are kept per device and handed back on an exact size match; everything is released by
// gpb_release_cached(), when an allocation fails, or when GPB_NO_POOL is set (then every block goes straight to cudaFree).
namespace {
struct DevicePool {
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, void *> free_blocks;   // (device, bytes) -> block
    std::map<void *, std::pair<int, size_t>> live;               // block -> (device, bytes)
    size_t cached_bytes = 0;
    std::vector<std::pair<size_t, void *>> pinned_free;           // page-locked host buffers (ID upload bounce buffers)

    bool enabled() const { return !getenv("GPB_NO_POOL"); }
    void release_all()
    {
        std::lock_guard<std::mutex> g(mu);
        int cur = 0;
        cudaGetDevice(&cur);
        for (auto &kv : free_blocks) {
            cudaSetDevice(kv.first.first);
            cudaFree(kv.second);
        }
        free_blocks.clear();
        cached_bytes = 0;
        for (auto &kv : pinned_free) cudaFreeHost(kv.second);
        pinned_free.clear();
        cudaSetDevice(cur);
    }
    cudaError_t alloc_pinned(void **out, size_t bytes)
    {
        {
            std::lock_guard<std::mutex> g(mu);
            for (size_t n = 0; n < pinned_free.size(); ++n)
                if (pinned_free[n].first == bytes) {
                    *out = pinned_free[n].second;
                    pinned_free.erase(pinned_free.begin() + n);
                    pinned_sizes[*out] = bytes;
                    return cudaSuccess;
                }
        }
        cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
        if (e == cudaSuccess) {
            std::lock_guard<std::mutex> g(mu);
            pinned_sizes[*out] = bytes;
        }
        return e;
    }
    void free_pinned(void *p)
    {
        if (!p) return;
        std::lock_guard<std::mutex> g(mu);
        auto it = pinned_sizes.find(p);
        const size_t bytes = it == pinned_sizes.end() ? 0 : it->second;
        if (it != pinned_sizes.end()) pinned_sizes.erase(it);
        if (enabled() && bytes) pinned_free.push_back({bytes, p});
        else cudaFreeHost(p);
    }
    std::map<void *, size_t> pinned_sizes;
    cudaError_t alloc(void **out, size_t bytes, int device)
    {
        {
            std::lock_guard<std::mutex> g(mu);
            auto it = free_blocks.find({device, bytes});
            if (it != free_blocks.end()) {
                *out = it->second;
                free_blocks.erase(it);
                cached_bytes -= bytes;
                live[*out] = {device, bytes};
                return cudaSuccess;
            }
        }
        cudaError_t e = cudaMalloc(out, bytes);
        if (e == cudaErrorMemoryAllocation) {   // make room and try once more
            cudaGetLastError();
            release_all();
            e = cudaMalloc(out, bytes);
        }
        if (e == cudaSuccess) {
            std::lock_guard<std::mutex> g(mu);
            live[*out] = {device, bytes};
        }
        return e;
    }
    void free(void *p)
    {
        if (!p) return;
        std::pair<int, size_t> info;
        {
            std::lock_guard<std::mutex> g(mu);
            auto it = live.find(p);
            if (it == live.end()) { cudaFree(p); return; }
            info = it->second;
            live.erase(it);
            if (enabled()) {
                free_blocks.insert({info, p});
                cached_bytes += info.second;
                return;
            }
        }
        cudaFree(p);
    }
};
DevicePool g_pool;

// The reference sets OMP_PROC_BIND=TRUE for every model (input_cmds_singleuse.py:78-80), so by the time solve_gpu is called
// the OpenMP runtime has bound the calling thread to ONE core.  Threads created from it inherit that mask: the CUDA driver's
// helper threads (created when the context is initialised) and this library's upload threads would all share the caller's
// core.  Entry points that may initialise CUDA or start threads therefore widen the calling thread's affinity to every CPU of
// the process's cpuset for their duration and restore it on return.
struct ScopedFullAffinity {
    cpu_set_t saved;
    bool have = false;
    ScopedFullAffinity()
    {
        if (sched_getaffinity(0, sizeof saved, &saved) != 0) return;
        cpu_set_t all;
        CPU_ZERO(&all);
        for (int c = 0; c < CPU_SETSIZE; ++c) CPU_SET(c, &all);
        have = sched_setaffinity(0, sizeof all, &all) == 0;
    }
    ~ScopedFullAffinity()
    {
        if (have) sched_setaffinity(0, sizeof saved, &saved);
    }
};

constexpr size_t kBounceBytes = 32u << 20;   // pinned bounce buffer / device stage of the ID upload

// host -> pinned copy split over a few threads (one thread saturates at ~10 GB/s)
void parallel_copy(char *dst, const char *src, size_t bytes, int nthreads)
{
    if (nthreads <= 1 || bytes < (4u << 20)) {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> th;
    const size_t part = (bytes / nthreads + 4095) & ~(size_t)4095;
    for (int t = 0; t < nthreads; ++t) {
        const size_t o = (size_t)t * part;
        if (o >= bytes) break;
        th.emplace_back([=] {
            // the caller's thread is often pinned to ONE core (the reference sets OMP_PROC_BIND=TRUE for every model,
            // input_cmds_singleuse.py:78-80, and the OpenMP runtime then binds the main thread): helper threads inherit that
            // mask and would all share the core (upload 22 -> 85 ms at 300^3).  Ask for every CPU; the kernel keeps the
            // process's cpuset.
            cpu_set_t all;
            CPU_ZERO(&all);
            for (int c = 0; c < CPU_SETSIZE; ++c) CPU_SET(c, &all);
            pthread_setaffinity_np(pthread_self(), sizeof all, &all);
            memcpy(dst + o, src + o, std::min(part, bytes - o));
        });
    }
    for (auto &t : th) t.join();
}
// uint32 -> IDT on the host, split over threads like parallel_copy; returns the largest value seen.  Narrowing before the
// transfer quarters the PCIe bytes and the pinned-buffer writes: the host side of the upload is memory-bound (a 12.9 GB slab
// per rank of a sharded run went through at 13 GB/s per rank with four ranks on one host)
template <typename IDT>
uint32_t parallel_narrow(IDT *dst, const uint32_t *src, size_t n, int nthreads)
{
    auto body = [](IDT *d, const uint32_t *s, size_t cnt) {
        uint32_t mx = 0;
        for (size_t q = 0; q < cnt; ++q) {
            const uint32_t v = s[q];
            mx = v > mx ? v : mx;
            d[q] = (IDT)v;
        }
        return mx;
    };
    if (nthreads <= 1 || n < (1u << 20)) return body(dst, src, n);
    std::vector<std::thread> th;
    std::vector<uint32_t> mx(nthreads, 0);
    const size_t part = (n / nthreads + 4095) & ~(size_t)4095;
    for (int t = 0; t < nthreads; ++t) {
        const size_t o = (size_t)t * part;
        if (o >= n) break;
        th.emplace_back([=, &mx] {
            cpu_set_t all;   // see parallel_copy
            CPU_ZERO(&all);
            for (int c = 0; c < CPU_SETSIZE; ++c) CPU_SET(c, &all);
            pthread_setaffinity_np(pthread_self(), sizeof all, &all);
            mx[t] = body(dst + o, src + o, std::min(part, n - o));
        });
    }
    for (auto &t : th) t.join();
    return *std::max_element(mx.begin(), mx.end());
}
}  // namespace

// ----------------------------------------------------------------------------------------------
// CUDA IPC mappings of neighbouring shards that live in other processes (gpb_link).  Two facts of cudaIpc* shape this:
//   * a handle stands for a whole DRIVER allocation; the runtime packs small cudaMalloc requests into shared blocks, so a
//     handle is taken for the base of the block (cuMemGetAddressRange) and the array is addressed base + offset;
//   * a process may open a given allocation only once: mappings are reference-counted by handle.
namespace {
typedef CUresult (*AddrRangeFn)(CUdeviceptr *, size_t *, CUdeviceptr);
AddrRangeFn addr_range_fn()
{
    static AddrRangeFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (AddrRangeFn)p;
    }
    return fn;
}
// handle of the allocation that contains `ptr` and the offset of `ptr` inside it
int ipc_export(const void *ptr, unsigned char handle[64], uint64_t *offset)
{
    CUdeviceptr base = (CUdeviceptr)(uintptr_t)ptr;
    size_t size = 0;
    AddrRangeFn fn = addr_range_fn();
    if (fn && fn(&base, &size, (CUdeviceptr)(uintptr_t)ptr) != CUDA_SUCCESS) base = (CUdeviceptr)(uintptr_t)ptr;
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, (void *)(uintptr_t)base) != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    memcpy(handle, &h, 64);
    *offset = (uint64_t)((uintptr_t)ptr - (uintptr_t)base);
    return 0;
}
struct IpcMappings {
    std::mutex mu;
    std::map<std::string, std::pair<void *, int>> open;   // handle bytes -> (mapped base, references)
    cudaError_t acquire(const unsigned char handle[64], void **base)
    {
        std::lock_guard<std::mutex> g(mu);
        const std::string key((const char *)handle, 64);
        auto it = open.find(key);
        if (it != open.end()) {
            ++it->second.second;
            *base = it->second.first;
            return cudaSuccess;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, handle, 64);
        cudaError_t e = cudaIpcOpenMemHandle(base, h, cudaIpcMemLazyEnablePeerAccess);
        if (e == cudaSuccess) open[key] = {*base, 1};
        return e;
    }
    // A mapping whose last user is gone stays open: the owner's device-memory cache hands the same allocation to its next
    // solver (the next trace, the next benchmark repetition), and mapping / unmapping a multi-GB allocation costs tenths of a
    // second.  gpb_release_cached() closes the idle ones.
    void release(void *base)
    {
        std::lock_guard<std::mutex> g(mu);
        for (auto &kv : open)
            if (kv.second.first == base) {
                if (kv.second.second > 0) --kv.second.second;
                return;
            }
    }
    void close_idle()
    {
        std::lock_guard<std::mutex> g(mu);
        for (auto it = open.begin(); it != open.end();) {
            if (it->second.second == 0) {
                cudaError_t e = cudaIpcCloseMemHandle(it->second.first);
                if (e != cudaSuccess) {
                    if (getenv("GPB_DEBUG")) fprintf(stderr, "[gpb] cudaIpcCloseMemHandle: %s\n", cudaGetErrorString(e));
                    cudaGetLastError();
                }
                it = open.erase(it);
            } else {
                ++it;
            }
        }
    }
};
IpcMappings g_ipc;
}  // namespace

// ----------------------------------------------------------------------------------------------
struct SolverBase {
    virtual ~SolverBase() {}
    virtual int run(int n) = 0;
    virtual int half_step(int phase, int part) = 0;
    virtual int reset() = 0;
    virtual int get_receivers(void *out, size_t bytes) = 0;
    virtual int get_snapshot(int idx, void *out6[6], size_t bytes_each) = 0;
    virtual int get_tline(int idx, void *v, void *i, size_t bytes_each) = 0;
    virtual int get_field(int comp, void *out, size_t bytes) = 0;
    virtual int set_field(int comp, const void *in, size_t bytes) = 0;
    virtual int halo(int which, void **a, void **b, size_t *bytes) = 0;
    virtual int profile(int n, double *ms4) = 0;
    virtual std::string kernel_path() const = 0;
    virtual int set_points(const gpb_model_t &m) = 0;
    virtual int link_info(gpb_link_t *out) = 0;
    virtual int link(const gpb_link_t *left, const gpb_link_t *right) = 0;
    int device = 0;
    int iteration = 0;
    double elapsed = 0;
    uint64_t mem = 0;
    uint64_t launches = 0;
    cudaStream_t stream = nullptr;
};

struct gpb_solver {
    SolverBase *impl;
};

template <typename R>
struct Solver : SolverBase {
    // model
    int nx, ny, nz, x_start, nplanes, pitch, iterations, nmat, maxpoles, form, order;
    long long plane, narr;  // elements per plane / per padded array (nplanes+2 planes)
    int idbytes;
    bool tabsmem;
    bool use_v4 = false;       // vectorised non-dispersive path (gpb_kernels_v4.cuh)
    bool use_tma = false;      // TMA-staged path (gpb_kernels_tma.cuh)
    int v4_xchunk = 16;        // planes marched by one thread of the v4 kernels
    int tma_ty = 0, tma_tz = 0, tma_stages = 0, tma_xchunk = 16, tma_pw = 0;   // tile, ring depth, planes per item, dedicated producer warp
    bool tma_persist = true;   // persistent CTAs with a continuous TMA pipeline across work items
    // diagnostic switches, read from the environment ONCE in build()
    bool tma_nofast = false, tma_zsplit = false, tma_znocoop = false, tma_nosplit = false;
    int tma_pf = 2;
    int *d_sched = 0;          // [2] work-item scheduler state of the persistent TMA kernels
    int sm_count = 148;
    TmaMaps4 maps_e, maps_h;
    int setup_tma();
    int launch_tma(int phase, int p0, int p1, int peer_store = 0);   // 1: boundary plane also stored to the neighbour, 2: and announced by the kernel
    // H and E half-steps of an iteration in ONE launch (gpb_kernels_pair.cuh): items of both phases from one queue, the E items
    // of a chunk about two chunks behind its H items, so E finds its operands in L2
    bool pair_he = false;
    int pair_xchunk = 4, pair_lag = 1;
    unsigned *d_progress = nullptr;   // [chunks + 1]: finished H warps per chunk, [chunks] = time-out flag
    int n_he_chunks = 0;
    int launch_pair();
    unsigned zslabs_e = 0, zslabs_h = 0;  // z slabs (bit per slab) handled by k_pml_slabs on the v4 path
    uint64_t graph_launches = 0, graph_k_launches = 0;
    size_t smem_bytes;
    // device memory
    std::vector<void *> allocs;
    R *F[6] = {0, 0, 0, 0, 0, 0};
    void *ID[6] = {0, 0, 0, 0, 0, 0};
    Coef4<R> *coefE = 0, *coefH = 0;
    R *srcE = 0, *srcH = 0;
    Cplx<R> *T[3] = {0, 0, 0};   // treal: the allocation holds R[maxpoles][narr] instead
    Cplx<R> *dcoef = 0;          // treal: R[nmat][maxpoles][3]
    bool treal = false;          // all dispersive coefficients real (Debye media): real-valued T
    // small grids: n iterations in ONE cooperative launch (gpb_kernels_coop.cuh)
    bool coop = false;
    int coop_grid = 0, coop_xchunk = 1, coop_gx = 0, coop_gy = 0;
    size_t coop_smem = 0;
    const void *coop_kernel() const;
    int setup_coop();
    int launch_coop(int n);
    bool tma_disp = false;       // dispersive E half-step on the TMA kernels
    int tma_tpf = 2;
    int *d_iter = 0;  // [0] current, [1] next
    // receivers / sources / snapshots
    int nrx = 0;
    int *d_rxc = 0;
    R *d_rxs = 0;
    int nsrc = 0, ntl = 0;
    SrcDev<R> *d_srcs = 0;
    TLDev<R> *d_tls = 0;
    std::vector<TLDev<R>> h_tls;
    std::vector<int> h_src_plane, h_src_phase;   // plane and phase (0: magnetic dipole, 1: voltage source / Hertzian dipole) of every point source
    std::vector<std::vector<R>> tl_v0, tl_c0;
    std::vector<R> tl_abc0;
    bool has_hsrc = false, has_esrc = false;
    std::vector<gpb_snapshot_t> snaps;
    std::vector<SnapDev<R>> snapdev;
    // phases
    PhaseParams<R> ph_h, ph_e;
    PointParams<R> pp;
    std::vector<std::pair<R *, size_t>> phis;
    bool v4_ok = true;
    // linked x-slab shards (halo pushed into the neighbours' ghost planes over peer memory, gpb_link)
    struct Peer {
        bool present = false;
        R *F = nullptr;              // base of the neighbour's field allocation (component c at F + c * narr_peer)
        unsigned *flags = nullptr;   // the neighbour's flag words
        long long narr = 0;
        int x_start = 0, nplanes = 0;
        void *ipc_F = nullptr, *ipc_flags = nullptr;   // mappings opened with cudaIpcOpenMemHandle (closed on unlink)
    };
    Peer left, right;
    bool linked = false;
    unsigned *d_flags = nullptr;   // [GPB_NFLAGS] written by the neighbours (peer stores) and by this shard's own kernels
    unsigned long long link_timeout_ns = 20000000000ull;
    bool link_nofused = false, link_late = false, link_early_credit = false;   // GPB_NO_FUSED_PUSH / GPB_LATE_SIGNAL / GPB_EARLY_CREDIT, read in link()
    bool rx_on_first_plane = false;                  // a receiver on my first plane reads the H ghost plane in the step prologue
    bool tl_on_first_plane() const
    {
        for (int t = 0; t < ntl; ++t)
            if (h_tls[t].i == x_start) return true;
        return false;
    }
    bool credit_src_on_first_plane() const           // (then the E half-step does not announce early: keep the credit late as well)
    {
        for (size_t q = 0; q < h_src_plane.size(); ++q)
            if (h_src_phase[q] == 1 && h_src_plane[q] == x_start) return true;
        return false;
    }
    bool snap_needs_right = false;
    bool snap_unlinked_ok = true;    // every snapshot cell of this slab averages planes of this slab only
    bool no_overlap_now = false;     // profile(): time the two half-step kernels one after the other   // some snapshot cell of this shard reads planes of the right neighbour
    int begin_run(int n);
    int enqueue_iterations(int n);
    int end_run();
    int finish_run();
    int unlink();
    int enqueue_linked_step(bool with_snap);
    int check_link_timeout();
    // graph
    cudaGraphExec_t graph = nullptr;    // one iteration
    cudaGraphExec_t graph_k = nullptr;  // graph_iters iterations in one graph (fewer graph boundaries; small grids)
    int graph_iters = 16;
    bool fuse_begin = true;             // GPB_FUSE_BEGIN=0: the step prologue stays a launch of its own inside the multi-iteration graph
    bool use_pdl = true;                // GPB_PDL=0: no programmatic dependent launches inside the step (gpb_kernels.cuh)
    bool host_driven = false;           // advanced through gpb_half_step
    bool pdl_on() const { return use_pdl && !linked && !host_driven; }
    bool use_graph = true;
    void drop_graphs()
    {
        if (graph) { cudaGraphExecDestroy(graph); graph = nullptr; }
        if (graph_k) { cudaGraphExecDestroy(graph_k); graph_k = nullptr; }
    }
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    ~Solver()
    {
        cudaSetDevice(device);
        drop_graphs();
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (stream) cudaStreamSynchronize(stream);   // cached blocks are handed to the next solver, which runs on another stream
        unlink();
        if (d_flags) cudaFree(d_flags);
        for (void *p : allocs) g_pool.free(p);
        if (stream) cudaStreamDestroy(stream);
        cudaError_t e = cudaGetLastError();   // nothing of the tear-down may linger as the "last error" of the next call
        if (e != cudaSuccess && getenv("GPB_DEBUG")) fprintf(stderr, "[gpb] solver tear-down left: %s\n", cudaGetErrorString(e));
    }

    template <typename T_>
    int dalloc(T_ **p, size_t n, bool zero = true)
    {
        void *q = nullptr;
        size_t bytes = std::max<size_t>(n, 1) * sizeof(T_);
        CK(g_pool.alloc(&q, bytes, device));
        allocs.push_back(q);
        mem += bytes;
        if (zero) CK(cudaMemsetAsync(q, 0, bytes, stream));
        *p = (T_ *)q;
        return 0;
    }
    template <typename T_>
    int upload(T_ **p, const T_ *h, size_t n)
    {
        if (dalloc(p, n, false)) return 1;
        if (n) CK(cudaMemcpyAsync(*p, h, n * sizeof(T_), cudaMemcpyHostToDevice, stream));
        return 0;
    }

    int upload_ids(const gpb_model_t &m);
    int build(const gpb_model_t &m);
    int setup_pml(const gpb_model_t &m);
    int setup_points(const gpb_model_t &m);
    int set_smem_attributes();
    void set_boxes();

    template <typename IDT>
    int launch_h(int p0, int p1);
    template <typename IDT>
    int launch_e(int p0, int p1);
    int launch_phase(int phase, int p0, int p1);
    int launch_sources(int phase, int p0, int p1, int t0, int t1, bool begin_next = false);
    int launch_begin();
    int launch_snapshots();
    int enqueue_step(bool with_snap, bool with_begin = true, bool begin_next = false);
    bool snapshot_due(int it) const
    {
        for (auto &s : snaps)
            if (s.time == it + 1) return true;
        return false;
    }

    int run(int n) override;
    int half_step(int phase, int part) override;
    int reset() override;
    int get_receivers(void *out, size_t bytes) override;
    int get_snapshot(int idx, void *out6[6], size_t bytes_each) override;
    int get_tline(int idx, void *v, void *i, size_t bytes_each) override;
    int get_field(int comp, void *out, size_t bytes) override;
    int set_field(int comp, const void *in, size_t bytes) override;
    int halo(int which, void **a, void **b, size_t *bytes) override;
    int profile(int n, double *ms4) override;
    std::string kernel_path() const override;
    int link_info(gpb_link_t *out) override;
    int link(const gpb_link_t *left, const gpb_link_t *right) override;
    int set_points(const gpb_model_t &m) override;
};

static int choose_pitch(int nzp1)
{
    // rows start on a 128-byte line when the row is long enough to make that worthwhile; always a
    // multiple of 16 elements so that even 1-byte ID rows satisfy the 16-byte stride rule of TMA
    if (nzp1 >= 48) return (nzp1 + 31) / 32 * 32;
    return (nzp1 + 15) / 16 * 16;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 4-D map over a component triple [ncomp][planes][rows][pitch] (one allocation) with a (b0 x b1 x 1 x bc) box
static int make_map(CUtensorMap *out, void *base, CUtensorMapDataType dt, size_t es, int pitch, int rows, int planes, long long comp_stride, int b0, int b1, int bc)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail("cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[4] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)planes, 3};
    const cuuint64_t strides[3] = {(cuuint64_t)pitch * es, (cuuint64_t)pitch * rows * es, (cuuint64_t)comp_stride * es};
    const cuuint32_t box[4] = {(cuuint32_t)b0, (cuuint32_t)b1, 1, (cuuint32_t)bc};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    if (const char *e = getenv("GPB_TMA_L2")) promo = atoi(e) == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : (atoi(e) == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : (atoi(e) == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : promo));
    CUresult r = fn(out, dt, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with code %d (pitch %d rows %d planes %d box %d x %d x 1 x %d)", (int)r, pitch, rows, planes, b0, b1, bc);
    return 0;
}

template <typename R>
int Solver<R>::upload_ids(const gpb_model_t &m)
{
    if (!m.ID) {
        // homogeneous domain: every edge carries the same material
        if (m.uniform_id < 0 || m.uniform_id >= nmat) return fail("uniform_id %d outside the %d materials", m.uniform_id, nmat);
        void *idbase = nullptr;
        CK(g_pool.alloc(&idbase, (size_t)narr * idbytes * 6, device));   // one allocation, like F
        allocs.push_back(idbase);
        mem += (size_t)narr * idbytes * 6;
        for (int c = 0; c < 6; ++c) {
            void *dst = (char *)idbase + (size_t)c * narr * idbytes;
            ID[c] = dst;
            if (idbytes == 1) k_fill_ids<uint8_t><<<148 * 8, 256, 0, stream>>>((uint8_t *)dst, narr, (uint8_t)m.uniform_id);
            else if (idbytes == 2) k_fill_ids<uint16_t><<<148 * 8, 256, 0, stream>>>((uint16_t *)dst, narr, (uint16_t)m.uniform_id);
            else k_fill_ids<uint32_t><<<148 * 8, 256, 0, stream>>>((uint32_t *)dst, narr, (uint32_t)m.uniform_id);
            CK(cudaGetLastError());
        }
        CK(cudaStreamSynchronize(stream));
        return 0;
    }
    // uint32 rows -> narrowed by several host threads into a pinned bounce buffer -> device stage -> placed into the pitched
    // layout, double-buffered: the host pass over chunk q+1 overlaps the PCIe transfer and the placement kernel of chunk q
    // (a plain cudaMemcpy from the pageable NumPy array ran at 11 GB/s: 59 ms of the 300^3 model's 62 ms set-up).  Models
    // with more than 65536 materials keep 32-bit IDs: copied as they are and "narrowed" (32 -> 32) on the device.
    const long long rows_per_plane = ny + 1;
    const long long total_rows = rows_per_plane * nplanes;
    const size_t row_elems = (size_t)(nz + 1);
    const bool host_narrow = idbytes < 4 && !getenv("GPB_DEVICE_NARROW");
    const size_t xfer_elem = host_narrow ? (size_t)idbytes : 4;   // bytes per ID that cross PCIe
    const long long chunk_rows = std::max<long long>(1, (long long)(kBounceBytes / (row_elems * xfer_elem)));
    // a shard may point into the caller's global ID array (component stride of the whole grid): no second host copy of the slab
    const size_t comp_stride = m.id_comp_stride > 0 ? (size_t)m.id_comp_stride : (size_t)total_rows * (nz + 1);
    void *stage[2] = {nullptr, nullptr};
    unsigned *d_max = nullptr;
    void *bounce[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    for (int b = 0; b < 2; ++b) {
        CK(g_pool.alloc(&stage[b], kBounceBytes, device));
        CK(g_pool.alloc_pinned(&bounce[b], kBounceBytes));
        CK(cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming));
    }
    CK(g_pool.alloc((void **)&d_max, sizeof(unsigned), device));
    CK(cudaMemsetAsync(d_max, 0, sizeof(unsigned), stream));
    void *idbase = nullptr;
    CK(g_pool.alloc(&idbase, (size_t)narr * idbytes * 6, device));   // one allocation, like F
    allocs.push_back(idbase);
    mem += (size_t)narr * idbytes * 6;
    CK(cudaMemsetAsync(idbase, 0, (size_t)narr * idbytes * 6, stream));
    // (configured CPUs, not std::thread::hardware_concurrency(): that one follows the calling thread's affinity mask, which is a
    //  single core once the OpenMP runtime has pinned the main thread)
    const int nthreads = (int)std::max(1l, std::min(8l, sysconf(_SC_NPROCESSORS_CONF) / 2));
    uint32_t host_max = 0;
    long long q = 0;
    for (int c = 0; c < 6; ++c) {
        void *dst = (char *)idbase + (size_t)c * narr * idbytes;
        ID[c] = dst;
        for (long long r0 = 0; r0 < total_rows; r0 += chunk_rows, ++q) {
            const int b = (int)(q & 1);
            const long long rows = std::min(chunk_rows, total_rows - r0);
            const size_t n = (size_t)rows * row_elems;
            const uint32_t *src = m.ID + (size_t)c * comp_stride + (size_t)r0 * row_elems;
            if (q >= 2) CK(cudaEventSynchronize(ev[b]));   // the transfer out of this bounce buffer two chunks ago
            if (host_narrow) {
                const uint32_t mx = idbytes == 1 ? parallel_narrow((uint8_t *)bounce[b], src, n, nthreads) : parallel_narrow((uint16_t *)bounce[b], src, n, nthreads);
                host_max = std::max(host_max, mx);
            } else {
                parallel_copy((char *)bounce[b], (const char *)src, n * 4, nthreads);
            }
            CK(cudaMemcpyAsync(stage[b], bounce[b], n * xfer_elem, cudaMemcpyHostToDevice, stream));
            CK(cudaEventRecord(ev[b], stream));
            const int blocks = (int)std::min<long long>(((long long)n + 255) / 256, 148 * 16);
            char *d = (char *)dst + ((size_t)plane + (size_t)r0 * pitch) * idbytes;   // plane 0 of the array is the ghost plane
            if (host_narrow) {
                if (idbytes == 1) k_place_ids<uint8_t><<<blocks, 256, 0, stream>>>((const uint8_t *)stage[b], (uint8_t *)d, rows, nz + 1, pitch);
                else k_place_ids<uint16_t><<<blocks, 256, 0, stream>>>((const uint16_t *)stage[b], (uint16_t *)d, rows, nz + 1, pitch);
            } else if (idbytes == 1) k_narrow_ids<uint8_t><<<blocks, 256, 0, stream>>>((const uint32_t *)stage[b], (uint8_t *)d, rows, nz + 1, pitch, d_max);
            else if (idbytes == 2) k_narrow_ids<uint16_t><<<blocks, 256, 0, stream>>>((const uint32_t *)stage[b], (uint16_t *)d, rows, nz + 1, pitch, d_max);
            else k_narrow_ids<uint32_t><<<blocks, 256, 0, stream>>>((const uint32_t *)stage[b], (uint32_t *)d, rows, nz + 1, pitch, d_max);
            CK(cudaGetLastError());
        }
    }
    CK(cudaStreamSynchronize(stream));
    for (int b = 0; b < 2; ++b) {
        cudaEventDestroy(ev[b]);
        g_pool.free(stage[b]);
        g_pool.free_pinned(bounce[b]);
    }
    unsigned mx = 0;
    CK(cudaMemcpy(&mx, d_max, sizeof mx, cudaMemcpyDeviceToHost));
    g_pool.free(d_max);
    mx = std::max(mx, host_max);
    if ((int)mx >= nmat) return fail("ID array references material %u but only %d materials were given", mx, nmat);
    return 0;
}

template <typename R>
void Solver<R>::set_boxes()
{
    // Index ranges of the CPU solver (fields_updates_ext.pyx), see SURVEY.md section 8(a).
    auto setbox = [](Box &b, int i0, int i1, int j0, int j1, int k0, int k1) {
        b.lo[0] = i0; b.hi[0] = i1; b.lo[1] = j0; b.hi[1] = j1; b.lo[2] = k0; b.hi[2] = k1;
    };
    Box none;
    setbox(none, 0, 0, 0, 0, 0, 0);
    // electric, :55-107 / :146-179 -- the same three boxes cover 3-D and every 2-D mode
    setbox(ph_e.box[0], 0, nx, 1, ny, 1, nz);
    setbox(ph_e.box[1], 1, nx, 0, ny, 1, nz);
    setbox(ph_e.box[2], 1, nx, 1, ny, 0, nz);
    // magnetic
    if (nx == 1 || ny == 1 || nz == 1) {  // :376-399
        ph_h.box[0] = ph_h.box[1] = ph_h.box[2] = none;
        if (ny == 1 || nz == 1) setbox(ph_h.box[0], 1, nx, 0, ny, 0, nz);
        if (nx == 1 || nz == 1) setbox(ph_h.box[1], 0, nx, 1, ny, 0, nz);
        if (nx == 1 || ny == 1) setbox(ph_h.box[2], 0, nx, 0, ny, 1, nz);
    } else {  // :401-412  Hx[i+1,j,k], Hy[i,j+1,k], Hz[i,j,k+1] for i<nx, j<ny, k<nz
        setbox(ph_h.box[0], 1, nx + 1, 0, ny, 0, nz);
        setbox(ph_h.box[1], 0, nx, 1, ny + 1, 0, nz);
        setbox(ph_h.box[2], 0, nx, 0, ny, 1, nz + 1);
    }
}

template <typename R>
int Solver<R>::setup_pml(const gpb_model_t &m)
{
    ph_e.nslabs = ph_h.nslabs = 0;
    if (m.npml > kMaxSlabs) return fail("at most %d PML slabs are supported, got %d", kMaxSlabs, m.npml);
    for (int n = 0; n < m.npml; ++n) {
        const gpb_pml_t &s = m.pmls[n];
        if (s.direction < 0 || s.direction > 5) return fail("PML slab %d: bad direction %d", n, s.direction);
        const int axis = s.direction % 3, minus = s.direction < 3;
        const int t = s.thickness;
        const int ext[3] = {s.xf - s.xs, s.yf - s.ys, s.zf - s.zs};
        if (t <= 0 || ext[axis] != t) return fail("PML slab %d: thickness %d does not match its extent %d", n, t, ext[axis]);
        const int lo0[3] = {s.xs, s.ys, s.zs}, hi0[3] = {s.xf, s.yf, s.zf};
        const R *tabs[8];
        const void *src[8] = {s.ERA, s.ERB, s.ERE, s.ERF, s.HRA, s.HRB, s.HRE, s.HRF};
        for (int q = 0; q < 8; ++q) {
            R *d = nullptr;
            if (upload(&d, (const R *)src[q], (size_t)order * t)) return 1;
            tabs[q] = d;
        }
        for (int phase = 0; phase < 2; ++phase) {  // 0 electric, 1 magnetic
            SlabDev<R> sd;
            memset(&sd, 0, sizeof sd);
            for (int a = 0; a < 3; ++a) { sd.lo[a] = lo0[a]; sd.hi[a] = hi0[a]; }
            // field index along the slab axis: electric minus  af - a  -> [as+1, af+1) ; magnetic minus af-(a+1) -> [as, af)
            // (pml_updates_electric_HORIPML_ext.pyx:73 / pml_updates_magnetic_HORIPML_ext.pyx:69); plus: a + as
            if (minus) {
                if (phase == 0) { sd.lo[axis] = lo0[axis] + 1; sd.hi[axis] = hi0[axis] + 1; sd.dref = hi0[axis]; }
                else { sd.dref = hi0[axis] - 1; }
            } else {
                sd.dref = lo0[axis];
            }
            sd.axis = axis;
            sd.minus = minus;
            sd.t = t;
            // restrict to the planes this handle owns
            const int glo = sd.lo[0], ghi = sd.hi[0];
            sd.lo[0] = std::max(glo, x_start);
            sd.hi[0] = std::min(ghi, x_start + nplanes);
            if (sd.hi[0] <= sd.lo[0]) continue;
            const long long n0 = sd.hi[0] - sd.lo[0];
            sd.n1 = sd.hi[1] - sd.lo[1];
            // x / y slabs: Phi rows padded to the field pitch so the vectorised kernels can use aligned
            // 128-bit accesses; z slabs: rows of `thickness` cells padded to whole 4-cell groups of the field rows
            sd.ko = axis == 2 ? (sd.lo[2] & ~3) : sd.lo[2];
            sd.n2 = axis == 2 ? (((sd.hi[2] + 3) & ~3) - sd.ko) : (sd.lo[2] == 0 ? pitch : sd.hi[2] - sd.lo[2]);
            if (axis != 2 && sd.lo[2] != 0) v4_ok = false;
            sd.ostride = n0 * sd.n1 * sd.n2;
            R *phi = nullptr;
            const size_t nphi = (size_t)sd.ostride * 2 * order;
            if (dalloc(&phi, nphi)) return 1;
            phis.push_back({phi, nphi});
            sd.phi = phi;
            const int o = phase == 0 ? 0 : 4;
            sd.RA = tabs[o]; sd.RB = tabs[o + 1]; sd.RE = tabs[o + 2]; sd.RF = tabs[o + 3];
            sd.d = (R)(float)s.d;  // the reference kernels take `float d` (e.g. pml_updates_electric_HORIPML_ext.pyx:48)
            sd.inv_d = (R)1 / sd.d;
            PhaseParams<R> &ph = phase == 0 ? ph_e : ph_h;
            ph.slab[ph.nslabs++] = sd;
        }
    }
    return 0;
}

// New sources / receivers / transmission lines / snapshots on the resident grid (the traces of a B-scan whose geometry is fixed
// only step these, model_build_run.py:294-330): fields, PML and T state are cleared, ID and coefficient arrays stay on the device.
template <typename R>
int Solver<R>::set_points(const gpb_model_t &m)
{
    CK(cudaSetDevice(device));
    if (m.nx != nx || m.ny != ny || m.nz != nz || m.iterations != iterations || m.nmaterials != nmat)
        return fail("gpb_set_points: the model has another grid, iteration count or material table than the resident one");
    CK(cudaStreamSynchronize(stream));
    drop_graphs();   // the captured launches hold the old point arrays
    has_hsrc = has_esrc = false;
    if (setup_points(m)) return 1;
    snap_unlinked_ok = true;
    for (auto &sn : snaps)
        for (int i = 0; i < sn.nx; ++i) {
            const int gi = sn.xs + i * sn.dx;
            if (gi >= x_start && gi < x_start + nplanes && gi + sn.dx >= x_start + nplanes) snap_unlinked_ok = false;
        }
    if (linked && !snaps.empty())
        for (auto &sn : snaps)
            for (int i = 0; i < sn.nx; ++i) {
                const int gi = sn.xs + i * sn.dx;
                if (gi < x_start || gi >= x_start + nplanes || gi + sn.dx < x_start + nplanes) continue;
                if (!right.present || gi + sn.dx >= right.x_start + right.nplanes) return fail("snapshot plane %d + %d lies beyond the right neighbour's planes", gi, sn.dx);
            }
    return reset();
}

template <typename R>
int Solver<R>::setup_points(const gpb_model_t &m)
{
    nrx = m.nrx;
    rx_on_first_plane = false;
    if (nrx) {
        if (upload(&d_rxc, m.rxcoords, (size_t)nrx * 3)) return 1;
        for (int r = 0; r < nrx; ++r) {
            const int *c = m.rxcoords + 3 * r;
            if (c[0] == m.x_start) rx_on_first_plane = true;
            if (c[0] < 0 || c[0] > nx || c[1] < 0 || c[1] > ny || c[2] < 0 || c[2] > nz) return fail("receiver %d outside the grid", r);
        }
    }
    if (dalloc(&d_rxs, (size_t)GPB_NRXOUT * iterations * std::max(nrx, 1))) return 1;
    nsrc = m.nsources;
    std::vector<SrcDev<R>> hs(nsrc);
    h_src_plane.clear();
    h_src_phase.clear();
    const double dd[3] = {m.dx, m.dy, m.dz};
    for (int s = 0; s < nsrc; ++s) {
        const gpb_source_t &g = m.sources[s];
        SrcDev<R> &d = hs[s];
        if (g.kind < 0 || g.kind > 2 || g.polarisation < 0 || g.polarisation > 2) return fail("source %d: bad kind/polarisation", s);
        if (g.i < 0 || g.i > nx || g.j < 0 || g.j > ny || g.k < 0 || g.k > nz) return fail("source %d outside the grid", s);
        d.kind = g.kind; d.i = g.i; d.j = g.j; d.k = g.k; d.pol = g.polarisation;
        h_src_plane.push_back(g.i);
        h_src_phase.push_back(g.kind == GPB_SRC_MAGNETIC ? 0 : 1);
        d.it_first = g.it_first; d.it_last = g.it_last; d.hard = 0; d.f1 = 0; d.f2 = 0;
        R *w = nullptr;
        if (upload(&w, (const R *)g.waveform, (size_t)iterations)) return 1;
        d.wave = w;
        if (g.kind == GPB_SRC_HERTZIAN) {  // sources.py:181-193
            d.f1 = (R)g.param;
            d.f2 = (R)(1 / (m.dx * m.dy * m.dz));
            has_esrc = true;
        } else if (g.kind == GPB_SRC_MAGNETIC) {  // sources.py:220-232
            d.f2 = (R)(1 / (m.dx * m.dy * m.dz));
            has_hsrc = true;
        } else {  // sources.py:97-117
            if (g.param != 0) {
                const int p = g.polarisation;
                const double d1 = p == 0 ? m.dy : m.dx, d2 = p == 2 ? m.dy : m.dz;
                d.f2 = (R)(1 / (g.param * d1 * d2));
            } else {
                d.hard = 1;
                d.f2 = (R)dd[g.polarisation];
            }
            has_esrc = true;
        }
    }
    if (nsrc && upload(&d_srcs, hs.data(), (size_t)nsrc)) return 1;
    ntl = m.ntlines;
    h_tls.resize(ntl);
    tl_v0.resize(ntl);
    tl_c0.resize(ntl);
    tl_abc0.resize(2 * (size_t)ntl);
    const double c0 = 299792458.0;
    for (int t = 0; t < ntl; ++t) {
        const gpb_tline_t &g = m.tlines[t];
        TLDev<R> &d = h_tls[t];
        if (g.nl < 2 || g.antpos >= g.nl || g.srcpos < 1 || g.srcpos >= g.nl) return fail("transmission line %d: bad line geometry", t);
        d.i = g.i; d.j = g.j; d.k = g.k; d.pol = g.polarisation;
        d.it_first = g.it_first; d.it_last = g.it_last;
        d.nl = g.nl; d.srcpos = g.srcpos; d.antpos = g.antpos;
        d.cdtdl = c0 * m.dt / g.dl;                           // sources.py:368-373
        d.coefV = g.resistance * d.cdtdl;
        d.coefI = (1 / g.resistance) * d.cdtdl;
        d.h = (c0 * m.dt - g.dl) / (c0 * m.dt + g.dl);        // :355
        const int p = g.polarisation;
        d.dpol = (R)dd[p];
        tl_v0[t].assign((const R *)g.voltage0, (const R *)g.voltage0 + g.nl);
        tl_c0[t].assign((const R *)g.current0, (const R *)g.current0 + g.nl);
        tl_abc0[2 * t] = (R)g.abcv0;
        tl_abc0[2 * t + 1] = (R)g.abcv1;
        R *v, *c, *abc, *ww, *wh, *vt, *itot;
        if (upload(&v, tl_v0[t].data(), (size_t)g.nl) || upload(&c, tl_c0[t].data(), (size_t)g.nl) ||
            upload(&abc, &tl_abc0[2 * t], 2) || upload(&ww, (const R *)g.wave_whole, (size_t)iterations) ||
            upload(&wh, (const R *)g.wave_half, (size_t)iterations) || dalloc(&vt, (size_t)iterations) || dalloc(&itot, (size_t)iterations))
            return 1;
        d.voltage = v; d.current = c; d.abcv = abc; d.wave_whole = ww; d.wave_half = wh; d.Vtotal = vt; d.Itotal = itot;
        has_hsrc = has_esrc = true;
    }
    if (ntl && upload(&d_tls, h_tls.data(), (size_t)ntl)) return 1;
    // snapshots
    snaps.assign(m.snapshots, m.snapshots + m.nsnapshots);
    snapdev.resize(snaps.size());
    for (size_t n = 0; n < snaps.size(); ++n) {
        const gpb_snapshot_t &s = snaps[n];
        if (s.nx <= 0 || s.ny <= 0 || s.nz <= 0 || s.dx < 1 || s.dy < 1 || s.dz < 1 || s.xs < 0 || s.ys < 0 || s.zs < 0 ||
            s.xs + s.nx * s.dx > nx || s.ys + s.ny * s.dy > ny || s.zs + s.nz * s.dz > nz)
            return fail("snapshot %zu does not fit the grid", n);
        SnapDev<R> &d = snapdev[n];
        d.xs = s.xs; d.ys = s.ys; d.zs = s.zs; d.dx = s.dx; d.dy = s.dy; d.dz = s.dz; d.nx = s.nx; d.ny = s.ny; d.nz = s.nz;
        for (int c = 0; c < 6; ++c)
            if (dalloc(&d.out[c], (size_t)s.nx * s.ny * s.nz)) return 1;
    }
    return 0;
}

template <typename R>
int Solver<R>::build(const gpb_model_t &m)
{
    nx = m.nx; ny = m.ny; nz = m.nz; x_start = m.x_start; nplanes = m.nx_planes;
    iterations = m.iterations; nmat = m.nmaterials; maxpoles = m.maxpoles;
    form = m.pml_formulation; order = m.npml ? m.pml_order : 1;
    if (nx < 1 || ny < 1 || nz < 1) return fail("grid must have at least one cell per axis");
    if (x_start < 0 || nplanes < 1 || x_start + nplanes > nx + 1) return fail("owned plane range [%d, %d) outside [0, %d]", x_start, x_start + nplanes, nx);
    if (nmat < 1 || !m.updatecoeffsE || !m.updatecoeffsH) return fail("material tables missing");
    if (m.npml && (order < 1 || order > 2 || form < 0 || form > 1)) return fail("unsupported PML formulation/order %d/%d", form, order);
    if (maxpoles < 0 || (maxpoles > 0 && !m.updatecoeffsdispersive)) return fail("dispersive coefficient table missing");
    // GPB_TIMING=1: host wall clock of the set-up phases on stderr (profiles/e2e_breakdown.py)
    const bool timing = getenv("GPB_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto tick = [&](const char *what) {
        if (!timing) return;
        cudaStreamSynchronize(stream);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[gpb] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&ev0));
    CK(cudaEventCreate(&ev1));
    pitch = choose_pitch(nz + 1);
    plane = (long long)(ny + 1) * pitch;
    narr = plane * (nplanes + 2);
    idbytes = nmat <= 256 ? 1 : (nmat <= 65536 ? 2 : 4);
    if (getenv("GPB_ID_BYTES")) idbytes = std::max(idbytes, atoi(getenv("GPB_ID_BYTES")) >= 4 ? 4 : (atoi(getenv("GPB_ID_BYTES")) >= 2 ? 2 : 1));
    use_graph = !getenv("GPB_NO_GRAPH");
    use_pdl = !(getenv("GPB_PDL") && atoi(getenv("GPB_PDL")) == 0);
    if (getenv("GPB_FUSE_BEGIN")) fuse_begin = atoi(getenv("GPB_FUSE_BEGIN")) != 0;
    if (getenv("GPB_GRAPH_ITERS")) graph_iters = std::max(1, std::min(64, atoi(getenv("GPB_GRAPH_ITERS"))));
    // all six components in one allocation: the TMA kernels address a triple (E or H) as one 4-D tensor
    if (dalloc(&F[0], (size_t)narr * 6)) return 1;
    for (int c = 1; c < 6; ++c) F[c] = F[0] + (size_t)c * narr;
    tick("stream + field arrays");
    if (upload_ids(m)) return 1;
    tick("material IDs");
    // coefficient rows: [CA, CBx, CBy, CBz | srce] (materials.py:200-201)
    std::vector<Coef4<R>> hE(nmat), hH(nmat);
    std::vector<R> sE(nmat), sH(nmat);
    const R *cE = (const R *)m.updatecoeffsE, *cH = (const R *)m.updatecoeffsH;
    for (int q = 0; q < nmat; ++q) {
        hE[q].a = cE[5 * q]; hE[q].bx = cE[5 * q + 1]; hE[q].by = cE[5 * q + 2]; hE[q].bz = cE[5 * q + 3]; sE[q] = cE[5 * q + 4];
        hH[q].a = cH[5 * q]; hH[q].bx = cH[5 * q + 1]; hH[q].by = cH[5 * q + 2]; hH[q].bz = cH[5 * q + 3]; sH[q] = cH[5 * q + 4];
    }
    if (upload(&coefE, hE.data(), (size_t)nmat) || upload(&coefH, hH.data(), (size_t)nmat) ||
        upload(&srcE, sE.data(), (size_t)nmat) || upload(&srcH, sH.data(), (size_t)nmat))
        return 1;
    CK(cudaStreamSynchronize(stream));  // host vectors above go out of scope
    smem_bytes = (size_t)nmat * (sizeof(Coef4<R>) + sizeof(R));
    tabsmem = smem_bytes <= 96 * 1024;
    if (!tabsmem) smem_bytes = 0;
    if (maxpoles) {
        // Debye poles give real coefficients (materials.py:102-108: w, q real), Lorentz / Drude complex ones.  If every entry
        // of the table is real, Im(T) stays zero for the whole run and T is held as a real array: half the traffic of the part
        // that dominates a dispersive half-step, the same E bits (GPB_DISP_COMPLEX keeps the complex form)
        const Cplx<R> *hc = (const Cplx<R> *)m.updatecoeffsdispersive;
        const size_t ncoef = (size_t)nmat * 3 * maxpoles;
        treal = !getenv("GPB_DISP_COMPLEX");
        for (size_t q = 0; q < ncoef && treal; ++q)
            if (hc[q].im != (R)0) treal = false;
        if (treal) {
            std::vector<R> re(ncoef);
            for (size_t q = 0; q < ncoef; ++q) re[q] = hc[q].re;
            R *d = nullptr;
            if (upload(&d, re.data(), ncoef)) return 1;
            CK(cudaStreamSynchronize(stream));
            dcoef = reinterpret_cast<Cplx<R> *>(d);
            for (int c = 0; c < 3; ++c) {
                R *t = nullptr;
                if (dalloc(&t, (size_t)narr * maxpoles)) return 1;
                T[c] = reinterpret_cast<Cplx<R> *>(t);
            }
        } else {
            for (int c = 0; c < 3; ++c)
                if (dalloc(&T[c], (size_t)narr * maxpoles)) return 1;
            if (upload(&dcoef, hc, ncoef)) return 1;
        }
    }
    if (dalloc(&d_iter, 2)) return 1;
    CK(cudaMalloc((void **)&d_flags, GPB_NFLAGS * sizeof(unsigned)));
    CK(cudaMemsetAsync(d_flags, 0, GPB_NFLAGS * sizeof(unsigned), stream));

    for (PhaseParams<R> *ph : {&ph_h, &ph_e}) {
        memset(ph, 0, sizeof *ph);
        ph->nx = nx; ph->ny = ny; ph->nz = nz; ph->x_start = x_start; ph->nplanes = nplanes;
        ph->pitch = pitch; ph->plane = plane;
        ph->Ex = F[0]; ph->Ey = F[1]; ph->Ez = F[2]; ph->Hx = F[3]; ph->Hy = F[4]; ph->Hz = F[5];
        ph->nmat = nmat; ph->form = form; ph->order = order;
        ph->p0 = 0; ph->p1 = nplanes;
    }
    for (int c = 0; c < 3; ++c) { ph_e.ID[c] = ID[c]; ph_h.ID[c] = ID[3 + c]; }
    ph_e.coef = coefE; ph_e.src = srcE; ph_h.coef = coefH; ph_h.src = srcH;
    ph_e.maxpoles = maxpoles; ph_e.dcoef = dcoef; ph_e.tstride = narr; ph_e.treal = treal ? 1 : 0;
    for (int c = 0; c < 3; ++c) ph_e.T[c] = T[c];
    set_boxes();
    tick("coefficient tables");
    if (setup_pml(m)) return 1;
    tick("PML slabs");
    // vectorised path: non-dispersive models (dispersive ones keep the generic scalar kernels)
    // vectorised path (dispersive E updates included; GPB_SCALAR forces the generic scalar kernels)
    use_v4 = v4_ok && !getenv("GPB_SCALAR");
    // small planes (2-D models, small 3-D grids): shorter marches so that there are enough blocks to fill the
    // GPU and the per-thread chain of dependent plane loads stays short
    {
        const long long bpp = (plane / 4 + kThreadsV4 - 1) / kThreadsV4;
        v4_xchunk = 16;
        while (v4_xchunk > 1 && bpp * ((nplanes + v4_xchunk - 1) / v4_xchunk) < 148ll * 8) v4_xchunk /= 2;
        if (getenv("GPB_V4_XCHUNK")) v4_xchunk = std::max(1, atoi(getenv("GPB_V4_XCHUNK")));
    }
    // TMA-staged path: 3-D grids with reasonably long z rows (2-D / thin grids stay on the flattened v4 path)
    // (below ~2.5 M nodes the pipeline fill of the TMA kernels costs more than it hides: 120^3 v4 22.7 vs TMA 20.2,
    //  150^3 v4 21.6 vs TMA 23.5 Gcells/s)
    use_tma = use_v4 && nz + 1 >= 32 && ny + 1 >= 8 && !getenv("GPB_NO_TMA") && (size_t)nmat * (sizeof(Coef4<R>) + sizeof(R)) <= 32 * 1024 &&
              ((long long)nplanes * plane >= 2500000ll || getenv("GPB_FORCE_TMA"));
    if (use_tma && setup_tma()) return 1;
    // dispersive E half-step on the TMA kernels: default tile, coefficient triples small enough for shared memory
    tma_disp = use_tma && maxpoles > 0 && tma_ty == 14 && tma_tz == 64 && tma_stages == 3 && tma_pw == 1 && !getenv("GPB_DISP_V4") &&
               (size_t)nmat * maxpoles * 3 * (treal ? 1 : 2) * sizeof(R) <= 24 * 1024;
    tma_tpf = getenv("GPB_TMA_TPF") ? std::max(0, atoi(getenv("GPB_TMA_TPF"))) : 2;
    // both half-steps in one launch (opt-in, GPB_PAIR=1): default tile with the producer warp, whole-domain handle (a shard's halo
    // protocol orders the half-steps itself), E half-step on the TMA kernels; models with something that acts on H between the two
    // half-steps (magnetic dipoles, transmission lines) keep the two launches (checked per step: has_hsrc).
    // Measured at 300^3 (profiles/README.md, r2 pair): DRAM traffic per iteration 2.36 -> 1.68 GB (E finds its operands in L2),
    // but 0.51 - 0.66 ms per iteration against 0.43 ms for two launches: the E items wait for H items that are still in flight
    // (close coupling) or lose the L2 reuse (loose coupling) -- 296 resident CTAs hold about as much data in flight as L2 keeps.
    pair_he = use_tma && tma_ty == 14 && tma_tz == 64 && tma_stages == 3 && tma_pw == 1 && tma_persist && nplanes == nx + 1 &&
              (!maxpoles || tma_disp) && getenv("GPB_PAIR") && !getenv("GPB_NO_PAIR");
    if (pair_he) {
        pair_xchunk = getenv("GPB_PAIR_XCHUNK") ? std::max(1, atoi(getenv("GPB_PAIR_XCHUNK"))) : 4;
        pair_lag = getenv("GPB_PAIR_LAG") ? std::max(0, atoi(getenv("GPB_PAIR_LAG"))) : 1;
        n_he_chunks = (nplanes + pair_xchunk - 1) / pair_xchunk;
        if (n_he_chunks >= (1 << 11)) pair_he = false;
    }
    if (pair_he && dalloc(&d_progress, (size_t)n_he_chunks + 1)) return 1;
    tick("tensor maps");
    if (use_v4) {
        for (int s = 0; s < ph_e.nslabs; ++s)
            if (ph_e.slab[s].axis == 2) zslabs_e |= 1u << s;
        for (int s = 0; s < ph_h.nslabs; ++s)
            if (ph_h.slab[s].axis == 2) zslabs_h |= 1u << s;
    }

    if (setup_coop()) return 1;

    memset(&pp, 0, sizeof pp);
    pp.x_start = x_start; pp.nplanes = nplanes; pp.ny = ny; pp.nz = nz; pp.pitch = pitch; pp.plane = plane;
    for (int c = 0; c < 6; ++c) { pp.F[c] = F[c]; pp.ID[c] = ID[c]; }
    pp.srcE = srcE; pp.srcH = srcH; pp.iter = d_iter; pp.iterations = iterations;
    pp.dx = (R)m.dx; pp.dy = (R)m.dy; pp.dz = (R)m.dz;
    if (setup_points(m)) return 1;
    // an unlinked slab cannot average snapshot cells with planes it does not own (gpb_link checks the linked case)
    snap_unlinked_ok = true;
    for (auto &sn : snaps)
        for (int i = 0; i < sn.nx; ++i) {
            const int gi = sn.xs + i * sn.dx;
            if (gi >= x_start && gi < x_start + nplanes && gi + sn.dx >= x_start + nplanes) snap_unlinked_ok = false;
        }
    if (set_smem_attributes()) return 1;
    CK(cudaStreamSynchronize(stream));
    tick("sources / receivers");
    return 0;
}

// ------------------------------------------------------------------------------------------ launches
template <typename R>
int Solver<R>::setup_tma()
{
    // tile shape: 4 cells per thread.  Default 14 x 64 cells = 7 consumer warps + a dedicated producer warp (measured fastest on
    // B200: 54.5 vs 50.7 Gcells/s at 300^3 for 16 x 64 with thread 0 of the first consumer warp producing); another shape
    // (producer = thread 0) only when it wastes >8 % fewer lanes on this grid
    const int cand[3][3] = {{14, 64, 1}, {32, 32, 0}, {8, 128, 0}};
    double best = 1e30;
    for (auto &c : cand) {
        const double waste = (double)((ny + 1 + c[0] - 1) / c[0] * c[0]) * ((pitch + c[1] - 1) / c[1] * c[1]) / ((double)(ny + 1) * (nz + 1));
        if (waste < best * 0.92) { best = waste; tma_ty = c[0]; tma_tz = c[1]; tma_pw = c[2]; }
    }
    if (getenv("GPB_TMA_PW") && atoi(getenv("GPB_TMA_PW")) == 0 && tma_pw) { tma_ty = 16; tma_tz = 64; tma_pw = 0; }
    if (getenv("GPB_TMA_TZ")) { tma_tz = atoi(getenv("GPB_TMA_TZ")); tma_ty = 1024 / tma_tz; tma_pw = 0; }
    if (getenv("GPB_TMA_TY")) tma_ty = atoi(getenv("GPB_TMA_TY"));
    tma_stages = getenv("GPB_TMA_STAGES") ? atoi(getenv("GPB_TMA_STAGES")) : 3;
    const CUtensorMapDataType fdt = sizeof(R) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    const CUtensorMapDataType idt = idbytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : (idbytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32);
    const int rows = ny + 1, planes = nplanes + 2, TY = tma_ty, TZ = tma_tz;
    // E phase: operands Hx,Hy,Hz (F[3..5]), own Ex,Ey,Ez (F[0..2]); H phase the other way round.  F and ID are single
    // allocations of six components each (build / upload_ids), so a triple is a 4-D tensor with component stride narr.
    for (int ph = 0; ph < 2; ++ph) {
        TmaMaps4 &mp = ph == 0 ? maps_h : maps_e;
        R *op = ph == 0 ? F[0] : F[3], *own = ph == 0 ? F[3] : F[0];
        void *ids = ph == 0 ? ID[3] : ID[0];
        if (make_map(&mp.op, op, fdt, sizeof(R), pitch, rows, planes, narr, TZ + 4, TY + 1, 3) ||
            make_map(&mp.opx, op, fdt, sizeof(R), pitch, rows, planes, narr, TZ + 4, TY + 1, 2) ||
            make_map(&mp.own, own, fdt, sizeof(R), pitch, rows, planes, narr, TZ, TY, 3) ||
            make_map(&mp.id, ids, idt, idbytes, pitch, rows, planes, narr, TZ, TY, 3))
            return 1;
    }
    CK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));   // (cudaGetDeviceProperties took 3 - 190 ms here)
    tma_persist = !getenv("GPB_TMA_NOPERSIST");
    tma_nofast = getenv("GPB_TMA_NOFAST") != nullptr;
    tma_nosplit = getenv("GPB_TMA_NOSPLIT") != nullptr;
    tma_zsplit = getenv("GPB_TMA_ZSPLIT") != nullptr;
    tma_znocoop = getenv("GPB_TMA_ZNOCOOP") != nullptr;
    tma_pf = getenv("GPB_TMA_PF") ? std::max(0, atoi(getenv("GPB_TMA_PF"))) : 2;
    if (dalloc(&d_sched, 2)) return 1;
    const long long tiles = (long long)((ny + 1 + TY - 1) / TY) * ((pitch + TZ - 1) / TZ);
    // planes per work item: 8 (every item pays one extra slot for its x-neighbour plane and a per-item set-up of the masks);
    // 4 only when there would otherwise be fewer than two items per resident CTA to balance (measured at 150^3 with the
    // persistent kernels: 2 planes 23.4, 4 planes 25.5, 8 planes 25.3 Gcells/s; 170^3: 33.1 / 33.9 for 4 / 8)
    tma_xchunk = 8;
    if (tiles * ((nplanes + 7) / 8) < 2ll * 2 * sm_count) tma_xchunk = 4;
    if (getenv("GPB_TMA_XCHUNK")) tma_xchunk = std::max(1, atoi(getenv("GPB_TMA_XCHUNK")));
    // work items travel as tile | chunk << 20 through the kernels' item ring
    if (tiles >= (1ll << 19) || (nplanes + tma_xchunk - 1) / tma_xchunk >= (1 << 11)) use_tma = false;
    return 0;
}

// E or H half-step of planes [p0, p1) on the TMA-staged kernels (gpb_tma_inst.cu)
template <typename R>
int Solver<R>::launch_tma(int phase, int p0, int p1, int peer_store)
{
    TmaLaunch<R> a;
    PhaseParams<R> &p = a.p;
    p = phase == 0 ? ph_h : ph_e;
    p.p0 = p0; p.p1 = p1; p.xchunk = tma_xchunk; p.persist = tma_persist ? 1 : 0;
    // planes on which a thread whose 4 cells are interior in (j,k) needs no mask / slab logic at all
    p.fast_i0 = std::max(p.box[0].lo[0], std::max(p.box[1].lo[0], p.box[2].lo[0]));
    p.fast_i1 = std::min(p.box[0].hi[0], std::min(p.box[1].hi[0], p.box[2].hi[0]));
    for (int s = 0; s < p.nslabs; ++s)
        if (p.slab[s].axis == 0) {
            if (p.slab[s].minus) p.fast_i0 = std::max(p.fast_i0, p.slab[s].hi[0]);
            else p.fast_i1 = std::min(p.fast_i1, p.slab[s].lo[0]);
        }
    if (tma_nofast) p.fast_i1 = p.fast_i0;
    p.zfused = tma_zsplit ? 0 : 1;
    p.znocoop = tma_znocoop ? 1 : 0;
    a.maps = phase == 0 ? &maps_h : &maps_e;
    a.phase = phase;
    a.ty = tma_ty; a.tz = tma_tz; a.stages = tma_stages; a.pw = tma_pw;
    a.idbytes = idbytes;
    a.pf_max = tma_pf;
    a.nosplit = tma_nosplit ? 1 : 0;
    a.disp = (phase == 1 && maxpoles) ? (treal ? 2 : 1) : 0;
    a.t_max = tma_tpf;
    a.sm_count = sm_count;
    a.sched = d_sched;
    a.stream = stream;
    a.pdl = pdl_on() && !peer_store ? 1 : 0;
    p.progress = nullptr;
    p.peer1 = p.peer2 = nullptr;
    p.peer_plane = -1;
    p.peer_flag = p.peer_counter = p.peer_flag2 = nullptr;
    p.peer_iter = d_iter;
    p.peer_add = 1;
    p.peer_need = 0;
    if (peer_store && phase == 0 && right.present) {          // Hy,Hz of my last plane -> right neighbour's ghost plane x_start-1
        p.peer1 = right.F + 4 * right.narr;
        p.peer2 = right.F + 5 * right.narr;
        p.peer_plane = nplanes;
        if (peer_store == 2) { p.peer_flag = right.flags + GPB_FLAG_H_READY; p.peer_counter = d_flags + GPB_FLAG_PUSH_COUNT; }
    } else if (peer_store && phase == 1 && left.present) {    // Ey,Ez of my first plane -> left neighbour's ghost plane x_end
        p.peer1 = left.F + 1 * left.narr + plane * (left.nplanes + 1);
        p.peer2 = left.F + 2 * left.narr + plane * (left.nplanes + 1);
        p.peer_plane = 1;
        if (peer_store >= 2) { p.peer_flag = left.flags + GPB_FLAG_E_READY; p.peer_counter = d_flags + GPB_FLAG_PUSH_COUNT + 1; }
        // the items of my first plane are also the last readers of my H ghost plane (unless the step prologue or a line reads
        // it, see enqueue_linked_step): the left neighbour may write the next H plane into it from now on
        if (peer_store == 3) p.peer_flag2 = left.flags + GPB_FLAG_H_FREE;
    }
    std::string err;
    const int pv = 2 * form + order - 1;
    int rc;
    switch (pv) {
    case 0: rc = tma_launch<R, 0>(a, &err); break;
    case 1: rc = tma_launch<R, 1>(a, &err); break;
    case 2: rc = tma_launch<R, 2>(a, &err); break;
    default: rc = tma_launch<R, 3>(a, &err); break;
    }
    if (rc) return fail("%s", err.c_str());
    ++launches;
    const unsigned zs = p.zfused ? 0u : (phase == 0 ? zslabs_h : zslabs_e);
    if (zs) {
        int cells = 0, planes = 0;
        for (int s = 0; s < p.nslabs; ++s)
            if ((zs >> s) & 1u) {
                cells = std::max(cells, (p.slab[s].hi[1] - p.slab[s].lo[1]) * (p.slab[s].hi[2] - p.slab[s].lo[2]));
                planes = std::max(planes, p.slab[s].hi[0] - p.slab[s].lo[0]);
            }
        dim3 g2((unsigned)((cells + 255) / 256), (unsigned)planes, (unsigned)__builtin_popcount(zs));
        if (idbytes == 1) k_pml_slabs<R, uint8_t><<<g2, 256, 0, stream>>>(p, phase, zs, p0, p1);
        else if (idbytes == 2) k_pml_slabs<R, uint16_t><<<g2, 256, 0, stream>>>(p, phase, zs, p0, p1);
        else k_pml_slabs<R, uint32_t><<<g2, 256, 0, stream>>>(p, phase, zs, p0, p1);
        CK(cudaGetLastError());
        ++launches;
    }
    return 0;
}

template <typename R>
template <typename IDT>
int Solver<R>::launch_h(int p0, int p1)
{
    PhaseParams<R> p = ph_h;
    p.p0 = p0; p.p1 = p1; p.xchunk = v4_xchunk;
    if (use_v4) {
        dim3 grid((unsigned)((plane / 4 + kThreadsV4 - 1) / kThreadsV4), (unsigned)((p1 - p0 + v4_xchunk - 1) / v4_xchunk));
        if (tabsmem) CK(launch_pdl(pdl_on(), k_update_h4<R, IDT, true>, grid, kThreadsV4, smem_bytes, stream, p));
        else CK(launch_pdl(pdl_on(), k_update_h4<R, IDT, false>, grid, kThreadsV4, 0, stream, p));
        CK(cudaGetLastError());
        ++launches;
        if (zslabs_h) {
            int cells = 0, planes = 0;
            for (int s = 0; s < p.nslabs; ++s)
                if ((zslabs_h >> s) & 1u) {
                    cells = std::max(cells, (p.slab[s].hi[1] - p.slab[s].lo[1]) * (p.slab[s].hi[2] - p.slab[s].lo[2]));
                    planes = std::max(planes, p.slab[s].hi[0] - p.slab[s].lo[0]);
                }
            dim3 g2((unsigned)((cells + 255) / 256), (unsigned)planes, (unsigned)__builtin_popcount(zslabs_h));
            CK(launch_pdl(pdl_on(), k_pml_slabs<R, IDT>, g2, 256, 0, stream, p, 0, zslabs_h, p0, p1));
            CK(cudaGetLastError());
            ++launches;
        }
        return 0;
    }
    dim3 grid((unsigned)((plane + kThreads - 1) / kThreads), (unsigned)((p1 - p0 + kXChunk - 1) / kXChunk));
    if (tabsmem) CK(launch_pdl(pdl_on(), k_update_h<R, IDT, true>, grid, kThreads, smem_bytes, stream, p));
    else CK(launch_pdl(pdl_on(), k_update_h<R, IDT, false>, grid, kThreads, 0, stream, p));
    CK(cudaGetLastError());
    ++launches;
    return 0;
}

template <typename R>
template <typename IDT>
int Solver<R>::launch_e(int p0, int p1)
{
    PhaseParams<R> p = ph_e;
    p.p0 = p0; p.p1 = p1; p.xchunk = v4_xchunk;
    if (use_v4) {
        dim3 grid((unsigned)((plane / 4 + kThreadsV4 - 1) / kThreadsV4), (unsigned)((p1 - p0 + v4_xchunk - 1) / v4_xchunk));
        if (maxpoles) {
            if (tabsmem) CK(launch_pdl(pdl_on(), k_update_e4<R, IDT, true, true>, grid, kThreadsV4, smem_bytes, stream, p));
            else CK(launch_pdl(pdl_on(), k_update_e4<R, IDT, false, true>, grid, kThreadsV4, 0, stream, p));
        } else {
            if (tabsmem) CK(launch_pdl(pdl_on(), k_update_e4<R, IDT, true, false>, grid, kThreadsV4, smem_bytes, stream, p));
            else CK(launch_pdl(pdl_on(), k_update_e4<R, IDT, false, false>, grid, kThreadsV4, 0, stream, p));
        }
        CK(cudaGetLastError());
        ++launches;
        if (zslabs_e) {
            int cells = 0, planes = 0;
            for (int s = 0; s < p.nslabs; ++s)
                if ((zslabs_e >> s) & 1u) {
                    cells = std::max(cells, (p.slab[s].hi[1] - p.slab[s].lo[1]) * (p.slab[s].hi[2] - p.slab[s].lo[2]));
                    planes = std::max(planes, p.slab[s].hi[0] - p.slab[s].lo[0]);
                }
            dim3 g2((unsigned)((cells + 255) / 256), (unsigned)planes, (unsigned)__builtin_popcount(zslabs_e));
            CK(launch_pdl(pdl_on(), k_pml_slabs<R, IDT>, g2, 256, 0, stream, p, 1, zslabs_e, p0, p1));
            CK(cudaGetLastError());
            ++launches;
        }
        return 0;
    }
    dim3 grid((unsigned)((plane + kThreads - 1) / kThreads), (unsigned)((p1 - p0 + kXChunk - 1) / kXChunk));
    if (maxpoles) {
        if (tabsmem) CK(launch_pdl(pdl_on(), k_update_e<R, IDT, true, true>, grid, kThreads, smem_bytes, stream, p));
        else CK(launch_pdl(pdl_on(), k_update_e<R, IDT, false, true>, grid, kThreads, 0, stream, p));
    } else {
        if (tabsmem) CK(launch_pdl(pdl_on(), k_update_e<R, IDT, true, false>, grid, kThreads, smem_bytes, stream, p));
        else CK(launch_pdl(pdl_on(), k_update_e<R, IDT, false, false>, grid, kThreads, 0, stream, p));
    }
    CK(cudaGetLastError());
    ++launches;
    return 0;
}

template <typename R>
int Solver<R>::launch_phase(int phase, int p0, int p1)
{
    if (p1 <= p0) return 0;
    if (use_tma && !(phase == 1 && maxpoles && !tma_disp)) return launch_tma(phase, p0, p1);
    if (phase == 0) {
        if (idbytes == 1) return launch_h<uint8_t>(p0, p1);
        if (idbytes == 2) return launch_h<uint16_t>(p0, p1);
        return launch_h<uint32_t>(p0, p1);
    }
    if (idbytes == 1) return launch_e<uint8_t>(p0, p1);
    if (idbytes == 2) return launch_e<uint16_t>(p0, p1);
    return launch_e<uint32_t>(p0, p1);
}

template <typename R>
int Solver<R>::launch_sources(int phase, int p0, int p1, int t0, int t1, bool begin_next)
{
    // point sources on the owned planes [p0, p1) and transmission lines on the owned planes [t0, t1) (local indices);
    // begin_next (electric phase, inside a multi-iteration graph): the next iteration's prologue rides in the same launch
    if (phase == 0 ? !has_hsrc : !has_esrc) return begin_next ? launch_begin() : 0;
    const int i_lo = x_start + p0, i_hi = x_start + p1, tl_lo = x_start + t0, tl_hi = x_start + t1;
    // (host copy of the source planes: no launch when nothing of this phase lies in the ranges)
    bool tl_any = false, src_any = false;
    for (int t = 0; t < ntl; ++t) tl_any = tl_any || (h_tls[t].i >= tl_lo && h_tls[t].i < tl_hi);
    for (size_t q = 0; q < h_src_plane.size(); ++q)
        src_any = src_any || (h_src_phase[q] == phase && h_src_plane[q] >= i_lo && h_src_plane[q] < i_hi);
    if (!src_any && !tl_any) return begin_next ? launch_begin() : 0;
    const int ntl_ = tl_any ? ntl : 0;
    if (begin_next) {
        if (idbytes == 1) CK(launch_pdl(pdl_on(), k_sources_begin<R, uint8_t>, 1, 128, 0, stream, pp, nsrc, d_srcs, ntl_, d_tls, i_lo, i_hi, tl_lo, tl_hi, d_iter, d_iter + 1, nrx, d_rxc, d_rxs, ntl));
        else if (idbytes == 2) CK(launch_pdl(pdl_on(), k_sources_begin<R, uint16_t>, 1, 128, 0, stream, pp, nsrc, d_srcs, ntl_, d_tls, i_lo, i_hi, tl_lo, tl_hi, d_iter, d_iter + 1, nrx, d_rxc, d_rxs, ntl));
        else CK(launch_pdl(pdl_on(), k_sources_begin<R, uint32_t>, 1, 128, 0, stream, pp, nsrc, d_srcs, ntl_, d_tls, i_lo, i_hi, tl_lo, tl_hi, d_iter, d_iter + 1, nrx, d_rxc, d_rxs, ntl));
        ++launches;
        return 0;
    }
    if (idbytes == 1) CK(launch_pdl(pdl_on(), k_sources<R, uint8_t>, 1, 32, 0, stream, pp, phase, nsrc, d_srcs, ntl_, d_tls, i_lo, i_hi, tl_lo, tl_hi));
    else if (idbytes == 2) CK(launch_pdl(pdl_on(), k_sources<R, uint16_t>, 1, 32, 0, stream, pp, phase, nsrc, d_srcs, ntl_, d_tls, i_lo, i_hi, tl_lo, tl_hi));
    else CK(launch_pdl(pdl_on(), k_sources<R, uint32_t>, 1, 32, 0, stream, pp, phase, nsrc, d_srcs, ntl_, d_tls, i_lo, i_hi, tl_lo, tl_hi));
    CK(cudaGetLastError());
    ++launches;
    return 0;
}

template <typename R>
int Solver<R>::launch_begin()
{
    CK(launch_pdl(pdl_on(), k_step_begin<R>, 1, 128, 0, stream, pp, d_iter, d_iter + 1, nrx, d_rxc, d_rxs, ntl, d_tls));
    CK(cudaGetLastError());
    ++launches;
    return 0;
}

template <typename R>
int Solver<R>::launch_snapshots()
{
    for (size_t n = 0; n < snaps.size(); ++n) {
        if (snaps[n].time != iteration + 1) continue;
        const long long cells = (long long)snaps[n].nx * snaps[n].ny * snaps[n].nz;
        const int blocks = (int)std::min<long long>((cells + 255) / 256, 148 * 8);
        k_snapshot<R><<<blocks, 256, 0, stream>>>(pp, snapdev[n]);
        CK(cudaGetLastError());
        ++launches;
    }
    return 0;
}

template <typename R>
int Solver<R>::enqueue_step(bool with_snap, bool with_begin, bool begin_next)
{
    // model_build_run.py:590-696 in order.  Inside the multi-iteration graph the prologue of every iteration but the first is
    // part of the previous iteration's last launch (with_begin = false there, begin_next = true on all but the last)
    if (with_begin && launch_begin()) return 1;
    if (with_snap && launch_snapshots()) return 1;
    if (pair_he && !has_hsrc) {
        if (launch_pair()) return 1;
        return launch_sources(1, 0, nplanes, 0, nplanes, begin_next);
    }
    if (launch_phase(0, 0, nplanes)) return 1;
    if (launch_sources(0, 0, nplanes, 0, nplanes)) return 1;
    if (launch_phase(1, 0, nplanes)) return 1;
    if (launch_sources(1, 0, nplanes, 0, nplanes, begin_next)) return 1;
    return 0;
}

// H and E half-steps of the whole grid in one launch of k_update_pair
template <typename R>
int Solver<R>::launch_pair()
{
    CK(cudaMemsetAsync(d_progress, 0, (size_t)n_he_chunks * sizeof(unsigned), stream));
    TmaLaunchPair<R> a;
    for (int phase = 0; phase < 2; ++phase) {
        PhaseParams<R> &p = phase == 0 ? a.ph : a.pe;
        p = phase == 0 ? ph_h : ph_e;
        p.p0 = 0; p.p1 = nplanes; p.xchunk = pair_xchunk; p.persist = 1;
        p.fast_i0 = std::max(p.box[0].lo[0], std::max(p.box[1].lo[0], p.box[2].lo[0]));
        p.fast_i1 = std::min(p.box[0].hi[0], std::min(p.box[1].hi[0], p.box[2].hi[0]));
        for (int s = 0; s < p.nslabs; ++s)
            if (p.slab[s].axis == 0) {
                if (p.slab[s].minus) p.fast_i0 = std::max(p.fast_i0, p.slab[s].hi[0]);
                else p.fast_i1 = std::min(p.fast_i1, p.slab[s].lo[0]);
            }
        if (tma_nofast) p.fast_i1 = p.fast_i0;
        p.zfused = 1;
        p.znocoop = tma_znocoop ? 1 : 0;
        p.progress = d_progress;
        p.peer1 = p.peer2 = nullptr;
        p.peer_plane = -1;
        p.peer_flag = p.peer_counter = p.peer_flag2 = nullptr;
        p.pair_lag = pair_lag;
        p.prog_flags = d_progress + n_he_chunks;
        p.prog_timeout_ns = 5000000000ull;
    }
    a.maps_h = &maps_h; a.maps_e = &maps_e;
    a.idbytes = idbytes; a.pf_max = tma_pf; a.t_max = tma_tpf;
    a.disp = maxpoles ? (treal ? 2 : 1) : 0;
    a.sm_count = sm_count; a.sched = d_sched; a.stream = stream;
    std::string err;
    const int pv = 2 * form + order - 1;
    int rc;
    switch (pv) {
    case 0: rc = tma_launch_pair<R, 0>(a, &err); break;
    case 1: rc = tma_launch_pair<R, 1>(a, &err); break;
    case 2: rc = tma_launch_pair<R, 2>(a, &err); break;
    default: rc = tma_launch_pair<R, 3>(a, &err); break;
    }
    if (rc) return fail("%s", err.c_str());
    ++launches;
    return 0;
}

// One iteration of a LINKED shard (gpb_link): same kernels on the same operands as enqueue_step, boundary plane first, with the
// halo planes pushed into the neighbours' ghost planes by k_halo_push between the boundary and the interior launch, and
// one-thread flag waits where a phase needs the neighbour's plane.  Flag protocol (values = iteration + 1, monotonic):
//   my H_READY  <- left neighbour:  its Hy,Hz of this iteration are in my ghost plane x_start-1     (before my E half-step)
//   my E_READY  <- right neighbour: its Ey,Ez of the previous iteration are in my ghost plane x_end (before my H half-step)
//   my H_FREE   <- right neighbour: it has finished reading the H ghost plane I filled last iteration (before I refill it)
// The Ey,Ez ghost needs no such credit: the right neighbour pushes E(it) only after it has seen my H(it), which I send after
// the only kernel that reads that ghost plane (the H update of my last plane).
template <typename R>
int Solver<R>::enqueue_linked_step(bool with_snap)
{
    const int n = nplanes;
    const int *it = d_iter;
    if (launch_begin()) return 1;
    // H_FREE credit for the left neighbour.  Readers of my H ghost plane: the E update of my first plane, and -- only if they
    // exist -- receivers on my first plane (currents Ix, Iy, Iz in the step prologue) and transmission lines there.  Without
    // those the E half-step kernel publishes the credit itself together with E_READY, as soon as the items of the first plane
    // are done (loose coupling); otherwise it is published here, after the prologue of the next iteration.
    const bool ghost_read_late = rx_on_first_plane || tl_on_first_plane();
    const bool fused_all = use_tma && !tma_zsplit && !link_nofused && !(maxpoles && !tma_disp) && !link_late;
    // (opt-in, GPB_EARLY_CREDIT=1: with three slabs on one device a run timed out once in the tests, and the effect at 8 GPUs is
    //  not measured yet -- by default the credit is published after the step prologue)
    const bool credit_early = link_early_credit && left.present && fused_all && !ghost_read_late && !credit_src_on_first_plane();
    if (left.present && !credit_early) {
        k_flag_signal<<<1, 1, 0, stream>>>(left.flags + GPB_FLAG_H_FREE, it, 0);
        ++launches;
    }
    if (with_snap && snapshot_due(iteration)) {
        // snapshot cells on my last planes average with planes of the right neighbour, read straight from its memory: it must
        // have finished the previous iteration and must not start this one before I am done
        if (left.present) { k_flag_signal<<<1, 1, 0, stream>>>(left.flags + GPB_FLAG_SNAP_READY, it, 1); ++launches; }
        if (right.present) { k_flag_wait<<<1, 1, 0, stream>>>(d_flags + GPB_FLAG_SNAP_READY, it, 1, d_flags, GPB_FLAG_SNAP_READY, link_timeout_ns); ++launches; }
        if (launch_snapshots()) return 1;
        if (right.present) { k_flag_signal<<<1, 1, 0, stream>>>(right.flags + GPB_FLAG_SNAP_DONE, it, 1); ++launches; }
        if (left.present) { k_flag_wait<<<1, 1, 0, stream>>>(d_flags + GPB_FLAG_SNAP_DONE, it, 1, d_flags, GPB_FLAG_SNAP_DONE, link_timeout_ns); ++launches; }
    }
    // ---- H half-step in ONE launch.  On the TMA kernels the halo exchange is part of it: the threads that update the last
    // owned plane store Hy,Hz into the right neighbour's ghost plane as well (PhaseParams::peer1/2).  Other kernel families (small
    // shards, z slabs in their own kernel) and planes touched by a point source afterwards are pushed by k_halo_push.  The flag is
    // published when the half-step is complete -- the neighbour needs the plane only at the start of ITS next half-step, which
    // begins when mine ends.  (Round 1 and the first linked version split every half-step into a boundary and an interior launch
    // to send early: at ~16 us of fill and tail per TMA launch the split cost more than the 30 us transfer it hid.)
    auto src_on_plane = [&](int phase, int gi) {
        for (size_t q = 0; q < h_src_plane.size(); ++q)
            if (h_src_phase[q] == phase && h_src_plane[q] == gi) return true;
        for (int t = 0; t < ntl; ++t)
            if (phase == 1 && h_tls[t].i == gi) return true;
        return false;
    };
    const bool fused = use_tma && !tma_zsplit && !link_nofused;
    const bool fused_e = fused && !(maxpoles && !tma_disp);   // (a dispersive E half-step on the register kernel has no peer stores)
    if (right.present) {
        k_flag_wait<<<1, 1, 0, stream>>>(d_flags + GPB_FLAG_E_READY, it, 0, d_flags, GPB_FLAG_E_READY, link_timeout_ns);
        k_flag_wait<<<1, 1, 0, stream>>>(d_flags + GPB_FLAG_H_FREE, it, 0, d_flags, GPB_FLAG_H_FREE, link_timeout_ns);
        launches += 2;
    }
    // (the kernel itself announces the plane as soon as its items are done, unless a point source changes it afterwards)
    const bool early_h = fused && right.present && !src_on_plane(0, x_start + n - 1) && !link_late;
    if (fused && right.present) {
        if (launch_tma(0, 0, n, early_h ? 2 : 1)) return 1;
    } else if (launch_phase(0, 0, n)) return 1;
    if (launch_sources(0, 0, n, 0, 0)) return 1;
    if (right.present && !early_h) {
        if (!fused || src_on_plane(0, x_start + n - 1)) {
            k_halo_push<R><<<128, 256, 0, stream>>>(F[4] + plane * n, F[5] + plane * n, right.F + 4 * right.narr, right.F + 5 * right.narr, plane,
                                                     right.flags + GPB_FLAG_H_READY, it, 1, d_flags + GPB_FLAG_PUSH_COUNT);
        } else {
            k_flag_signal<<<1, 1, 0, stream>>>(right.flags + GPB_FLAG_H_READY, it, 1);
        }
        CK(cudaGetLastError());
        ++launches;
    }
    // ---- E half-step
    if (left.present) {
        k_flag_wait<<<1, 1, 0, stream>>>(d_flags + GPB_FLAG_H_READY, it, 1, d_flags, GPB_FLAG_H_READY, link_timeout_ns);
        ++launches;
    }
    // transmission-line currents after the wait: a line on my first plane reads H of the ghost plane (sources.py:444-452)
    if (ntl && launch_sources(0, 0, 0, 0, n)) return 1;
    const bool early_e = fused_e && left.present && !src_on_plane(1, x_start) && !link_late;
    if (fused_e && left.present) {
        if (launch_tma(1, 0, n, early_e ? (credit_early ? 3 : 2) : 1)) return 1;
    } else if (launch_phase(1, 0, n)) return 1;
    if (launch_sources(1, 0, n, 0, n)) return 1;
    if (left.present && !early_e) {
        if (!fused_e || src_on_plane(1, x_start)) {
            k_halo_push<R><<<128, 256, 0, stream>>>(F[1] + plane, F[2] + plane, left.F + 1 * left.narr + plane * (left.nplanes + 1),
                                                     left.F + 2 * left.narr + plane * (left.nplanes + 1), plane, left.flags + GPB_FLAG_E_READY, it, 1,
                                                     d_flags + GPB_FLAG_PUSH_COUNT + 1);
        } else {
            k_flag_signal<<<1, 1, 0, stream>>>(left.flags + GPB_FLAG_E_READY, it, 1);
        }
        CK(cudaGetLastError());
        ++launches;
    }
    CK(cudaGetLastError());
    return 0;
}

template <typename R>
int Solver<R>::check_link_timeout()
{
    if (!linked) return 0;
    unsigned t = 0;
    CK(cudaMemcpyAsync(&t, d_flags + GPB_FLAG_TIMEOUT, sizeof t, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (t) return fail("linked shard [%d, %d): a halo flag wait timed out (mask 0x%x: bit 0 H_READY, 1 E_READY, 2 H_FREE, 3 SNAP_READY, 4 SNAP_DONE) -- "
                       "a neighbouring shard stopped or was never started", x_start, x_start + nplanes, t);
    return 0;
}

template <typename R>
int Solver<R>::link_info(gpb_link_t *out)
{
    CK(cudaSetDevice(device));
    memset(out, 0, sizeof *out);
    out->process_id = (uint64_t)getpid();
    out->device_id = device;
    out->dtype = sizeof(R) == 4 ? GPB_F32 : GPB_F64;
    out->x_start = x_start; out->nx_planes = nplanes; out->ny = ny; out->nz = nz;
    out->plane_elems = (uint64_t)plane; out->array_elems = (uint64_t)narr;
    out->fields_ptr = (uint64_t)(uintptr_t)F[0];
    out->flags_ptr = (uint64_t)(uintptr_t)d_flags;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "gpb_link_t carries 64-byte IPC handles");
    // (a process that only ever links shards of its own does not need the handles: failure to export is not an error here)
    if (ipc_export(F[0], out->fields_ipc, &out->fields_ipc_offset) || ipc_export(d_flags, out->flags_ipc, &out->flags_ipc_offset)) {
        memset(out->fields_ipc, 0, 64);
        memset(out->flags_ipc, 0, 64);
    }
    return 0;
}

template <typename R>
int Solver<R>::unlink()
{
    if (!linked && !left.present && !right.present && !left.ipc_F && !left.ipc_flags && !right.ipc_F && !right.ipc_flags) return 0;
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    for (Peer *p : {&left, &right}) {
        if (p->ipc_F) g_ipc.release(p->ipc_F);
        if (p->ipc_flags) g_ipc.release(p->ipc_flags);
        *p = Peer();
    }
    drop_graphs();
    for (int c = 0; c < 6; ++c) pp.Fr[c] = nullptr;
    linked = false;
    return 0;
}

template <typename R>
int Solver<R>::link(const gpb_link_t *l, const gpb_link_t *r)
{
    CK(cudaSetDevice(device));
    unlink();
    if (!l && !r) return 0;
    if (const char *e = getenv("GPB_LINK_TIMEOUT_MS")) link_timeout_ns = (unsigned long long)std::max(1, atoi(e)) * 1000000ull;
    link_nofused = getenv("GPB_NO_FUSED_PUSH") != nullptr;
    link_late = getenv("GPB_LATE_SIGNAL") != nullptr;
    link_early_credit = getenv("GPB_EARLY_CREDIT") != nullptr;
    const uint64_t me = (uint64_t)getpid();
    for (int side = 0; side < 2; ++side) {
        const gpb_link_t *q = side == 0 ? l : r;
        if (!q) continue;
        Peer &p = side == 0 ? left : right;
        if (q->dtype != (sizeof(R) == 4 ? GPB_F32 : GPB_F64) || q->ny != ny || q->nz != nz || (long long)q->plane_elems != plane)
            return fail("neighbour shard has another grid or float type");
        if (side == 0 ? (q->x_start + q->nx_planes != x_start) : (q->x_start != x_start + nplanes))
            return fail("%s neighbour owns planes [%d, %d), which do not adjoin mine [%d, %d)", side == 0 ? "left" : "right", q->x_start, q->x_start + q->nx_planes,
                        x_start, x_start + nplanes);
        if (q->process_id == me) {
            if (q->device_id != device) {
                int can = 0;
                CK(cudaDeviceCanAccessPeer(&can, device, q->device_id));
                if (!can) return fail("device %d cannot access the memory of device %d (no peer path)", device, q->device_id);
                cudaError_t e = cudaDeviceEnablePeerAccess(q->device_id, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail("cudaDeviceEnablePeerAccess(%d) failed: %s", q->device_id, cudaGetErrorString(e));
                cudaGetLastError();
            }
            p.F = (R *)(uintptr_t)q->fields_ptr;
            p.flags = (unsigned *)(uintptr_t)q->flags_ptr;
        } else {
            static const unsigned char none[64] = {0};
            if (!memcmp(q->fields_ipc, none, 64)) return fail("the neighbour could not export its arrays through CUDA IPC");
            CK(g_ipc.acquire(q->fields_ipc, &p.ipc_F));
            CK(g_ipc.acquire(q->flags_ipc, &p.ipc_flags));
            p.F = (R *)((char *)p.ipc_F + q->fields_ipc_offset);
            p.flags = (unsigned *)((char *)p.ipc_flags + q->flags_ipc_offset);
        }
        p.narr = (long long)q->array_elems;
        p.x_start = q->x_start;
        p.nplanes = q->nx_planes;
        p.present = true;
    }
    // snapshots: cells of my planes average with plane gi + dx, which may belong to the right neighbour
    snap_needs_right = false;
    for (auto &sn : snaps)
        for (int i = 0; i < sn.nx; ++i) {
            const int gi = sn.xs + i * sn.dx;
            if (gi < x_start || gi >= x_start + nplanes || gi + sn.dx < x_start + nplanes) continue;
            if (!right.present || gi + sn.dx >= right.x_start + right.nplanes)
                return fail("snapshot plane %d + %d lies beyond the right neighbour's planes: shards thinner than the snapshot stride are not supported", gi, sn.dx);
            snap_needs_right = true;
        }
    for (int c = 0; c < 6; ++c) pp.Fr[c] = right.present ? right.F + (size_t)c * right.narr : nullptr;
    pp.xr_start = right.x_start;
    pp.xr_planes = right.nplanes;
    CK(cudaMemsetAsync(d_flags, 0, GPB_NFLAGS * sizeof(unsigned), stream));
    CK(cudaStreamSynchronize(stream));
    linked = true;
    return 0;
}


// Opt in to more than 48 KB of dynamic shared memory for the kernels that stage the coefficient rows there (models with more
// than ~2450 materials in float32).  Called once from build(): run(), half_step() and profile() all launch these kernels.
template <typename R>
int Solver<R>::set_smem_attributes()
{
    if (!(tabsmem && smem_bytes > 48 * 1024)) return 0;
    const int n = (int)smem_bytes;
#define GPB_OPTIN(k) CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, n))
#define GPB_OPTIN_IDT(IDT)                      \
    GPB_OPTIN((k_update_h4<R, IDT, true>));        \
    GPB_OPTIN((k_update_e4<R, IDT, true, false>)); \
    GPB_OPTIN((k_update_e4<R, IDT, true, true>));  \
    GPB_OPTIN((k_update_h<R, IDT, true>));         \
    GPB_OPTIN((k_update_e<R, IDT, true, false>));  \
    GPB_OPTIN((k_update_e<R, IDT, true, true>))
    GPB_OPTIN_IDT(uint8_t);
    GPB_OPTIN_IDT(uint16_t);
    GPB_OPTIN_IDT(uint32_t);
#undef GPB_OPTIN_IDT
#undef GPB_OPTIN
    return 0;
}

template <typename R>
const void *Solver<R>::coop_kernel() const
{
    if (idbytes == 1) return maxpoles ? (const void *)k_run_coop<R, uint8_t, true> : (const void *)k_run_coop<R, uint8_t, false>;
    if (idbytes == 2) return maxpoles ? (const void *)k_run_coop<R, uint16_t, true> : (const void *)k_run_coop<R, uint16_t, false>;
    return maxpoles ? (const void *)k_run_coop<R, uint32_t, true> : (const void *)k_run_coop<R, uint32_t, false>;
}

// Grids that are launch-bound on the kernel-per-half-step path (everything the register-vectorised kernels serve: 2-D models
// and 3-D grids below the TMA threshold) run n iterations in one cooperative launch, provided the whole domain is on this
// handle (a shard's half-steps are ordered by the halo protocol) and the device can keep enough CTAs resident.
template <typename R>
int Solver<R>::setup_coop()
{
    coop = false;
    // Opt-in (GPB_COOP=1).  Measured on B200 (profiles/README.md, r2 coop): bit-identical, but slower than the graph of small
    // kernels -- cylinder_Ascan_2D 30 us per iteration against 18 us, 100^3 168 us against 52 us: a grid-wide barrier over ~600
    // resident CTAs costs more than the two kernel boundaries it replaces, and the half-step bodies, called as functions with
    // coherent loads, spill.
    if (!use_v4 || use_tma || nplanes != nx + 1 || !getenv("GPB_COOP") || getenv("GPB_NO_COOP")) return 0;
    int can = 0;
    CK(cudaDeviceGetAttribute(&can, cudaDevAttrCooperativeLaunch, device));
    if (!can) return 0;
    coop_smem = (size_t)nmat * 2 * (sizeof(Coef4<R>) + sizeof(R));
    if (coop_smem > 96 * 1024) return 0;
    const void *kern = coop_kernel();
    if (coop_smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)coop_smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreadsV4, coop_smem));
    CK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    if (per_sm < 1) return 0;
    coop_gx = (int)((plane / 4 + kThreadsV4 - 1) / kThreadsV4);
    const int resident = per_sm * sm_count;
    // planes per item: as long as possible (the x-neighbour plane rides in registers) while every resident CTA still gets a
    // few items per half-step
    coop_xchunk = 16;
    while (coop_xchunk > 1 && (long long)coop_gx * ((nplanes + coop_xchunk - 1) / coop_xchunk) < 3ll * resident) coop_xchunk /= 2;
    if (getenv("GPB_COOP_XCHUNK")) coop_xchunk = std::max(1, atoi(getenv("GPB_COOP_XCHUNK")));
    coop_gy = (nplanes + coop_xchunk - 1) / coop_xchunk;
    coop_grid = (int)std::min<long long>((long long)coop_gx * coop_gy, resident);
    coop = true;
    return 0;
}

template <typename R>
int Solver<R>::launch_coop(int n)
{
    CoopParams<R> cp;
    cp.ph = ph_h; cp.pe = ph_e;
    for (PhaseParams<R> *p : {&cp.ph, &cp.pe}) { p->p0 = 0; p->p1 = nplanes; p->xchunk = coop_xchunk; }
    cp.pp = pp;
    cp.nrx = nrx; cp.rxc = d_rxc; cp.rxs = d_rxs;
    cp.nsrc = nsrc; cp.srcs = d_srcs;
    cp.zs_h = zslabs_h; cp.zs_e = zslabs_e;
    cp.gx = coop_gx; cp.gy = coop_gy;
    cp.it0 = iteration; cp.n_iters = n;
    cp.d_iter = d_iter;
    void *args[] = {&cp};
    CK(cudaLaunchCooperativeKernel(coop_kernel(), dim3((unsigned)coop_grid), dim3(kThreadsV4), args, coop_smem, stream));
    ++launches;
    return 0;
}

template <typename R>
int Solver<R>::run(int n)
{
    if (begin_run(n) || enqueue_iterations(n) || end_run()) return 1;
    return finish_run();
}

// run() in four steps, so that a sharded run can interleave the slabs' launches (ShardedSolver::run)
template <typename R>
int Solver<R>::begin_run(int n)
{
    CK(cudaSetDevice(device));
    if (n < 0 || iteration + n > iterations) return fail("cannot run %d iterations from %d: model has %d", n, iteration, iterations);
    if (use_graph && n > 1 && !(coop && !ntl && !linked)) {
        // the step is identical every iteration (the iteration index lives on the device), so it is captured once and replayed:
        // one graph launch per time step instead of 4-6 kernel launches -- and, for unlinked solvers, a second graph that holds
        // graph_iters steps: launches inside a graph follow each other more closely than two graph launches do (measured with
        // 16 steps per graph: 2-D A-scan 18.4 -> 17.6, 100^3 53.3 -> 50.8, 300^3 436.5 -> 433.6 us per iteration)
        for (int k = 0; k < 2; ++k) {
            cudaGraphExec_t &target = k == 0 ? graph : graph_k;
            const int steps = k == 0 ? 1 : graph_iters;
            if (target || (k == 1 && (linked || graph_iters < 2 || n < graph_iters))) continue;
            cudaGraph_t g = nullptr;
            const uint64_t l0 = launches;
            CK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
            int rc = 0;
            for (int q = 0; q < steps && !rc; ++q)
                rc = linked ? enqueue_linked_step(false) : (k == 0 ? enqueue_step(false) : enqueue_step(false, q == 0 || !fuse_begin, fuse_begin && q + 1 < steps));
            cudaError_t e = cudaStreamEndCapture(stream, &g);
            (k == 0 ? graph_launches : graph_k_launches) = launches - l0;
            launches = l0;
            if (rc) { if (g) cudaGraphDestroy(g); return 1; }
            if (e != cudaSuccess) return fail("graph capture failed: %s", cudaGetErrorString(e));
            CK(cudaGraphInstantiate(&target, g, 0));
            cudaGraphDestroy(g);
        }
    }
    CK(cudaEventRecord(ev0, stream));
    return 0;
}

template <typename R>
int Solver<R>::enqueue_iterations(int n)
{
    CK(cudaSetDevice(device));
    if (coop && !ntl && !linked) {
        // cooperative whole-run kernel, split around the iterations that take a snapshot (those run as a plain step)
        while (n > 0) {
            if (snapshot_due(iteration)) {
                if (enqueue_step(true)) return 1;
                ++iteration;
                --n;
                continue;
            }
            int k = n;
            for (auto &sn : snaps)
                if (sn.time - 1 > iteration) k = std::min(k, sn.time - 1 - iteration);
            if (launch_coop(k)) return 1;
            iteration += k;
            n -= k;
        }
        return 0;
    }
    for (int s = 0; s < n; ++s) {
        if (graph_k && s + graph_iters <= n) {
            bool snap = false;
            for (int q = 0; q < graph_iters; ++q) snap = snap || snapshot_due(iteration + q);
            if (!snap) {
                CK(cudaGraphLaunch(graph_k, stream));
                launches += graph_k_launches;
                iteration += graph_iters;
                s += graph_iters - 1;
                continue;
            }
        }
        if (graph && !snapshot_due(iteration)) {
            CK(cudaGraphLaunch(graph, stream));
            launches += graph_launches;
        } else if (linked ? enqueue_linked_step(true) : enqueue_step(true)) {
            return 1;
        }
        ++iteration;
    }
    return 0;
}

template <typename R>
int Solver<R>::end_run()
{
    CK(cudaSetDevice(device));
    CK(cudaEventRecord(ev1, stream));
    return 0;
}

template <typename R>
int Solver<R>::finish_run()
{
    CK(cudaSetDevice(device));
    CK(cudaEventSynchronize(ev1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ev0, ev1));
    elapsed += ms * 1e-3;
    CK(cudaGetLastError());
    if (pair_he) {
        unsigned t = 0;
        CK(cudaMemcpy(&t, d_progress + n_he_chunks, sizeof t, cudaMemcpyDeviceToHost));
        if (t) return fail("k_update_pair: an E item waited more than 5 s for the H items it depends on (GPB_NO_PAIR=1 runs the half-steps as two launches)");
    }
    return check_link_timeout();
}

// Which kernel family each half-step runs on ("H:<kernel> E:<kernel>"), for logs and the bench line
template <typename R>
std::string Solver<R>::kernel_path() const
{
    if (coop && !ntl && !linked)
        return std::string("H+E:k_run_coop (whole run in one cooperative launch; ") + (maxpoles ? (treal ? "h4_body / e4_body<DISP=real>" : "h4_body / e4_body<DISP=complex>") : "h4_body / e4_body") + ")";
    auto name = [&](int phase) -> std::string {
        const char *dn = treal ? "DISP=real" : "DISP=complex";
        if (use_tma && !(phase == 1 && maxpoles && !tma_disp)) {
            char b[128];
            if (phase == 1 && maxpoles) snprintf(b, sizeof b, "k_update_tma<%dx%d,PHASE=1,%s>", tma_ty, tma_tz, dn);
            else snprintf(b, sizeof b, "k_update_tma<%dx%d,PHASE=%d>", tma_ty, tma_tz, phase);
            return std::string(b) + ((pair_he && !has_hsrc) ? " [one launch: k_update_pair]" : "");
        }
        if (use_v4) return phase == 0 ? "k_update_h4" : (maxpoles ? std::string("k_update_e4<") + dn + ">" : "k_update_e4");
        return phase == 0 ? "k_update_h" : (maxpoles ? std::string("k_update_e<") + dn + ">" : "k_update_e");
    };
    return "H:" + name(0) + " E:" + name(1);
}

// Per-kernel device times of `n` iterations launched WITHOUT the graph, CUDA events between the
// kernels on the launching stream: ms4 = {step prologue, H update, E update, source kernels} summed
// over the n iterations.  Used by bench.py for the roofline of the dominant kernel.
template <typename R>
int Solver<R>::profile(int n, double *ms4)
{
    CK(cudaSetDevice(device));
    if (n < 0 || iteration + n > iterations) return fail("cannot profile %d iterations from %d: model has %d", n, iteration, iterations);
    if (linked) return fail("gpb_profile works on unlinked handles only");
    cudaEvent_t ev[6];
    for (auto &e : ev) CK(cudaEventCreate(&e));
    // (profile() always times the half-step kernels one after the other; gpb_profile_step times the iteration as gpb_run runs it)
    for (int q = 0; q < 4; ++q) ms4[q] = 0;
    for (int s = 0; s < n; ++s) {
        CK(cudaEventRecord(ev[0], stream));
        if (launch_begin() || launch_snapshots()) return 1;
        CK(cudaEventRecord(ev[1], stream));
        if (launch_phase(0, 0, nplanes)) return 1;
        CK(cudaEventRecord(ev[2], stream));
        if (launch_sources(0, 0, nplanes, 0, nplanes)) return 1;
        CK(cudaEventRecord(ev[3], stream));
        if (launch_phase(1, 0, nplanes)) return 1;
        CK(cudaEventRecord(ev[4], stream));
        if (launch_sources(1, 0, nplanes, 0, nplanes)) return 1;
        CK(cudaEventRecord(ev[5], stream));
        CK(cudaEventSynchronize(ev[5]));
        float t01, t12, t23, t34, t45;
        CK(cudaEventElapsedTime(&t01, ev[0], ev[1]));
        CK(cudaEventElapsedTime(&t12, ev[1], ev[2]));
        CK(cudaEventElapsedTime(&t23, ev[2], ev[3]));
        CK(cudaEventElapsedTime(&t34, ev[3], ev[4]));
        CK(cudaEventElapsedTime(&t45, ev[4], ev[5]));
        ms4[0] += t01; ms4[1] += t12; ms4[2] += t34; ms4[3] += t23 + t45;
        ++iteration;
    }
    for (auto &e : ev) cudaEventDestroy(e);
    return 0;
}

template <typename R>
int Solver<R>::half_step(int phase, int part)
{
    // Host-driven sharded stepping (the caller moves the halo planes between the calls, e.g. over NCCL).  Transmission-line
    // currents are advanced at the start of phase 1, i.e. after the caller has received the H halo: a line on the first owned
    // plane reads H of the ghost plane (sources.py:444-452).
    CK(cudaSetDevice(device));
    if (linked) return fail("this shard is linked to its neighbours: advance it with gpb_run");
    host_driven = true;   // the caller orders these launches against its own transfers with events: plain launches from now on
    if (!snaps.empty() && !snap_unlinked_ok)
        return fail("a snapshot spans the cut plane at x = %d: link the shards (gpb_link / gpb_create_sharded), host-driven half-steps cannot reach the neighbour's planes",
                    x_start + nplanes);
    const bool first = part != 1, second = part != 0;
    if (phase == 0) {
        if (first) {
            if (iteration >= iterations) return fail("all %d iterations already done", iterations);
            if (launch_begin() || launch_snapshots()) return 1;
            // boundary plane first (its Hy,Hz feed the right neighbour) so the host can start the halo
            // transfer while the interior runs
            // (with the point sources that sit on it: the plane is sent as soon as this part is done)
            if (launch_phase(0, nplanes - 1, nplanes) || launch_sources(0, nplanes - 1, nplanes, 0, 0)) return 1;
        }
        if (second && (launch_phase(0, 0, nplanes - 1) || launch_sources(0, 0, nplanes - 1, 0, 0))) return 1;
    } else {
        // first owned plane: its Ey,Ez feed the left neighbour
        if (first && ((ntl && launch_sources(0, 0, 0, 0, nplanes)) || launch_phase(1, 0, 1) || launch_sources(1, 0, 1, 0, 1))) return 1;
        if (second) {
            if (launch_phase(1, 1, nplanes) || launch_sources(1, 1, nplanes, 1, nplanes)) return 1;
            ++iteration;
        }
    }
    return 0;
}

template <typename R>
int Solver<R>::reset()
{
    CK(cudaSetDevice(device));
    for (int c = 0; c < 6; ++c) CK(cudaMemsetAsync(F[c], 0, (size_t)narr * sizeof(R), stream));
    for (int c = 0; c < 3; ++c)
        if (T[c]) CK(cudaMemsetAsync(T[c], 0, (size_t)narr * maxpoles * (treal ? sizeof(R) : sizeof(Cplx<R>)), stream));
    for (auto &ph : phis) CK(cudaMemsetAsync(ph.first, 0, ph.second * sizeof(R), stream));
    CK(cudaMemsetAsync(d_rxs, 0, (size_t)GPB_NRXOUT * iterations * std::max(nrx, 1) * sizeof(R), stream));
    CK(cudaMemsetAsync(d_iter, 0, 2 * sizeof(int), stream));
    // (linked shards: the caller resets all of them before any runs again -- the flags count iterations)
    CK(cudaMemsetAsync(d_flags, 0, GPB_NFLAGS * sizeof(unsigned), stream));
    for (int t = 0; t < ntl; ++t) {
        CK(cudaMemcpyAsync(h_tls[t].voltage, tl_v0[t].data(), tl_v0[t].size() * sizeof(R), cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(h_tls[t].current, tl_c0[t].data(), tl_c0[t].size() * sizeof(R), cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(h_tls[t].abcv, &tl_abc0[2 * t], 2 * sizeof(R), cudaMemcpyHostToDevice, stream));
        CK(cudaMemsetAsync(h_tls[t].Vtotal, 0, (size_t)iterations * sizeof(R), stream));
        CK(cudaMemsetAsync(h_tls[t].Itotal, 0, (size_t)iterations * sizeof(R), stream));
    }
    CK(cudaStreamSynchronize(stream));
    iteration = 0;
    elapsed = 0;
    return 0;
}

template <typename R>
int Solver<R>::get_receivers(void *out, size_t bytes)
{
    CK(cudaSetDevice(device));
    const size_t need = (size_t)GPB_NRXOUT * iterations * nrx * sizeof(R);
    if (bytes != need) return fail("receiver buffer is %zu bytes, expected %zu", bytes, need);
    if (need) CK(cudaMemcpyAsync(out, d_rxs, need, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return 0;
}

template <typename R>
int Solver<R>::get_snapshot(int idx, void *out6[6], size_t bytes_each)
{
    CK(cudaSetDevice(device));
    if (idx < 0 || idx >= (int)snaps.size()) return fail("snapshot index %d out of range", idx);
    const size_t need = (size_t)snaps[idx].nx * snaps[idx].ny * snaps[idx].nz * sizeof(R);
    if (bytes_each != need) return fail("snapshot buffer is %zu bytes, expected %zu", bytes_each, need);
    for (int c = 0; c < 6; ++c) CK(cudaMemcpyAsync(out6[c], snapdev[idx].out[c], need, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return 0;
}

template <typename R>
int Solver<R>::get_tline(int idx, void *v, void *i, size_t bytes_each)
{
    CK(cudaSetDevice(device));
    if (idx < 0 || idx >= ntl) return fail("transmission line index %d out of range", idx);
    const size_t need = (size_t)iterations * sizeof(R);
    if (bytes_each != need) return fail("transmission line buffer is %zu bytes, expected %zu", bytes_each, need);
    CK(cudaMemcpyAsync(v, h_tls[idx].Vtotal, need, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(i, h_tls[idx].Itotal, need, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return 0;
}

template <typename R>
int Solver<R>::get_field(int comp, void *out, size_t bytes)
{
    CK(cudaSetDevice(device));
    if (comp < 0 || comp > 5) return fail("field component %d out of range", comp);
    const size_t row = (size_t)(nz + 1) * sizeof(R), rows = (size_t)nplanes * (ny + 1);
    if (bytes != row * rows) return fail("field buffer is %zu bytes, expected %zu", bytes, row * rows);
    CK(cudaMemcpy2DAsync(out, row, F[comp] + plane, (size_t)pitch * sizeof(R), row, rows, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return 0;
}

template <typename R>
int Solver<R>::set_field(int comp, const void *in, size_t bytes)
{
    CK(cudaSetDevice(device));
    if (comp < 0 || comp > 5) return fail("field component %d out of range", comp);
    const size_t row = (size_t)(nz + 1) * sizeof(R), rows = (size_t)nplanes * (ny + 1);
    if (bytes != row * rows) return fail("field buffer is %zu bytes, expected %zu", bytes, row * rows);
    CK(cudaMemcpy2DAsync(F[comp] + plane, (size_t)pitch * sizeof(R), in, row, row, rows, cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream));
    return 0;
}

template <typename R>
int Solver<R>::halo(int which, void **a, void **b, size_t *bytes)
{
    // E halo: Ey,Ez of the first owned plane go to the left neighbour's ghost plane (x_start+nplanes there);
    // H halo: Hy,Hz of the last owned plane go to the right neighbour's ghost plane (x_start-1 there).
    const long long first = plane, last = plane * nplanes, ghost_lo = 0, ghost_hi = plane * (nplanes + 1);
    switch (which) {
    case 0: *a = F[1] + first; *b = F[2] + first; break;
    case 1: *a = F[1] + ghost_hi; *b = F[2] + ghost_hi; break;
    case 2: *a = F[4] + last; *b = F[5] + last; break;
    case 3: *a = F[4] + ghost_lo; *b = F[5] + ghost_lo; break;
    default: return fail("halo selector %d out of range", which);
    }
    *bytes = (size_t)plane * sizeof(R);
    return 0;
}

// ------------------------------------------------------------------------------------------ one domain over several devices
// gpb_create_sharded: the nx + 1 node planes are cut into contiguous x-slabs, one Solver per device, neighbouring slabs linked
// over peer memory (Solver::link).  Every slab advances with its own CUDA graph on its own stream; the only coupling is the
// flag-announced halo pushes, so the host just replays one graph per device and iteration.  The reference has no
// counterpart: it rejects a model that does not fit one GPU (grid.py:239-241) and keeps only gpus[0] (gprMax.py:141-144).
template <typename R>
struct ShardedSolver : SolverBase {
    std::vector<Solver<R> *> sh;
    int nx = 0, ny = 0, nz = 0, iterations = 0, nrx = 0, ntl = 0;
    std::vector<gpb_snapshot_t> snaps;

    ~ShardedSolver()
    {
        for (auto *s : sh)
            if (s) { cudaSetDevice(s->device); if (s->stream) cudaStreamSynchronize(s->stream); }
        for (auto *s : sh)
            if (s) s->unlink();
        for (auto *s : sh) delete s;
    }

    int build(const gpb_model_t &m, const int *devices, int ndev)
    {
        nx = m.nx; ny = m.ny; nz = m.nz; iterations = m.iterations; nrx = m.nrx; ntl = m.ntlines;
        snaps.assign(m.snapshots, m.snapshots + m.nsnapshots);
        if (m.x_start != 0 || m.nx_planes != m.nx + 1) return fail("gpb_create_sharded takes the global model (x_start 0, nx_planes nx + 1)");
        if (ndev < 1 || ndev > m.nx + 1) return fail("cannot cut %d planes into %d slabs", m.nx + 1, ndev);
        sh.assign(ndev, nullptr);
        // balanced contiguous ranges of the nx + 1 node planes (the first slabs take the remainder)
        std::vector<int> x0(ndev + 1, 0);
        for (int d = 0; d < ndev; ++d) x0[d + 1] = x0[d] + (m.nx + 1) / ndev + (d < (m.nx + 1) % ndev ? 1 : 0);
        const size_t gplane = (size_t)(m.ny + 1) * (m.nz + 1);
        std::vector<std::string> errs(ndev);
        std::vector<int> rcs(ndev, 0);
        std::vector<std::thread> th;
        for (int d = 0; d < ndev; ++d) {
            sh[d] = new Solver<R>();
            sh[d]->device = devices[d];
            th.emplace_back([&, d] {
                gpb_model_t ms = m;
                ms.x_start = x0[d];
                ms.nx_planes = x0[d + 1] - x0[d];
                if (m.ID) {   // slab by slab straight out of the caller's array
                    ms.ID = m.ID + (size_t)x0[d] * gplane;
                    ms.id_comp_stride = m.id_comp_stride > 0 ? m.id_comp_stride : (int64_t)((size_t)(m.nx + 1) * gplane);
                }
                rcs[d] = sh[d]->build(ms);
                if (rcs[d]) errs[d] = g_err;
            });
        }
        for (auto &t : th) t.join();
        for (int d = 0; d < ndev; ++d)
            if (rcs[d]) return fail("slab %d (device %d, planes [%d, %d)): %s", d, devices[d], x0[d], x0[d + 1], errs[d].c_str());
        std::vector<gpb_link_t> info(ndev);
        for (int d = 0; d < ndev; ++d)
            if (sh[d]->link_info(&info[d])) return 1;
        if (ndev > 1)
            for (int d = 0; d < ndev; ++d)
                if (sh[d]->link(d > 0 ? &info[d - 1] : nullptr, d + 1 < ndev ? &info[d + 1] : nullptr)) return 1;
        for (auto *s : sh) mem += s->mem;
        device = devices[0];
        stream = sh[0]->stream;
        return 0;
    }

    int run(int n) override
    {
        if (n < 0 || iteration + n > iterations) return fail("cannot run %d iterations from %d: model has %d", n, iteration, iterations);
        // the slabs' iterations are enqueued interleaved, a few at a time: the flags order the slabs among themselves on the
        // devices, but a slab whose launches all sat in front of its neighbour's in a full launch queue would wait for work the
        // host could not submit yet
        for (auto *s : sh)
            if (s->begin_run(n)) return 1;
        for (int done = 0; done < n; done += 16)
            for (auto *s : sh)
                if (s->enqueue_iterations(std::min(16, n - done))) return 1;
        for (auto *s : sh)
            if (s->end_run()) return 1;
        double worst = 0;
        int rc = 0;
        for (auto *s : sh) {
            const double e0 = s->elapsed;
            if (s->finish_run()) rc = 1;
            worst = std::max(worst, s->elapsed - e0);
        }
        if (rc) return 1;
        elapsed += worst;
        iteration += n;
        launches = 0;
        for (auto *s : sh) launches += s->launches;
        return 0;
    }
    int half_step(int, int) override { return fail("a sharded handle advances with gpb_run"); }
    int profile(int, double *) override { return fail("gpb_profile works on single-device handles only"); }
    int halo(int, void **, void **, size_t *) override { return fail("a sharded handle exchanges its halo planes itself"); }
    int link_info(gpb_link_t *) override { return fail("a sharded handle is already linked internally"); }
    int link(const gpb_link_t *, const gpb_link_t *) override { return fail("a sharded handle is already linked internally"); }
    std::string kernel_path() const override
    {
        char b[64];
        snprintf(b, sizeof b, "%d x-slabs, each ", (int)sh.size());
        return b + sh[0]->kernel_path();
    }
    int reset() override
    {
        for (auto *s : sh)
            if (s->reset()) return 1;
        iteration = 0;
        elapsed = 0;
        return 0;
    }
    int set_points(const gpb_model_t &m) override
    {
        for (auto *s : sh)
            if (s->set_points(m)) return 1;
        nrx = m.nrx; ntl = m.ntlines;
        snaps.assign(m.snapshots, m.snapshots + m.nsnapshots);
        iteration = 0;
        elapsed = 0;
        return 0;
    }
    // every receiver / transmission line / snapshot cell is owned by exactly one slab and zero in the others: the sum is exact
    int sum_into(std::vector<R> &acc, const std::vector<R> &part)
    {
        for (size_t q = 0; q < acc.size(); ++q) acc[q] += part[q];
        return 0;
    }
    int get_receivers(void *out, size_t bytes) override
    {
        const size_t n = (size_t)GPB_NRXOUT * iterations * nrx;
        if (bytes != n * sizeof(R)) return fail("receiver buffer is %zu bytes, expected %zu", bytes, n * sizeof(R));
        std::vector<R> acc(n, (R)0), part(n);
        for (auto *s : sh) {
            if (s->get_receivers(part.data(), bytes)) return 1;
            sum_into(acc, part);
        }
        memcpy(out, acc.data(), bytes);
        return 0;
    }
    int get_snapshot(int idx, void *out6[6], size_t bytes_each) override
    {
        if (idx < 0 || idx >= (int)snaps.size()) return fail("snapshot index %d out of range", idx);
        const size_t n = (size_t)snaps[idx].nx * snaps[idx].ny * snaps[idx].nz;
        if (bytes_each != n * sizeof(R)) return fail("snapshot buffer is %zu bytes, expected %zu", bytes_each, n * sizeof(R));
        std::vector<std::vector<R>> acc(6, std::vector<R>(n, (R)0)), part(6, std::vector<R>(n));
        void *pp6[6];
        for (int c = 0; c < 6; ++c) pp6[c] = part[c].data();
        for (auto *s : sh) {
            if (s->get_snapshot(idx, pp6, bytes_each)) return 1;
            for (int c = 0; c < 6; ++c) sum_into(acc[c], part[c]);
        }
        for (int c = 0; c < 6; ++c) memcpy(out6[c], acc[c].data(), bytes_each);
        return 0;
    }
    int get_tline(int idx, void *v, void *i, size_t bytes_each) override
    {
        if (idx < 0 || idx >= ntl) return fail("transmission line index %d out of range", idx);
        const size_t n = (size_t)iterations;
        if (bytes_each != n * sizeof(R)) return fail("transmission line buffer is %zu bytes, expected %zu", bytes_each, n * sizeof(R));
        std::vector<R> av(n, (R)0), ai(n, (R)0), pv(n), pi(n);
        for (auto *s : sh) {
            if (s->get_tline(idx, pv.data(), pi.data(), bytes_each)) return 1;
            sum_into(av, pv);
            sum_into(ai, pi);
        }
        memcpy(v, av.data(), bytes_each);
        memcpy(i, ai.data(), bytes_each);
        return 0;
    }
    int get_field(int comp, void *out, size_t bytes) override
    {
        const size_t gplane = (size_t)(ny + 1) * (nz + 1) * sizeof(R);
        if (bytes != gplane * (nx + 1)) return fail("field buffer is %zu bytes, expected %zu", bytes, gplane * (nx + 1));
        for (auto *s : sh)
            if (s->get_field(comp, (char *)out + gplane * s->x_start, gplane * s->nplanes)) return 1;
        return 0;
    }
    int set_field(int comp, const void *in, size_t bytes) override
    {
        const size_t gplane = (size_t)(ny + 1) * (nz + 1) * sizeof(R);
        if (bytes != gplane * (nx + 1)) return fail("field buffer is %zu bytes, expected %zu", bytes, gplane * (nx + 1));
        for (auto *s : sh)
            if (s->set_field(comp, (const char *)in + gplane * s->x_start, gplane * s->nplanes)) return 1;
        return 0;
    }
};

// ------------------------------------------------------------------------------------------ C ABI
extern "C" {

#define NEED(h) if (!(h) || !(h)->impl) return fail("null handle")

int gpb_release_cached(void)
{
    g_ipc.close_idle();
    g_pool.release_all();
    return 0;
}

const char *gpb_last_error(void) { return g_err.c_str(); }
const char *gpb_version(void) { return "gprmax_b200 0.1 (sm_100a)"; }

int gpb_device_count(int *count)
{
    if (!count) return fail("null argument");
    ScopedFullAffinity aff;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        cudaGetLastError();
        return fail("no CUDA device available: %s", cudaGetErrorString(e));
    }
    *count = n;
    return 0;
}

int gpb_device_info(int device_id, gpb_device_info_t *out)
{
    if (!out) return fail("null argument");
    ScopedFullAffinity aff;
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, device_id));
    memset(out, 0, sizeof *out);
    out->device_id = device_id;
    strncpy(out->name, pr.name, sizeof out->name - 1);
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device_id) == cudaSuccess) strncpy(out->pci_bus_id, bus, sizeof out->pci_bus_id - 1);
    out->total_mem = pr.totalGlobalMem;
    out->const_mem = pr.totalConstMem;
    out->sm_count = pr.multiProcessorCount;
    out->cc_major = pr.major;
    out->cc_minor = pr.minor;
    return 0;
}

int gpb_create(const gpb_model_t *model, int device_id, gpb_handle *out)
{
    if (!model || !out) return fail("null argument");
    ScopedFullAffinity aff;
    *out = nullptr;
    if (model->abi_version != GPB_ABI_VERSION) return fail("ABI version mismatch: library %d, caller %d", GPB_ABI_VERSION, model->abi_version);
    int n = 0;
    if (gpb_device_count(&n)) return 1;
    {   // a failed call of the application (or of another library in this thread) must not surface in this library's checks
        cudaError_t stale = cudaGetLastError();
        if (stale != cudaSuccess && getenv("GPB_DEBUG")) fprintf(stderr, "[gpb] stale CUDA error at gpb_create: %s\n", cudaGetErrorString(stale));
    }
    if (device_id < 0 || device_id >= n) return fail("GPU with device ID %d does not exist (%d device(s) present)", device_id, n);
    SolverBase *s = nullptr;
    int rc;
    if (model->dtype == GPB_F32) {
        auto *p = new Solver<float>();
        p->device = device_id;
        rc = p->build(*model);
        s = p;
    } else if (model->dtype == GPB_F64) {
        auto *p = new Solver<double>();
        p->device = device_id;
        rc = p->build(*model);
        s = p;
    } else {
        return fail("unknown dtype %d", model->dtype);
    }
    if (rc) {
        std::string keep = g_err;
        delete s;
        cudaGetLastError();
        g_err = keep;
        return 1;
    }
    *out = new gpb_solver{s};
    return 0;
}

int gpb_create_sharded(const gpb_model_t *model, const int *device_ids, int ndevices, gpb_handle *out)
{
    if (!model || !out || !device_ids) return fail("null argument");
    ScopedFullAffinity aff;
    *out = nullptr;
    if (model->abi_version != GPB_ABI_VERSION) return fail("ABI version mismatch: library %d, caller %d", GPB_ABI_VERSION, model->abi_version);
    int n = 0;
    if (gpb_device_count(&n)) return 1;
    for (int d = 0; d < ndevices; ++d) {
        if (device_ids[d] < 0 || device_ids[d] >= n) return fail("GPU with device ID %d does not exist (%d device(s) present)", device_ids[d], n);
        for (int e = 0; e < d; ++e)
            if (device_ids[e] == device_ids[d] && !getenv("GPB_ALLOW_SAME_DEVICE")) return fail("device %d is listed twice", device_ids[d]);
    }
    SolverBase *s = nullptr;
    int rc;
    if (model->dtype == GPB_F32) {
        auto *p = new ShardedSolver<float>();
        rc = p->build(*model, device_ids, ndevices);
        s = p;
    } else if (model->dtype == GPB_F64) {
        auto *p = new ShardedSolver<double>();
        rc = p->build(*model, device_ids, ndevices);
        s = p;
    } else {
        return fail("unknown dtype %d", model->dtype);
    }
    if (rc) {
        std::string keep = g_err;
        delete s;
        cudaGetLastError();
        g_err = keep;
        return 1;
    }
    *out = new gpb_solver{s};
    return 0;
}

int gpb_set_points(gpb_handle h, const gpb_model_t *model)
{
    NEED(h);
    if (!model) return fail("null argument");
    if (model->abi_version != GPB_ABI_VERSION) return fail("ABI version mismatch: library %d, caller %d", GPB_ABI_VERSION, model->abi_version);
    return h->impl->set_points(*model);
}
int gpb_link_info(gpb_handle h, gpb_link_t *out) { NEED(h); if (!out) return fail("null argument"); return h->impl->link_info(out); }
int gpb_link(gpb_handle h, const gpb_link_t *left, const gpb_link_t *right) { NEED(h); return h->impl->link(left, right); }

int gpb_destroy(gpb_handle h)
{
    if (!h) return 0;
    delete h->impl;
    delete h;
    return 0;
}


int gpb_run(gpb_handle h, int n_iters) { NEED(h); return h->impl->run(n_iters); }
int gpb_half_step(gpb_handle h, int phase, int part) { NEED(h); if (phase != 0 && phase != 1) return fail("phase must be 0 or 1"); if (part < -1 || part > 1) return fail("part must be -1, 0 or 1"); return h->impl->half_step(phase, part); }
int gpb_reset(gpb_handle h) { NEED(h); return h->impl->reset(); }
int gpb_iteration(gpb_handle h, int *it) { NEED(h); if (!it) return fail("null argument"); *it = h->impl->iteration; return 0; }
int gpb_elapsed_seconds(gpb_handle h, double *s) { NEED(h); if (!s) return fail("null argument"); *s = h->impl->elapsed; return 0; }
int gpb_mem_used(gpb_handle h, uint64_t *b) { NEED(h); if (!b) return fail("null argument"); *b = h->impl->mem; return 0; }
int gpb_kernel_launches(gpb_handle h, uint64_t *c) { NEED(h); if (!c) return fail("null argument"); *c = h->impl->launches; return 0; }
int gpb_get_receivers(gpb_handle h, void *out, size_t bytes) { NEED(h); if (!out && bytes) return fail("null argument"); return h->impl->get_receivers(out, bytes); }
int gpb_get_snapshot(gpb_handle h, int idx, void *out6[6], size_t bytes_each) { NEED(h); if (!out6) return fail("null argument"); return h->impl->get_snapshot(idx, out6, bytes_each); }
int gpb_get_tline(gpb_handle h, int idx, void *v, void *i, size_t bytes_each) { NEED(h); if (!v || !i) return fail("null argument"); return h->impl->get_tline(idx, v, i, bytes_each); }
int gpb_get_field(gpb_handle h, int comp, void *out, size_t bytes) { NEED(h); if (!out) return fail("null argument"); return h->impl->get_field(comp, out, bytes); }
int gpb_set_field(gpb_handle h, int comp, const void *in, size_t bytes) { NEED(h); if (!in) return fail("null argument"); return h->impl->set_field(comp, in, bytes); }
int gpb_halo(gpb_handle h, int which, void **a, void **b, size_t *bytes) { NEED(h); if (!a || !b || !bytes) return fail("null argument"); return h->impl->halo(which, a, b, bytes); }
int gpb_profile(gpb_handle h, int n_iters, double *ms4) { NEED(h); if (!ms4) return fail("null argument"); return h->impl->profile(n_iters, ms4); }
int gpb_kernel_path(gpb_handle h, char *buf, size_t buf_bytes)
{
    NEED(h);
    if (!buf || !buf_bytes) return fail("null argument");
    const std::string s = h->impl->kernel_path();
    snprintf(buf, buf_bytes, "%s", s.c_str());
    return 0;
}
int gpb_stream(gpb_handle h, void **s) { NEED(h); if (!s) return fail("null argument"); *s = (void *)h->impl->stream; return 0; }
int gpb_synchronize(gpb_handle h)
{
    NEED(h);
    CK(cudaSetDevice(h->impl->device));
    CK(cudaStreamSynchronize(h->impl->stream));
    return 0;
}

}  // extern "C"
