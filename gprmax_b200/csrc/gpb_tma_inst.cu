// gpb_tma_inst.cu -- one translation unit per (GPB_TMA_R, GPB_TMA_PV): the TMA-staged E/H kernels of that float type
// and PML variant in every tile shape and ID width, plus their launcher.  Split like this so that (a) the PML
// formulation and order are compile-time constants inside the kernels and (b) the variants compile in parallel.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "gpb_kernels_pair.cuh"

#ifndef GPB_TMA_R
#define GPB_TMA_R float
#endif
#ifndef GPB_TMA_PV
#define GPB_TMA_PV 0
#endif

namespace gpb {

static int tma_fail(std::string *err, const char *what, cudaError_t e)
{
    char buf[512];
    snprintf(buf, sizeof buf, "%s failed: %s", what, cudaGetErrorString(e));
    if (err) *err = buf;
    return 1;
}

template <typename R, int PV, typename IDT, int TY, int TZ, int S, int PW, int DISP = 0>
static int tma_launch_cfg(const TmaLaunch<R> &a, std::string *err)
{
    using L = StageLayout<R, IDT, TY, TZ>;
    constexpr int order = (PV & 1) + 1;
    PhaseParams<R> p = a.p;
    const int tiles_k = (p.pitch + TZ - 1) / TZ, tiles_j = (p.ny + 1 + TY - 1) / TY;
    // shared memory: barriers | coefficient rows | PML R tables | per-thread Phi prefetch slots | stage ring
    p.tmax = 1;
    for (int s = 0; s < p.nslabs; ++s) p.tmax = std::max(p.tmax, p.slab[s].t);
    const size_t fixed = 128 + (size_t)((p.nmat * (sizeof(Coef4<R>) + sizeof(R)) + 127) / 128 * 128) +
                         (size_t)((p.nslabs * 4 * order * p.tmax * sizeof(R) + 127) / 128 * 128) + (size_t)S * L::bytes;
    const size_t pf_unit = (size_t)2 * order * (TY * TZ / 4) * 4 * sizeof(R);   // one plane of Phi per thread
    const size_t smem_cap = (sizeof(R) == 4 ? (size_t)(227 * 1024) / GPB_TMA_CTAS : (size_t)227 * 1024) - 1024;
    // dispersive E half-step: coefficient triples + T prefetch slots; T is touched by every cell, Phi only inside the slabs, so T
    // gets the first call on the shared memory that is left
    size_t disp_fixed = 0, t_unit = 0;
    p.t_depth = 0;
    if (DISP) {
        const int tw = DISP == 1 ? 2 : 1;
        disp_fixed = (size_t)((p.nmat * p.maxpoles * 3 * tw * sizeof(R) + 127) / 128 * 128);
        t_unit = (size_t)3 * p.maxpoles * tw * (TY * TZ / 4) * 4 * sizeof(R);
        if (fixed + disp_fixed > smem_cap) {
            if (err) *err = "dispersive coefficient table does not fit the shared memory of the TMA kernels";
            return 1;
        }
        p.t_depth = fixed + disp_fixed + 2 * t_unit <= smem_cap ? 2 : (fixed + disp_fixed + t_unit <= smem_cap ? 1 : 0);
        p.t_depth = std::min(p.t_depth, a.t_max);
    }
    const size_t fixed2 = fixed + disp_fixed + p.t_depth * t_unit;
    p.pf_depth = p.nslabs == 0 ? 0 : (fixed2 + 2 * pf_unit <= smem_cap ? 2 : (fixed2 + pf_unit <= smem_cap ? 1 : 0));
    p.pf_depth = std::min(p.pf_depth, a.pf_max);
    const size_t smem = fixed2 + p.pf_depth * pf_unit;
    const int tiles = tiles_k * tiles_j, nchunks = (p.p1 - p.p0 + p.xchunk - 1) / p.xchunk;
    p.peer_need = (unsigned)(tiles * (TY * TZ / 4 / 32));
    // persistent: as many CTAs as fit on the GPU at once (2 per SM for fp32), pulling (tile, chunk) items from a.sched
    const int resident = (sizeof(R) == 4 ? GPB_TMA_CTAS : 1) * a.sm_count;
    p.monotone = 0;
    p.progress = nullptr;
    // persistent launches hand the last ~wave of chunks out in halves (chunk_range)
    const int nsplit = (p.persist && p.xchunk >= 4 && !a.nosplit) ? std::min(nchunks, (resident + tiles - 1) / tiles + 1) : 0;
    const dim3 grid = p.persist ? dim3((unsigned)std::min(tiles * (nchunks + nsplit), resident)) : dim3((unsigned)tiles, (unsigned)nchunks);
    cudaError_t e;
    if (a.phase == 0) {
        auto kern = k_update_tma<R, IDT, TY, TZ, S, 0, PW, PV>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return tma_fail(err, "cudaFuncSetAttribute", e);
        if ((e = launch_pdl(a.pdl != 0, kern, grid, dim3(TY * TZ / 4 + 32 * PW), smem, a.stream, p, *a.maps, tiles_k, tiles, nchunks, nsplit, a.sched)) != cudaSuccess)
            return tma_fail(err, "k_update_tma launch", e);
    } else {
        auto kern = k_update_tma<R, IDT, TY, TZ, S, 1, PW, PV, DISP>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return tma_fail(err, "cudaFuncSetAttribute", e);
        if ((e = launch_pdl(a.pdl != 0, kern, grid, dim3(TY * TZ / 4 + 32 * PW), smem, a.stream, p, *a.maps, tiles_k, tiles, nchunks, nsplit, a.sched)) != cudaSuccess)
            return tma_fail(err, "k_update_tma launch", e);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return tma_fail(err, "k_update_tma launch", e);
    return 0;
}

template <typename R, int PV, typename IDT>
static int tma_launch_idt(const TmaLaunch<R> &a, std::string *err)
{
    // dispersive electric half-step (a.disp: 1 complex T, 2 real T): instantiated for the default tile only
    if (a.disp && a.phase == 1) {
        if (a.ty == 14 && a.tz == 64 && a.stages == 3 && a.pw == 1)
            return a.disp == 1 ? tma_launch_cfg<R, PV, IDT, 14, 64, 3, 1, 1>(a, err) : tma_launch_cfg<R, PV, IDT, 14, 64, 3, 1, 2>(a, err);
        if (err) *err = "the dispersive TMA kernels exist for the 14 x 64 tile only";
        return 1;
    }
#define GPB_TMA_CASE(TY_, TZ_, S_, PW_) \
    if (a.ty == TY_ && a.tz == TZ_ && a.stages == S_ && a.pw == PW_) return tma_launch_cfg<R, PV, IDT, TY_, TZ_, S_, PW_>(a, err)
    GPB_TMA_CASE(14, 64, 3, 1);
    GPB_TMA_CASE(16, 64, 3, 0);
    GPB_TMA_CASE(8, 128, 3, 0);
    GPB_TMA_CASE(32, 32, 3, 0);
#ifdef GPB_TMA_SWEEP   // ring-depth sweep variants of profiles/README.md (not built by default)
    GPB_TMA_CASE(14, 64, 4, 1);
    GPB_TMA_CASE(16, 64, 2, 0);
#endif
#undef GPB_TMA_CASE
    char buf[256];
    snprintf(buf, sizeof buf, "no TMA kernel instantiated for tile %d x %d with %d stages (producer warp %d)", a.ty, a.tz, a.stages, a.pw);
    if (err) *err = buf;
    return 1;
}

template <typename R, int PV>
int tma_launch(const TmaLaunch<R> &a, std::string *err)
{
    if (a.idbytes == 1) return tma_launch_idt<R, PV, uint8_t>(a, err);
    if (a.idbytes == 2) return tma_launch_idt<R, PV, uint16_t>(a, err);
    return tma_launch_idt<R, PV, uint32_t>(a, err);
}

template <typename R, int PV, typename IDT, int DISP>
static int tma_launch_pair_cfg(const TmaLaunchPair<R> &a, std::string *err)
{
    constexpr int TY = 14, TZ = 64, S = 3;
    using L = StageLayout<R, IDT, TY, TZ>;
    constexpr int order = (PV & 1) + 1;
    constexpr int threads = TY * TZ / 4;
    PhaseParams<R> ph = a.ph, pe = a.pe;
    const int tiles_k = (ph.pitch + TZ - 1) / TZ, tiles_j = (ph.ny + 1 + TY - 1) / TY;
    ph.tmax = pe.tmax = 1;
    for (int s = 0; s < ph.nslabs; ++s) ph.tmax = std::max(ph.tmax, ph.slab[s].t);
    for (int s = 0; s < pe.nslabs; ++s) pe.tmax = std::max(pe.tmax, pe.slab[s].t);
    auto a128 = [](size_t x) { return (x + 127) / 128 * 128; };
    const size_t fixed = 128 + 2 * a128(ph.nmat * (sizeof(Coef4<R>) + sizeof(R))) + a128(ph.nslabs * 4 * order * ph.tmax * sizeof(R)) +
                         a128(pe.nslabs * 4 * order * pe.tmax * sizeof(R)) + (size_t)S * L::bytes;
    const size_t pf_unit = (size_t)2 * order * threads * 4 * sizeof(R);
    const size_t smem_cap = (sizeof(R) == 4 ? (size_t)(227 * 1024) / GPB_TMA_CTAS : (size_t)227 * 1024) - 1024;
    size_t disp_fixed = 0, t_unit = 0;
    ph.t_depth = pe.t_depth = 0;
    if (DISP) {
        const int tw = DISP == 1 ? 2 : 1;
        disp_fixed = a128(pe.nmat * pe.maxpoles * 3 * tw * sizeof(R));
        t_unit = (size_t)3 * pe.maxpoles * tw * threads * 4 * sizeof(R);
        if (fixed + disp_fixed > smem_cap) {
            if (err) *err = "dispersive coefficient table does not fit the shared memory of the TMA kernels";
            return 1;
        }
        pe.t_depth = fixed + disp_fixed + 2 * t_unit <= smem_cap ? 2 : (fixed + disp_fixed + t_unit <= smem_cap ? 1 : 0);
        pe.t_depth = std::min(pe.t_depth, a.t_max);
    }
    const size_t fixed2 = fixed + disp_fixed + pe.t_depth * t_unit;
    const int nslabs = std::max(ph.nslabs, pe.nslabs);
    int pfd = nslabs == 0 ? 0 : (fixed2 + 2 * pf_unit <= smem_cap ? 2 : (fixed2 + pf_unit <= smem_cap ? 1 : 0));
    pfd = std::min(pfd, a.pf_max);
    ph.pf_depth = ph.nslabs ? pfd : 0;
    pe.pf_depth = pe.nslabs ? pfd : 0;
    const size_t smem = fixed2 + pfd * pf_unit;
    const int tiles = tiles_k * tiles_j, nchunks = (ph.p1 - ph.p0 + ph.xchunk - 1) / ph.xchunk;
    ph.monotone = pe.monotone = 1;
    ph.persist = pe.persist = 1;
    ph.prog_need = pe.prog_need = (unsigned)(tiles * (threads / 32));
    const int resident = (sizeof(R) == 4 ? GPB_TMA_CTAS : 1) * a.sm_count;
    const long long W = 2ll * nchunks * tiles;
    const dim3 grid((unsigned)std::min<long long>(W, resident));
    auto kern = k_update_pair<R, IDT, TY, TZ, S, PV, DISP>;
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return tma_fail(err, "cudaFuncSetAttribute", e);
    kern<<<grid, threads + 32, smem, a.stream>>>(ph, pe, *a.maps_h, *a.maps_e, tiles_k, tiles, nchunks, a.sched);
    if ((e = cudaGetLastError()) != cudaSuccess) return tma_fail(err, "k_update_pair launch", e);
    return 0;
}

template <typename R, int PV>
int tma_launch_pair(const TmaLaunchPair<R> &a, std::string *err)
{
#define GPB_PAIR_IDT(IDT)                                                      \
    do {                                                                       \
        if (a.disp == 0) return tma_launch_pair_cfg<R, PV, IDT, 0>(a, err);    \
        if (a.disp == 1) return tma_launch_pair_cfg<R, PV, IDT, 1>(a, err);    \
        return tma_launch_pair_cfg<R, PV, IDT, 2>(a, err);                     \
    } while (0)
    if (a.idbytes == 1) GPB_PAIR_IDT(uint8_t);
    if (a.idbytes == 2) GPB_PAIR_IDT(uint16_t);
    GPB_PAIR_IDT(uint32_t);
#undef GPB_PAIR_IDT
}

template int tma_launch<GPB_TMA_R, GPB_TMA_PV>(const TmaLaunch<GPB_TMA_R> &, std::string *);
template int tma_launch_pair<GPB_TMA_R, GPB_TMA_PV>(const TmaLaunchPair<GPB_TMA_R> &, std::string *);

}  // namespace gpb
