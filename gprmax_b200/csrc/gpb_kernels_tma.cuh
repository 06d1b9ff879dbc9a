// gpb_kernels_tma.cuh -- TMA-staged E/H half-step kernels (sm_100a), the bandwidth path.
//
// Why: the register-only vectorised kernels (gpb_kernels_v4.cuh) need ~120 registers per thread to
// keep one plane of operands in flight, which caps occupancy at 16 warps/SM and leaves them
// latency-bound at ~3.4 TB/s (profiles/r1b_*).  Here the operand planes are staged in shared memory
// by the Tensor Memory Accelerator instead:
//   * a CTA owns a TY x TZ tile of (j,k) and marches along x (E phase: +x, H phase: -x);
//   * for every plane a producer issues 3 `cp.async.bulk.tensor.4d` loads -- the operand triple with its one-row /
//     one-column halo, the triple being updated and the material-ID triple (each triple is one allocation, so the
//     component is the 4th tensor dimension) -- into one stage of a kStages-deep ring; completion is tracked with an
//     mbarrier (expect_tx / complete_tx); out-of-range halo coordinates are zero-filled by the TMA unit: no edge code;
//   * consumers read 128-bit rows from shared memory, so the j+-1 / k+-1 re-reads never touch L2, and
//     kStages planes (~26 KB each) are in flight per CTA regardless of register pressure;
//   * the x-neighbour plane (i-1 for E, i+1 for H) rides in a register queue as before;
//   * results go straight from registers to global memory with 128-bit stores;
//   * a consumer warp hands a stage back (one `empty` arrival per warp) only after EVERYTHING it reads from the stage --
//     operands, own fields, material IDs, and for slab warps the re-reads of the PML corrections -- has been read: the
//     producer refills the stage at once, and with persistent CTAs the refill may belong to another tile.
// The dispersive recursion of the E half-step (T read / written once by the owning thread, prefetched like Phi) and, for linked
// x-slab shards, the halo exchange (the boundary plane is stored into the neighbour's ghost plane by the threads that update it
// and announced by the warp that completes the last item holding it) are part of the same pass: see gpb_tma_item.inc.
// All PML slabs are applied in the same pass: x / y slabs vectorised and warp-uniform with Phi prefetched per thread by
// cp.async (`prefetch` in the kernel), z slabs one cell per lane with the corrections handed to the owning thread by shuffle.
#pragma once
#include <cuda.h>

#include "gpb_kernels_v4.cuh"
#include "gpb_tma.h"

namespace gpb {


__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Wait for the phase with the given parity.  try_wait carries a suspend-time hint: the waiting thread is parked by the hardware
// until the phase completes (or the hint expires) instead of spinning -- without it the spin loops of the producer warp and of
// consumers waiting for data were 16 % of all issued instructions of an issue-bound kernel (profiles/README.md, r1k).
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x989680u)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// per-thread asynchronous global -> shared copy of one 4-cell vector (Phi prefetch), L2 only
template <typename R>
__device__ __forceinline__ void cp_async_v4(void *sdst, const R *g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sdst)), "l"(g) : "memory");
    if (sizeof(R) == 8)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sdst) + 16), "l"(reinterpret_cast<const char *>(g) + 16) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// PML coefficient set of one depth from the shared-memory copy of a slab's R tables, tb = [RA|RB|RE|RF][order][tmax]
template <typename R>
__device__ __forceinline__ PmlCo<R> pml_load_s(int form, int order, const R *tb, int tmax, int depth)
{
    const R *RA = tb + depth, *RB = tb + order * tmax + depth, *RE = tb + 2 * order * tmax + depth, *RF = tb + 3 * order * tmax + depth;
    const int o1 = order == 2 ? tmax : 0;
    return pml_co(form, order, RA[0], RA[o1], RB[0], RB[o1], RE[0], RE[o1], RF[0], RF[o1]);
}

// shared-memory layout of one stage (byte offsets), every sub-buffer 128-byte aligned
template <typename R, typename IDT, int TY, int TZ>
struct StageLayout {
    static constexpr int a128(int x) { return (x + 127) / 128 * 128; }
    static constexpr int PA = TZ + 4;               // row pitch (elements) of the operand tiles
    static constexpr int CS = (TY + 1) * PA;        // elements per operand component
    static constexpr int OS = TY * TZ;              // elements per own / id component
    static constexpr int oOp = 0, szOp = a128(3 * CS * (int)sizeof(R));
    static constexpr int oOwn = oOp + szOp, szOwn = a128(3 * OS * (int)sizeof(R));
    static constexpr int oId = oOwn + szOwn, szId = a128(3 * OS * (int)sizeof(IDT));
    static constexpr int bytes = oId + szId;
    // bytes the TMA unit delivers per stage (full boxes, out-of-range parts are zero-filled)
    static constexpr int tx = 3 * CS * (int)sizeof(R) + 3 * OS * (int)sizeof(R) + 3 * OS * (int)sizeof(IDT);
    static constexpr int tx_x = 2 * CS * (int)sizeof(R);   // x-neighbour slot: operand components B and C (staged in the own buffer)
    static_assert(2 * CS <= 3 * OS, "x-neighbour tiles must fit the own buffer");
};

template <typename IDT>
__device__ __forceinline__ Ids4 lds_ids4(const unsigned char *base, int elem);
template <>
__device__ __forceinline__ Ids4 lds_ids4<uint8_t>(const unsigned char *base, int elem)
{
    const unsigned v = *reinterpret_cast<const unsigned *>(base + elem);
    return {v & 0xffu, (v >> 8) & 0xffu, (v >> 16) & 0xffu, v >> 24};
}
template <>
__device__ __forceinline__ Ids4 lds_ids4<uint16_t>(const unsigned char *base, int elem)
{
    const uint2 v = *reinterpret_cast<const uint2 *>(base + 2 * elem);
    return {v.x & 0xffffu, v.x >> 16, v.y & 0xffffu, v.y >> 16};
}
template <>
__device__ __forceinline__ Ids4 lds_ids4<uint32_t>(const unsigned char *base, int elem)
{
    const uint4 v = *reinterpret_cast<const uint4 *>(base + 4 * elem);
    return {v.x, v.y, v.z, v.w};
}

// The material IDs of a thread's 4 cells as loaded (packed): the common case "all four equal" is one multiply and one compare on
// the packed word, and only mixed quads are unpacked.
template <typename IDT>
struct IdQ;
template <>
struct IdQ<uint8_t> {
    unsigned v;
    __device__ __forceinline__ static IdQ load(const unsigned char *base, int elem) { return {*reinterpret_cast<const unsigned *>(base + elem)}; }
    __device__ __forceinline__ bool uniform() const { return v == (v & 0xffu) * 0x01010101u; }
    __device__ __forceinline__ unsigned at(int q) const { return (v >> (8 * q)) & 0xffu; }
};
template <>
struct IdQ<uint16_t> {
    uint2 v;
    __device__ __forceinline__ static IdQ load(const unsigned char *base, int elem) { return {*reinterpret_cast<const uint2 *>(base + 2 * elem)}; }
    __device__ __forceinline__ bool uniform() const { return v.x == v.y && (v.x & 0xffffu) == (v.x >> 16); }
    __device__ __forceinline__ unsigned at(int q) const { return ((q < 2 ? v.x : v.y) >> (16 * (q & 1))) & 0xffffu; }
};
template <>
struct IdQ<uint32_t> {
    uint4 v;
    __device__ __forceinline__ static IdQ load(const unsigned char *base, int elem) { return {*reinterpret_cast<const uint4 *>(base + 4 * elem)}; }
    __device__ __forceinline__ bool uniform() const { return v.x == v.y && v.x == v.z && v.x == v.w; }
    __device__ __forceinline__ unsigned at(int q) const { return q == 0 ? v.x : (q == 1 ? v.y : (q == 2 ? v.z : v.w)); }
};
template <typename R, typename IDT>
__device__ __forceinline__ void coef4q(const Coef4<R> *coef, const IdQ<IDT> &id, Coef4<R> &c0, Coef4<R> &c1, Coef4<R> &c2, Coef4<R> &c3)
{
    c0 = coef[id.at(0)];
    if (id.uniform()) {
        c1 = c2 = c3 = c0;
    } else {
        c1 = coef[id.at(1)];
        c2 = coef[id.at(2)];
        c3 = coef[id.at(3)];
    }
}

// Default order in which the x chunks are handed out (no sorted item list): both ends first, the middle last.  The chunks that
// cross the x slabs take 2-3 times as long as the others; handed out last they formed a final, half-empty wave of slow items.
__device__ __forceinline__ int chunk_of(int q, int nchunks)
{
    return (q & 1) ? nchunks - 1 - (q >> 1) : (q >> 1);
}
// Plane range [l0, l1) of hand-out slot v.  The last `nsplit` chunks of the hand-out order (about one wave of items) are
// handed out as two halves each, so that the SMs run dry within half an item of each other at the end of the kernel.
__device__ __forceinline__ void chunk_range(int v, int nchunks, int nsplit, int xc, int p0, int p1, int &l0, int &l1, int monotone = 0)
{
    const int nbig = nchunks - nsplit;
    int c, lo = 0, len = xc;
    if (monotone) {
        c = v;   // increasing x (k_update_pair)
    } else if (v < nbig) {
        c = chunk_of(v, nchunks);
    } else {
        const int u = v - nbig;
        c = chunk_of(nbig + (u >> 1), nchunks);
        lo = (u & 1) * (xc >> 1);
        len = (u & 1) ? xc - (xc >> 1) : (xc >> 1);
    }
    l0 = min(p0 + c * xc + lo, p1);
    l1 = min(l0 + len, p1);
}

// slabs whose x range holds plane i (CTA-uniform)
template <typename R>
__device__ __forceinline__ unsigned slabs_on_plane(const PhaseParams<R> &p, int i)
{
    unsigned act = 0;
#pragma unroll
    for (int s = 0; s < kMaxSlabs; ++s)
        if (s < p.nslabs && i >= p.slab[s].lo[0] && i < p.slab[s].hi[0]) act |= 1u << s;
    return act;
}

// One PML component of one slab on the thread's 4 cells, Phi included: load (prefetch slot or global), apply, store.
// Deliberately one component at a time with few values live: the slab path shares its register allocation with the
// straight-line update of every thread (two inlined variants that held both components' Phi spilled the hot loop:
// 47 -> 15 Gcells/s at 300^3; real calls spilled the register queue around them).
template <typename R, int DSTEP>
__device__ __forceinline__ void pml_comp(int form, int order, const R *tb, int tmax, int depth0, int dsign, R inv_d, unsigned m, const Ids4 &id,
                                         const R *src, R sign, const V4<R> &dF, V4<R> &F, R *phi, long long ostride2, const V4<R> *slot, int slot_stride2)
{
    V4<R> P0, P1;
    if (slot) {
        P0 = slot[0];
        P1 = order == 2 ? slot[slot_stride2] : P0;
    } else {
        P0 = ld4(phi);
        P1 = order == 2 ? ld4(phi + ostride2) : P0;
    }
    if (DSTEP == 0) {
        const PmlCo<R> co = pml_load_s(form, order, tb, tmax, depth0);
        if (m & 1u) F.x = fma_(sign, mul_(src[id.a], pml_apply(form, order, co, mul_(dF.x, inv_d), P0.x, P1.x)), F.x);
        if (m & 2u) F.y = fma_(sign, mul_(src[id.b], pml_apply(form, order, co, mul_(dF.y, inv_d), P0.y, P1.y)), F.y);
        if (m & 4u) F.z = fma_(sign, mul_(src[id.c], pml_apply(form, order, co, mul_(dF.z, inv_d), P0.z, P1.z)), F.z);
        if (m & 8u) F.w = fma_(sign, mul_(src[id.d], pml_apply(form, order, co, mul_(dF.w, inv_d), P0.w, P1.w)), F.w);
    } else {
        if (m & 1u) { const PmlCo<R> co = pml_load_s(form, order, tb, tmax, depth0); F.x = fma_(sign, mul_(src[id.a], pml_apply(form, order, co, mul_(dF.x, inv_d), P0.x, P1.x)), F.x); }
        if (m & 2u) { const PmlCo<R> co = pml_load_s(form, order, tb, tmax, depth0 + dsign); F.y = fma_(sign, mul_(src[id.b], pml_apply(form, order, co, mul_(dF.y, inv_d), P0.y, P1.y)), F.y); }
        if (m & 4u) { const PmlCo<R> co = pml_load_s(form, order, tb, tmax, depth0 + 2 * dsign); F.z = fma_(sign, mul_(src[id.c], pml_apply(form, order, co, mul_(dF.z, inv_d), P0.z, P1.z)), F.z); }
        if (m & 8u) { const PmlCo<R> co = pml_load_s(form, order, tb, tmax, depth0 + 3 * dsign); F.w = fma_(sign, mul_(src[id.d], pml_apply(form, order, co, mul_(dF.w, inv_d), P0.w, P1.w)), F.w); }
    }
    st4(phi, P0);
    if (order == 2) st4(phi + ostride2, P1);
}

// ------------------------------------------------------------------------------------------
// PHASE 1: electric half-step (marches +x, operands H, queue = Hy,Hz of plane i-1)
// PHASE 0: magnetic half-step (marches -x, operands E, queue = Ey,Ez of plane i+1)
// Arithmetic identical to k_update_e4 / k_update_h4.
// Shared memory: [kStages full + kStages empty mbarriers | item ring][coefficient rows][kStages stages]
//
// Work items = (tile, x-chunk) pairs.  Non-persistent launch: one CTA per item.  Persistent launch (p.persist):
// GPB_TMA_CTAS CTAs per SM pull items from an atomic counter and keep ONE continuous TMA pipeline running across item
// boundaries, so the ring never drains between marches.  Every item starts with a slot that carries only the two
// x-neighbour operand tiles (register-queue initialisation); a final empty slot carries the end-of-work signal.
// sched[0] = next item, sched[1] = finished CTAs (the last one resets both for the next launch).
//
// Producer (3 TMA instructions per plane: the operand, own and ID triples are 4-D boxes).  PW = 1: a dedicated warp --
// one lane walks the CTA's slot sequence and waits only for the stages' `empty` mbarriers (TY = 14: 7 consumer warps +
// the producer = 8 warps at 128 registers; a ninth warp would cap the kernel at 96 registers and spill, measured
// 40.2 vs 47.0 Gcells/s).  PW = 0: thread 0 of the first consumer warp refills right after its warp released a stage,
// before its own arithmetic.  (r1g ncu source page of the first design, where thread 0 issued 9 loads per plane
// after its arithmetic: 24 % of all stall samples sat on the `full` wait of the other seven warps.)
// ------------------------------------------------------------------------------------------
#ifndef GPB_TMA_CTAS
#define GPB_TMA_CTAS 2
#endif
template <typename R, typename IDT, int TY, int TZ, int kStages, int PHASE, int PW, int PV, int DISP = 0>
__global__ void __launch_bounds__(TY * TZ / 4 + 32 * PW, (sizeof(R) == 4 ? GPB_TMA_CTAS : 1))
k_update_tma(const PhaseParams<R> p, const __grid_constant__ TmaMaps4 maps, int tiles_k, int tiles, int nchunks, int nsplit, int *sched)
{
    constexpr int kTmaThreads = TY * TZ / 4;  // consumer threads: every thread owns 4 consecutive z cells of the tile
    static_assert(kTmaThreads % 32 == 0 && kTmaThreads + 32 * PW <= 256, "tile shape");
    // PML formulation and order are compile-time (PV = 2 * form + order - 1): with both at run time the four inlined variants of
    // every correction spilled ~500 bytes of the straight-line update of every thread (47 -> 15 Gcells/s at 300^3)
    constexpr int PFORM = PV >> 1, PORDER = (PV & 1) + 1;
    // DISP (electric half-step only): 0 no dispersive media, 1 complex T, 2 real T (Debye media) -- the polarisation arrays are
    // read and written once per step by the thread that owns the cells, prefetched like Phi (per-thread cp.async, p.t_depth)
    static_assert(DISP == 0 || PHASE == 1, "dispersive update belongs to the electric half-step");
    constexpr int KTW = DISP == 1 ? 2 : 1;   // V4 vectors per 4 cells of T
    using L = StageLayout<R, IDT, TY, TZ>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);   // [kStages] TMA bytes landed
    uint64_t *empty = full + kStages;                           // [kStages] all consumer warps have read the stage
    volatile int *ring = reinterpret_cast<volatile int *>(empty + kStages);  // [4] item ids, producer -> consumers
    Coef4<R> *scoef = reinterpret_cast<Coef4<R> *>(smem_raw + 128);
    R *ssrc = reinterpret_cast<R *>(scoef + p.nmat);
    const int coef_bytes = (int)((p.nmat * (sizeof(Coef4<R>) + sizeof(R)) + 127) / 128 * 128);
    R *stab = reinterpret_cast<R *>(smem_raw + 128 + coef_bytes);   // PML R tables [nslabs][RA|RB|RE|RF][order][tmax]
    const int tab_bytes = (int)((p.nslabs * 4 * PORDER * p.tmax * sizeof(R) + 127) / 128 * 128);
    V4<R> *spf = reinterpret_cast<V4<R> *>(smem_raw + 128 + coef_bytes + tab_bytes);   // Phi prefetch [pf_depth][2*order][threads]
    const int pf_bytes = p.pf_depth * 2 * PORDER * kTmaThreads * (int)sizeof(V4<R>);
    // dispersive: coefficient triples [nmat][poles][3] (R or complex) and the T prefetch slots [t_depth][3 comps][poles][TW][threads]
    R *sdc = reinterpret_cast<R *>(smem_raw + 128 + coef_bytes + tab_bytes + pf_bytes);
    const int dc_bytes = DISP ? (int)((p.nmat * p.maxpoles * 3 * KTW * sizeof(R) + 127) / 128 * 128) : 0;
    V4<R> *stf = reinterpret_cast<V4<R> *>(smem_raw + 128 + coef_bytes + tab_bytes + pf_bytes + dc_bytes);
    const int tslot = DISP ? 3 * p.maxpoles * KTW * kTmaThreads : 0;   // V4 vectors per plane of T slots
    const int tf_bytes = DISP ? p.t_depth * tslot * (int)sizeof(V4<R>) : 0;
    unsigned char *stages = smem_raw + 128 + coef_bytes + tab_bytes + pf_bytes + dc_bytes + tf_bytes;

    const int tid = threadIdx.x, lane = tid & 31;
    const int W = tiles * (nchunks + nsplit);

    pdl_launch_dependents();
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, kTmaThreads / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid == 32) {
        tma_prefetch_desc(&maps.op);
        tma_prefetch_desc(&maps.opx);
        tma_prefetch_desc(&maps.own);
        tma_prefetch_desc(&maps.id);
    }
    for (int m = tid; m < p.nmat; m += kTmaThreads + 32 * PW) {
        scoef[m] = p.coef[m];
        ssrc[m] = p.src[m];
    }
    if (DISP) {
        const R *src = reinterpret_cast<const R *>(p.dcoef);
        for (int m = tid; m < p.nmat * p.maxpoles * 3 * KTW; m += kTmaThreads + 32 * PW) sdc[m] = src[m];
    }
    for (int s = 0; s < p.nslabs; ++s) {
        const SlabDev<R> &sl = p.slab[s];
        for (int m = tid; m < 4 * PORDER * sl.t; m += kTmaThreads + 32 * PW) {
            const int q = m / (PORDER * sl.t), o = (m / sl.t) % PORDER, dd = m % sl.t;
            const R *src = q == 0 ? sl.RA : (q == 1 ? sl.RB : (q == 2 ? sl.RE : sl.RF));
            stab[((s * 4 + q) * PORDER + o) * p.tmax + dd] = src[o * sl.t + dd];
        }
    }
    __syncthreads();
    // everything above reads tables that are constant during a run; from here on the kernel touches what its predecessor in the
    // stream wrote (the work queue it re-armed, fields, Phi, T)
    pdl_wait();

    // ---------------- producer: one slot of the CTA's slot sequence per call (3 TMA instructions per plane)
    int p_item = -1, p_n = -1, p_q = 0, p_g = 0;   // item being loaded, its next plane (-1 = x-neighbour slot), item ordinal, slot
    int p_j0 = 0, p_k0 = 0, p_l0 = 0, p_l1 = 0;
    bool p_done = false;
    auto fetch = [&]() {
        int w;
        if (p.persist) w = atomicAdd(sched, 1);
        else w = p_q == 0 ? (int)(blockIdx.y * gridDim.x + blockIdx.x) : W;
        // x chunks from both ends inwards (chunk_of), tiles in row-major order: neighbouring tiles run at the same time and share
        // their halo rows / columns in L2.  (Handing the items out by estimated cost, most expensive first, to shorten the tail --
        // 7 % of the SM cycles are idle at the end -- broke that locality: 57.5 -> 53.8 Gcells/s at 300^3, 75.2 -> 66.8 at 500^3.)
        p_item = w < W ? ((w % tiles) | ((w / tiles) << 20)) : -1;
        if (p_item >= 0) {
            const int tile = p_item & 0x7ffff, chunk = p_item >> 20;
            p_k0 = (tile % tiles_k) * TZ;
            p_j0 = (tile / tiles_k) * TY;
            chunk_range(chunk, nchunks, nsplit, p.xchunk, p.p0, p.p1, p_l0, p_l1, p.monotone);
        }
        p_n = -1;
    };
    auto produce = [&]() {
        const int stg = p_g % kStages;
        if (p_g >= kStages) mbar_wait(empty + stg, (uint32_t)(((p_g / kStages) - 1) & 1));
        unsigned char *st = stages + (size_t)stg * L::bytes;
        uint64_t *bar = full + stg;
        ++p_g;
        if (p_item < 0) {   // end of work: publish the sentinel and complete the phase without bytes
            ring[p_q & 3] = -1;
            __threadfence_block();
            mbar_arrive(bar);
            p_done = true;
            return;
        }
        const int ck = PHASE == 1 ? p_k0 - 4 : p_k0, cj = PHASE == 1 ? p_j0 - 1 : p_j0;
        if (p_n < 0) {
            // x-neighbour plane (i-1 for E, i+1 for H) of the item's first plane: operand components B and C.  They land in the
            // stage's `own` buffer: a 4-D box is dense, so components B and C of the operand buffer do not start on the 128-byte
            // boundary a TMA destination needs.
            ring[p_q & 3] = p_item;
            __threadfence_block();
            mbar_expect_tx(bar, (uint32_t)L::tx_x);
            tma_load_4d(st + L::oOwn, &maps.opx, bar, ck, cj, PHASE == 1 ? p_l0 : p_l1 + 1, 1);
        } else {
            const int pl = PHASE == 1 ? (p_l0 + p_n + 1) : (p_l1 - 1 - p_n + 1);
            mbar_expect_tx(bar, (uint32_t)L::tx);
            tma_load_4d(st + L::oOp, &maps.op, bar, ck, cj, pl, 0);
            tma_load_4d(st + L::oOwn, &maps.own, bar, p_k0, p_j0, pl, 0);
            tma_load_4d(st + L::oId, &maps.id, bar, p_k0, p_j0, pl, 0);
        }
        if (++p_n == p_l1 - p_l0) {
            ++p_q;
            fetch();
        }
    };
    if (PW) {
        // dedicated producer warp: one lane walks the slot sequence, held back only by the `empty` barriers
        if (tid >= kTmaThreads) {
            if (lane == 0) {
                fetch();
                while (!p_done) produce();
            }
            return;
        }
    } else if (tid == 0) {
        // thread 0 of the first consumer warp: fill the ring now, then one slot per consumed slot (right after the warp
        // has released the slot it read, before its own arithmetic)
        fetch();
        for (int s = 0; s < kStages && !p_done; ++s) produce();
    }

    // ---------------- consumers
    const int r = tid / (TZ / 4), c = (tid % (TZ / 4)) * 4;
    const int e = r * TZ + c;

    int g = 0;   // consumer position in the slot sequence
    for (int q = 0;; ++q) {
    // ---------------- item prologue: the x-neighbour slot (or the end-of-work signal)
    mbar_wait(full + (g % kStages), (uint32_t)((g / kStages) & 1));
    const int item = ring[q & 3];
    if (item < 0) break;
#include "gpb_tma_item.inc"
    }   // items

    // the last CTA to finish re-arms the scheduler for the next launch
    if (p.persist && tid == 0) {
        __threadfence();
        if (atomicAdd(sched + 1, 1) == (int)(gridDim.x * gridDim.y) - 1) {
            sched[0] = 0;
            sched[1] = 0;
            __threadfence();
        }
    }
}

}  // namespace gpb
