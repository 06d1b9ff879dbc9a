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
        c = v;   // increasing x (concurrent H / E kernels)
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
    constexpr int TW = DISP == 1 ? 2 : 1;   // V4 vectors per 4 cells of T
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
    const int dc_bytes = DISP ? (int)((p.nmat * p.maxpoles * 3 * TW * sizeof(R) + 127) / 128 * 128) : 0;
    V4<R> *stf = reinterpret_cast<V4<R> *>(smem_raw + 128 + coef_bytes + tab_bytes + pf_bytes + dc_bytes);
    const int tslot = DISP ? 3 * p.maxpoles * TW * kTmaThreads : 0;   // V4 vectors per plane of T slots
    const int tf_bytes = DISP ? p.t_depth * tslot * (int)sizeof(V4<R>) : 0;
    unsigned char *stages = smem_raw + 128 + coef_bytes + tab_bytes + pf_bytes + dc_bytes + tf_bytes;

    const int tid = threadIdx.x, lane = tid & 31;
    const int W = tiles * (nchunks + nsplit);

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
        for (int m = tid; m < p.nmat * p.maxpoles * 3 * TW; m += kTmaThreads + 32 * PW) sdc[m] = src[m];
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
            const int tile = p_item & 0xfffff, chunk = p_item >> 20;
            p_k0 = (tile % tiles_k) * TZ;
            p_j0 = (tile / tiles_k) * TY;
            chunk_range(chunk, nchunks, nsplit, p.xchunk, p.p0, p.p1, p_l0, p_l1, p.monotone);
            if (PHASE == 1 && p.progress) {
                // concurrent H kernel: the H fields of this chunk's planes and of the plane in front of them must be final.
                // The finished items were published with fence + atomic by the H kernel's warps; acquire here, then order the
                // TMA (async proxy) loads that follow behind what was acquired.
                unsigned long long t0;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                for (int cc = chunk > 0 ? chunk - 1 : chunk; cc <= chunk; ++cc) {
                    unsigned v;
                    for (;;) {
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.progress + cc) : "memory");
                        if (v >= p.prog_need) break;
                        __nanosleep(100);
                        unsigned long long t1;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                        if (t1 - t0 > p.prog_timeout_ns) { atomicOr(p.prog_flags, 1u); break; }
                    }
                }
                asm volatile("fence.proxy.async;" ::: "memory");
            }
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
    // element offsets of my 4 cells inside the operand / own tiles
    const int eo = PHASE == 1 ? ((r + 1) * L::PA + c + 4) : (r * L::PA + c);   // centre
    const int eoj = PHASE == 1 ? (eo - L::PA) : (eo + L::PA);                    // j-1 (E) / j+1 (H)
    const int e = r * TZ + c;

    // fields this phase writes (the operand arrays are read-only in this phase)
    R *__restrict__ F0 = PHASE == 1 ? p.Ex : p.Hx;
    R *__restrict__ F1 = PHASE == 1 ? p.Ey : p.Hy;
    R *__restrict__ F2 = PHASE == 1 ? p.Ez : p.Hz;

    int g = 0;   // consumer position in the slot sequence
    for (int q = 0;; ++q) {
    // ---------------- item prologue: the x-neighbour slot (or the end-of-work signal)
    mbar_wait(full + (g % kStages), (uint32_t)((g / kStages) & 1));
    const int item = ring[q & 3];
    if (item < 0) break;
    const int tile = item & 0xfffff, chunkid = item >> 20;
    const int k0 = (tile % tiles_k) * TZ, j0 = (tile / tiles_k) * TY;
    const int j = j0 + r, k = k0 + c;
    int l0, l1;
    chunk_range(chunkid, nchunks, nsplit, p.xchunk, p.p0, p.p1, l0, l1, p.monotone);
    const int nl = l1 - l0;
    V4<R> qb, qc;   // register queue: operand B / C of the x-neighbour plane at my cells
    {
        const R *sX = reinterpret_cast<const R *>(stages + (size_t)(g % kStages) * L::bytes + L::oOwn);
        qb = ld4(sX + eo);
        qc = ld4(sX + L::CS + eo);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + (g % kStages));
        if (!PW && tid == 0 && !p_done) produce();
        ++g;
    }

    const bool valid = j <= p.ny && k <= p.nz;
    const JK4 bx = jk4_of(p.box[0].lo, p.box[0].hi, j, k), by = jk4_of(p.box[1].lo, p.box[1].hi, j, k), bz = jk4_of(p.box[2].lo, p.box[2].hi, j, k);
    unsigned smask = 0;
#pragma unroll
    for (int s = 0; s < kMaxSlabs; ++s)
        // (slabs that do not meet this tile in (j, k) at all are skipped with a CTA-uniform test)
        if (s < p.nslabs && (p.zfused || p.slab[s].axis != 2) && valid && p.slab[s].lo[1] < j0 + TY && p.slab[s].hi[1] > j0 && p.slab[s].lo[2] < k0 + TZ &&
            p.slab[s].hi[2] > k0)
            smask |= jk4_of(p.slab[s].lo, p.slab[s].hi, j, k).kmask << (4 * s);
    const bool any = valid && ((bx.kmask | by.kmask | bz.kmask) != 0u || smask != 0u);
    // fast cells: all 4 inside all three update boxes and outside every y- and z-slab footprint; on the planes between the x
    // slabs (p.fast_i0 <= i < p.fast_i1) such a thread needs no masks and no slab logic
    unsigned yfoot = 0;
#pragma unroll
    for (int s = 0; s < kMaxSlabs; ++s)
        if (s < p.nslabs && p.slab[s].axis != 0) yfoot |= (smask >> (4 * s)) & 0xfu;
    const bool fast_jk = valid && bx.kmask == 0xfu && by.kmask == 0xfu && bz.kmask == 0xfu && yfoot == 0u;
    const long long eoff = valid ? ((long long)j * p.pitch + k) : 0;

    // Phi prefetch: the lowest-numbered slab that touches my cells on a plane is fetched one plane ahead (pf_depth 2) or at the
    // top of the same iteration (pf_depth 1) with per-thread cp.async into a private shared-memory slot, so the DRAM latency of
    // Phi hides behind the TMA wait and the base update; any further slab on the same cells (edges, corners) loads directly
    unsigned smask6 = 0;   // bit s: slab s touches at least one of my cells
#pragma unroll
    for (int s = 0; s < kMaxSlabs; ++s)
        if ((smask >> (4 * s)) & 0xfu) smask6 |= 1u << s;
    // Cooperative z-slab corrections.  A z slab covers ~10 cells at one end of every z row: handled per thread, 3 of the 16
    // threads of a row would run the whole correction for their 4 cells while the other 13 idle (r1k: the z corrections were the
    // largest PML cost, executed by 40 % of all warps on 6 of 32 lanes).  Instead lane c of a row takes ONE cell (column zkq0 + c
    // of its own row) and both components, and hands the correction to the owner of the cell by warp shuffle.  Needs a single z
    // slab in the tile with at most TZ/4 columns here; otherwise that slab stays on the per-thread path.
    constexpr int LPR = TZ / 4;   // lanes per tile row
    int zs = -1;
    if (p.zfused && !p.znocoop) {
#pragma unroll
        for (int s = 0; s < kMaxSlabs; ++s)
            if (s < p.nslabs && p.slab[s].axis == 2 && min(k0 + TZ, p.slab[s].hi[2]) > max(k0, p.slab[s].lo[2])) zs = zs == -1 ? s : -2;
    }
    int zkq0 = 0;
    if (zs >= 0) {
        zkq0 = max(p.slab[zs].ko, k0);
        if (LPR > 32 || min(p.slab[zs].hi[2], k0 + TZ) - zkq0 > LPR) zs = -1;
    }
    const bool zcoop = zs >= 0;
    const int zc = zcoop ? zs : 0;                      // index used for parameter reads (any valid slab when unused)
    const int zk = zkq0 + (tid % LPR);                  // my cooperative cell: own row j, column zk
    const bool zok = zcoop && j <= p.ny && j >= p.slab[zc].lo[1] && j < p.slab[zc].hi[1] && zk >= p.slab[zc].lo[2] && zk < p.slab[zc].hi[2];
    const int zdepth = p.slab[zc].minus ? (p.slab[zc].dref - zk) : (zk - p.slab[zc].dref);
    const long long zphi_off = zok ? ((long long)(j - p.slab[zc].lo[1]) * p.slab[zc].n2 + (zk - p.slab[zc].ko)) : 0;
    const unsigned zm = zcoop ? ((smask >> (4 * zc)) & 0xfu) : 0u;   // cells of my own quad inside that slab
    if (zcoop) {   // the per-thread machinery (prefetch, slab loop) no longer sees the slab
        smask6 &= ~(1u << zc);
    }
    const bool pf = smask6 != 0u && p.pf_depth > 0;
    auto i_of = [&](int n) { return p.x_start + (PHASE == 1 ? (l0 + n + 1) : (l1 - 1 - n + 1)) - 1; };
    // Order of the slab corrections on a cell: x and y slabs in G.pmls order, then the z slabs -- the same order in every
    // kernel family (the register-vectorised path applies its z slabs in a second kernel, k_pml_slabs), so that edge and
    // corner cells, which take two or three corrections, get the same bits whichever family a grid or a shard runs on.
    unsigned zbits = 0;
#pragma unroll
    for (int s = 0; s < kMaxSlabs; ++s)
        if (s < p.nslabs && p.slab[s].axis == 2) zbits |= 1u << s;
    auto first_of = [&](unsigned m) { return __ffs((m & ~zbits) ? (m & ~zbits) : m) - 1; };
    auto prefetch = [&](V4<R> *slot, unsigned pmn, int ii) {   // Phi of the first slab (in application order) of pmn at plane ii -> my slot
        const SlabDev<R> &sl = p.slab[first_of(pmn)];
        const R *phi = sl.phi + ((long long)(ii - sl.lo[0]) * sl.n1 + (j - sl.lo[1])) * sl.n2 + (k - sl.ko);
        for (int q = 0; q < 2 * PORDER; ++q) cp_async_v4<R>(slot + q * kTmaThreads, phi + q * sl.ostride);
    };
    // slabs on a plane (CTA-uniform): the same set on every plane of the item unless the item straddles the edge of an x slab
    unsigned act_all = 0, act_some = 0;
    {
        const int ilo = min(i_of(0), i_of(nl - 1)), ihi = max(i_of(0), i_of(nl - 1));
#pragma unroll
        for (int s = 0; s < kMaxSlabs; ++s)
            if (s < p.nslabs) {
                if (ilo >= p.slab[s].lo[0] && ihi < p.slab[s].hi[0]) act_all |= 1u << s;
                if (ihi >= p.slab[s].lo[0] && ilo < p.slab[s].hi[0]) act_some |= 1u << s;
            }
    }
    const bool act_same = act_all == act_some;
    auto act_of = [&](int n) { return act_same ? act_all : slabs_on_plane(p, i_of(n)); };
    // T prefetch (dispersive): every thread with cells on the grid fetches its 4 cells of every pole and component
    const bool tpf = DISP && p.t_depth > 0 && any;
    auto tprefetch = [&](V4<R> *slot, int n) {   // T of plane n of this item -> my slots [comp][pole][TW]
        const long long off = (long long)(PHASE == 1 ? (l0 + n + 1) : (l1 - 1 - n + 1)) * p.plane + eoff;
        for (int cq = 0; cq < 3 * p.maxpoles; ++cq) {
            const R *g = reinterpret_cast<const R *>(p.T[cq / p.maxpoles]) + ((long long)(cq % p.maxpoles) * p.tstride + off) * TW;
            for (int w = 0; w < TW; ++w) cp_async_v4<R>(slot + (cq * TW + w) * kTmaThreads, g + 4 * w);
        }
    };
    // One cp.async group per plane, committed at the top of the loop body, holds whatever is fetched then: Phi and / or T of
    // this plane (distance 1) or of the next one (distance 2).  Data of distance 2 is complete after wait_group 1, of distance 1
    // after wait_group 0.  A distance-2 user needs one group in front of the loop for plane 0.
    const bool pf2 = pf && p.pf_depth == 2, tpf2 = tpf && p.t_depth == 2;
    // (the commits are uniform per thread; whether a thread has anything in the group does not matter)
    const bool grouped = (p.pf_depth > 0 && p.nslabs > 0) || (DISP && p.t_depth > 0);
    const bool lead_group = (p.pf_depth == 2 && p.nslabs > 0) || (DISP && p.t_depth == 2);
    if (lead_group) {
        if (pf2 && (act_of(0) & smask6)) prefetch(spf + tid, act_of(0) & smask6, i_of(0));
        if (tpf2) tprefetch(stf + tid, 0);
        cp_async_commit();
    }

    // dispersive sum of one component on my 4 cells (all poles): T from my prefetch slot or from global memory, advanced and
    // written back; u = base update - srce * phi on the cells inside the component's update box
    auto disp_apply = [&](int comp, const IdQ<IDT> &idq, unsigned m, const V4<R> &eold, V4<R> &u, int n, int pl) {
        if (!m) return;
        float ph0 = 0, ph1 = 0, ph2 = 0, ph3 = 0;
        const unsigned i0 = idq.at(0);
        unsigned i1 = i0, i2 = i0, i3 = i0;
        if (!idq.uniform()) { i1 = idq.at(1); i2 = idq.at(2); i3 = idq.at(3); }
        const int P = p.maxpoles;
        R *Tg = reinterpret_cast<R *>(p.T[comp]) + ((long long)pl * p.plane + eoff) * TW;
        const V4<R> *slot = tpf ? stf + (size_t)(n % p.t_depth) * tslot + (size_t)comp * P * TW * kTmaThreads + tid : nullptr;
        for (int q = 0; q < P; ++q) {
            R *Tq = Tg + (long long)q * p.tstride * TW;
            if (DISP == 2) {
                V4<R> t = slot ? slot[(size_t)q * kTmaThreads] : ld4(Tq);
                if (m & 1u) disp_cell_r(sdc + ((size_t)i0 * P + q) * 3, eold.x, t.x, ph0);
                if (m & 2u) disp_cell_r(sdc + ((size_t)i1 * P + q) * 3, eold.y, t.y, ph1);
                if (m & 4u) disp_cell_r(sdc + ((size_t)i2 * P + q) * 3, eold.z, t.z, ph2);
                if (m & 8u) disp_cell_r(sdc + ((size_t)i3 * P + q) * 3, eold.w, t.w, ph3);
                st4(Tq, t);
            } else {
                const Cplx<R> *dc = reinterpret_cast<const Cplx<R> *>(sdc);
                V4<R> t01 = slot ? slot[(size_t)(2 * q) * kTmaThreads] : ld4(Tq), t23 = slot ? slot[(size_t)(2 * q + 1) * kTmaThreads] : ld4(Tq + 4);
                if (m & 1u) disp_cell_c(dc + ((size_t)i0 * P + q) * 3, eold.x, t01.x, t01.y, ph0);
                if (m & 2u) disp_cell_c(dc + ((size_t)i1 * P + q) * 3, eold.y, t01.z, t01.w, ph1);
                if (m & 4u) disp_cell_c(dc + ((size_t)i2 * P + q) * 3, eold.z, t23.x, t23.y, ph2);
                if (m & 8u) disp_cell_c(dc + ((size_t)i3 * P + q) * 3, eold.w, t23.z, t23.w, ph3);
                st4(Tq, t01);
                st4(Tq + 4, t23);
            }
        }
        u.x = disp_sub(u.x, ssrc[i0], ph0);
        u.y = disp_sub(u.y, ssrc[i1], ph1);
        u.z = disp_sub(u.z, ssrc[i2], ph2);
        u.w = disp_sub(u.w, ssrc[i3], ph3);
    };

    for (int n = 0; n < nl; ++n, ++g) {
        const unsigned act = act_of(n);
        if (grouped) {
            if (pf) {
                const int np = n + p.pf_depth - 1;   // plane fetched now
                const unsigned pmn = (p.pf_depth == 2 ? (n + 1 < nl ? act_of(n + 1) : 0u) : act) & smask6;
                if (pmn) prefetch(spf + (size_t)(np % p.pf_depth) * 2 * PORDER * kTmaThreads + tid, pmn, i_of(np));
            }
            if (tpf) {
                const int np = n + p.t_depth - 1;
                if (np < nl) tprefetch(stf + (size_t)(np % p.t_depth) * tslot + tid, np);
            }
            cp_async_commit();
        }
        const int pl = PHASE == 1 ? (l0 + n + 1) : (l1 - 1 - n + 1);
        const int i = p.x_start + pl - 1;
        // cooperative z slab: my cell's Phi (both components) straight into registers, before the wait for the plane
        const bool zact = zcoop && i >= p.slab[zc].lo[0] && i < p.slab[zc].hi[0];   // CTA-uniform
        R *zphi = nullptr;
        R zP[2 * PORDER];
        if (zact && zok) {
            zphi = p.slab[zc].phi + (long long)(i - p.slab[zc].lo[0]) * p.slab[zc].n1 * p.slab[zc].n2 + zphi_off;
#pragma unroll
            for (int q = 0; q < 2 * PORDER; ++q) zP[q] = zphi[q * p.slab[zc].ostride];
        }
        const unsigned char *st = stages + (size_t)(g % kStages) * L::bytes;
        mbar_wait(full + (g % kStages), (uint32_t)((g / kStages) & 1));
        const R *sOp = reinterpret_cast<const R *>(st + L::oOp);
        const R *sOwn = reinterpret_cast<const R *>(st + L::oOwn);
        // E phase: A = Hx (needs j-1, k-1), B = Hy (k-1), C = Hz (j-1) ; H phase: A = Ex (j+1, k+1), B = Ey (k+1), C = Ez (j+1)
        const V4<R> a_c = ld4(sOp + eo), a_j = ld4(sOp + eoj);
        const V4<R> b_c = ld4(sOp + L::CS + eo);
        const V4<R> c_c = ld4(sOp + 2 * L::CS + eo), c_j = ld4(sOp + 2 * L::CS + eoj);
        R a_k, b_k;
        if (PHASE == 1) {
            a_k = __shfl_up_sync(0xffffffffu, a_c.w, 1);
            b_k = __shfl_up_sync(0xffffffffu, b_c.w, 1);
            if (c == 0 || lane == 0) {
                a_k = sOp[eo - 1];
                b_k = sOp[L::CS + eo - 1];
            }
        } else {
            a_k = __shfl_down_sync(0xffffffffu, a_c.x, 1);
            b_k = __shfl_down_sync(0xffffffffu, b_c.x, 1);
            if (c == TZ - 4 || lane == 31) {
                a_k = sOp[eo + 4];
                b_k = sOp[L::CS + eo + 4];
            }
        }
        V4<R> f0 = ld4(sOwn + e), f1 = ld4(sOwn + L::OS + e), f2 = ld4(sOwn + 2 * L::OS + e);
        const unsigned pm = act & smask6;   // slabs on my cells on this plane; 0 for fast threads by construction
        // A warp without slab cells on this plane has now taken what it needs from the stage and hands it back to the producer.
        // A warp with slab cells keeps it until its PML corrections are done: they re-read their operands and IDs from the stage
        // instead of keeping them in registers (which spilled the straight-line update of every thread).
        // (the IDs are read here, before the release: the stage is refilled as soon as the last warp has arrived)
        const IdQ<IDT> id0 = IdQ<IDT>::load(st + L::oId, e), id1 = IdQ<IDT>::load(st + L::oId, L::OS + e), id2 = IdQ<IDT>::load(st + L::oId, 2 * L::OS + e);
        const bool wpml = __any_sync(0xffffffffu, pm != 0u) || zact;
        if (!wpml) {
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + (g % kStages));
            if (!PW && tid == 0 && !p_done) produce();
        }

        bool w0 = false, w1 = false, w2 = false;
        if (any) {
            // one straight-line update for every thread; threads with cells outside an update box (domain faces, the
            // two x-slab plane ranges) put the old value back per cell afterwards
            const bool fast = fast_jk && i >= p.fast_i0 && i < p.fast_i1;
            unsigned m0 = 0xfu, m1 = 0xfu, m2 = 0xfu;
            if (!fast) {
                m0 = (i >= p.box[0].lo[0] && i < p.box[0].hi[0]) ? bx.kmask : 0u;
                m1 = (i >= p.box[1].lo[0] && i < p.box[1].hi[0]) ? by.kmask : 0u;
                m2 = (i >= p.box[2].lo[0] && i < p.box[2].hi[0]) ? bz.kmask : 0u;
            }
            // E phase: backward differences (c - neighbour); H phase: forward (neighbour - c)
            // E phase: Ex = CA Ex + CBy dHz/dy - CBz dHy/dz ; Ey = CA Ey + CBz dHx/dz - CBx dHz/dx ; Ez = CA Ez + CBx dHy/dx - CBy dHx/dy
            // H phase: Hx = DA Hx - DBy dEz/dy + DBz dEy/dz ; Hy = DA Hy - DBz dEx/dz + DBx dEz/dx ; Hz = DA Hz - DBx dEy/dx + DBy dEx/dy
            {
                V4<R> dC_dy, dB_dz;
                if (PHASE == 1) {
                    dB_dz = {b_c.x - b_k, b_c.y - b_c.x, b_c.z - b_c.y, b_c.w - b_c.z};       // dHy/dz
                    dC_dy = {c_c.x - c_j.x, c_c.y - c_j.y, c_c.z - c_j.z, c_c.w - c_j.w};     // dHz/dy
                } else {
                    dB_dz = {b_c.y - b_c.x, b_c.z - b_c.y, b_c.w - b_c.z, b_k - b_c.w};       // dEy/dz
                    dC_dy = {c_j.x - c_c.x, c_j.y - c_c.y, c_j.z - c_c.z, c_j.w - c_c.w};     // dEz/dy
                }
                Coef4<R> q0, q1, q2, q3;
                coef4q(scoef, id0, q0, q1, q2, q3);
                V4<R> u;
                if (PHASE == 1) {
                    u.x = upd3(q0.a, f0.x, q0.by, dC_dy.x, -q0.bz, dB_dz.x);
                    u.y = upd3(q1.a, f0.y, q1.by, dC_dy.y, -q1.bz, dB_dz.y);
                    u.z = upd3(q2.a, f0.z, q2.by, dC_dy.z, -q2.bz, dB_dz.z);
                    u.w = upd3(q3.a, f0.w, q3.by, dC_dy.w, -q3.bz, dB_dz.w);
                } else {
                    u.x = upd3(q0.a, f0.x, -q0.by, dC_dy.x, q0.bz, dB_dz.x);
                    u.y = upd3(q1.a, f0.y, -q1.by, dC_dy.y, q1.bz, dB_dz.y);
                    u.z = upd3(q2.a, f0.z, -q2.by, dC_dy.z, q2.bz, dB_dz.z);
                    u.w = upd3(q3.a, f0.w, -q3.by, dC_dy.w, q3.bz, dB_dz.w);
                }
                if (DISP) {
                    if (tpf) { if (p.t_depth == 2) cp_async_wait1(); else cp_async_wait0(); }
                    disp_apply(0, id0, m0, f0, u, n, pl);
                }
                if (!fast) { u.x = sel(m0, 0, u.x, f0.x); u.y = sel(m0, 1, u.y, f0.y); u.z = sel(m0, 2, u.z, f0.z); u.w = sel(m0, 3, u.w, f0.w); }
                f0 = u;
            }
            {
                V4<R> dA_dz, dC_dx;
                if (PHASE == 1) {
                    dA_dz = {a_c.x - a_k, a_c.y - a_c.x, a_c.z - a_c.y, a_c.w - a_c.z};       // dHx/dz
                    dC_dx = {c_c.x - qc.x, c_c.y - qc.y, c_c.z - qc.z, c_c.w - qc.w};         // dHz/dx
                } else {
                    dA_dz = {a_c.y - a_c.x, a_c.z - a_c.y, a_c.w - a_c.z, a_k - a_c.w};       // dEx/dz
                    dC_dx = {qc.x - c_c.x, qc.y - c_c.y, qc.z - c_c.z, qc.w - c_c.w};         // dEz/dx
                }
                Coef4<R> q0, q1, q2, q3;
                coef4q(scoef, id1, q0, q1, q2, q3);
                V4<R> u;
                if (PHASE == 1) {
                    u.x = upd3(q0.a, f1.x, q0.bz, dA_dz.x, -q0.bx, dC_dx.x);
                    u.y = upd3(q1.a, f1.y, q1.bz, dA_dz.y, -q1.bx, dC_dx.y);
                    u.z = upd3(q2.a, f1.z, q2.bz, dA_dz.z, -q2.bx, dC_dx.z);
                    u.w = upd3(q3.a, f1.w, q3.bz, dA_dz.w, -q3.bx, dC_dx.w);
                } else {
                    u.x = upd3(q0.a, f1.x, -q0.bz, dA_dz.x, q0.bx, dC_dx.x);
                    u.y = upd3(q1.a, f1.y, -q1.bz, dA_dz.y, q1.bx, dC_dx.y);
                    u.z = upd3(q2.a, f1.z, -q2.bz, dA_dz.z, q2.bx, dC_dx.z);
                    u.w = upd3(q3.a, f1.w, -q3.bz, dA_dz.w, q3.bx, dC_dx.w);
                }
                if (DISP) disp_apply(1, id1, m1, f1, u, n, pl);
                if (!fast) { u.x = sel(m1, 0, u.x, f1.x); u.y = sel(m1, 1, u.y, f1.y); u.z = sel(m1, 2, u.z, f1.z); u.w = sel(m1, 3, u.w, f1.w); }
                f1 = u;
            }
            {
                V4<R> dB_dx, dA_dy;
                if (PHASE == 1) {
                    dA_dy = {a_c.x - a_j.x, a_c.y - a_j.y, a_c.z - a_j.z, a_c.w - a_j.w};     // dHx/dy
                    dB_dx = {b_c.x - qb.x, b_c.y - qb.y, b_c.z - qb.z, b_c.w - qb.w};         // dHy/dx
                } else {
                    dA_dy = {a_j.x - a_c.x, a_j.y - a_c.y, a_j.z - a_c.z, a_j.w - a_c.w};     // dEx/dy
                    dB_dx = {qb.x - b_c.x, qb.y - b_c.y, qb.z - b_c.z, qb.w - b_c.w};         // dEy/dx
                }
                Coef4<R> q0, q1, q2, q3;
                coef4q(scoef, id2, q0, q1, q2, q3);
                V4<R> u;
                if (PHASE == 1) {
                    u.x = upd3(q0.a, f2.x, q0.bx, dB_dx.x, -q0.by, dA_dy.x);
                    u.y = upd3(q1.a, f2.y, q1.bx, dB_dx.y, -q1.by, dA_dy.y);
                    u.z = upd3(q2.a, f2.z, q2.bx, dB_dx.z, -q2.by, dA_dy.z);
                    u.w = upd3(q3.a, f2.w, q3.bx, dB_dx.w, -q3.by, dA_dy.w);
                } else {
                    u.x = upd3(q0.a, f2.x, -q0.bx, dB_dx.x, q0.by, dA_dy.x);
                    u.y = upd3(q1.a, f2.y, -q1.bx, dB_dx.y, q1.by, dA_dy.y);
                    u.z = upd3(q2.a, f2.z, -q2.bx, dB_dx.z, q2.by, dA_dy.z);
                    u.w = upd3(q3.a, f2.w, -q3.bx, dB_dx.w, q3.by, dA_dy.w);
                }
                if (DISP) disp_apply(2, id2, m2, f2, u, n, pl);
                if (!fast) { u.x = sel(m2, 0, u.x, f2.x); u.y = sel(m2, 1, u.y, f2.y); u.z = sel(m2, 2, u.z, f2.z); u.w = sel(m2, 3, u.w, f2.w); }
                f2 = u;
            }
            w0 = m0 != 0u; w1 = m1 != 0u; w2 = m2 != 0u;
        }
        if (wpml) {   // warp-uniform
            {
                const V4<R> *slot = nullptr;
                if (pf && pm) {
                    if (p.pf_depth == 2) cp_async_wait1();
                    else cp_async_wait0();
                    slot = spf + (size_t)(n % p.pf_depth) * 2 * PORDER * kTmaThreads + tid;
                }
                // slabs any lane of this warp has to apply on this plane: x / y slabs in G.pmls order, then the z slabs
                // (warp-uniform: the cooperative z step needs every lane)
                const unsigned wtodo = __reduce_or_sync(0xffffffffu, pm) | (zact ? 1u << zs : 0u);
                for (int zpass = 0; zpass < 2; ++zpass)
                for (unsigned todo = wtodo & (zpass ? zbits : ~zbits); todo; todo &= todo - 1) {
                    const int s = __ffs(todo) - 1;
                    if (zact && s == zs) {
                        // ---- cooperative z slab: one cell per lane, both components
                        const SlabDev<R> &sl = p.slab[s];
                        R corr_a = 0, corr_b = 0;
                        if (zok) {
                            const PmlCo<R> co = pml_load_s(PFORM, PORDER, stab + (size_t)s * 4 * PORDER * p.tmax, p.tmax, zdepth);
                            const int zkk = zk - k0;
                            const int o = PHASE == 1 ? ((r + 1) * L::PA + zkk + 4) : (r * L::PA + zkk);
                            R dB, dA;   // d(operand B)/dz for the first component, d(operand A)/dz for the second, as in the quad path
                            if (PHASE == 1) { dB = sOp[L::CS + o] - sOp[L::CS + o - 1]; dA = sOp[o] - sOp[o - 1]; }
                            else { dB = sOp[L::CS + o + 1] - sOp[L::CS + o]; dA = sOp[o + 1] - sOp[o]; }
                            const IDT *sid = reinterpret_cast<const IDT *>(st + L::oId);
                            const unsigned ida = sid[r * TZ + zkk], idb = sid[L::OS + r * TZ + zkk];
                            R Pa0 = zP[0], Pb0 = zP[1], Pa1 = PORDER == 2 ? zP[2 * (PORDER - 1)] : Pa0, Pb1 = PORDER == 2 ? zP[2 * (PORDER - 1) + 1] : Pb0;
                            corr_a = mul_(ssrc[ida], pml_apply(PFORM, PORDER, co, mul_(dB, sl.inv_d), Pa0, Pa1));
                            corr_b = mul_(ssrc[idb], pml_apply(PFORM, PORDER, co, mul_(dA, sl.inv_d), Pb0, Pb1));
                            zphi[0] = Pa0;
                            zphi[sl.ostride] = Pb0;
                            if (PORDER == 2) { zphi[2 * sl.ostride] = Pa1; zphi[3 * sl.ostride] = Pb1; }
                        }
                        // to the owners: cell e of my quad was computed by lane (row base) + (k - zkq0) + e
                        const int src0 = (lane & ~(LPR - 1)) + (k - zkq0);
                        const R sa = PHASE == 1 ? (R)-1 : (R)1;   // E: Ex -= , Ey += ; H: Hx += , Hy -=
                        const R ca0 = __shfl_sync(0xffffffffu, corr_a, src0 & 31), ca1 = __shfl_sync(0xffffffffu, corr_a, (src0 + 1) & 31);
                        const R ca2 = __shfl_sync(0xffffffffu, corr_a, (src0 + 2) & 31), ca3 = __shfl_sync(0xffffffffu, corr_a, (src0 + 3) & 31);
                        const R cb0 = __shfl_sync(0xffffffffu, corr_b, src0 & 31), cb1 = __shfl_sync(0xffffffffu, corr_b, (src0 + 1) & 31);
                        const R cb2 = __shfl_sync(0xffffffffu, corr_b, (src0 + 2) & 31), cb3 = __shfl_sync(0xffffffffu, corr_b, (src0 + 3) & 31);
                        if (zm & 1u) { f0.x = fma_(sa, ca0, f0.x); f1.x = fma_(-sa, cb0, f1.x); }
                        if (zm & 2u) { f0.y = fma_(sa, ca1, f0.y); f1.y = fma_(-sa, cb1, f1.y); }
                        if (zm & 4u) { f0.z = fma_(sa, ca2, f0.z); f1.z = fma_(-sa, cb2, f1.z); }
                        if (zm & 8u) { f0.w = fma_(sa, ca3, f0.w); f1.w = fma_(-sa, cb3, f1.w); }
                        if (zm) w0 = w1 = true;
                        continue;
                    }
                    if (!((pm >> s) & 1u)) continue;
                    const SlabDev<R> &sl = p.slab[s];
                    const unsigned m = (smask >> (4 * s)) & 0xfu;
                    const R *tb = stab + (size_t)s * 4 * PORDER * p.tmax;
                    R *phi = sl.phi + ((long long)(i - sl.lo[0]) * sl.n1 + (j - sl.lo[1])) * sl.n2 + (k - sl.ko);
                    const V4<R> *slb = slot ? slot + kTmaThreads : nullptr;
                    // Component / derivative / sign table of SURVEY.md section 8a (identical in all 48 reference kernels):
                    //   E phase  x: Ey -= dHz/dx, Ez += dHy/dx   y: Ex += dHz/dy, Ez -= dHx/dy   z: Ex -= dHy/dz, Ey += dHx/dz
                    //   H phase  x: Hy += dEz/dx, Hz -= dEy/dx   y: Hx -= dEz/dy, Hz += dEx/dy   z: Hx += dEy/dz, Hy -= dEx/dz
                    // The derivatives are formed again from the stage (same operands, same expression as above).
                    if (sl.axis == 0) {
                        const int depth = sl.minus ? (sl.dref - i) : (i - sl.dref);
                        V4<R> dF;
                        if (PHASE == 1) dF = {c_c.x - qc.x, c_c.y - qc.y, c_c.z - qc.z, c_c.w - qc.w};
                        else dF = {qc.x - c_c.x, qc.y - c_c.y, qc.z - c_c.z, qc.w - c_c.w};
                        pml_comp<R, 0>(PFORM, PORDER, tb, p.tmax, depth, 0, sl.inv_d, m, lds_ids4<IDT>(st + L::oId, L::OS + e), ssrc, PHASE == 1 ? (R)-1 : (R)1, dF, f1, phi,
                                       2 * sl.ostride, slot, 2 * kTmaThreads);
                        if (PHASE == 1) dF = {b_c.x - qb.x, b_c.y - qb.y, b_c.z - qb.z, b_c.w - qb.w};
                        else dF = {qb.x - b_c.x, qb.y - b_c.y, qb.z - b_c.z, qb.w - b_c.w};
                        pml_comp<R, 0>(PFORM, PORDER, tb, p.tmax, depth, 0, sl.inv_d, m, lds_ids4<IDT>(st + L::oId, 2 * L::OS + e), ssrc, PHASE == 1 ? (R)1 : (R)-1, dF, f2,
                                       phi + sl.ostride, 2 * sl.ostride, slb, 2 * kTmaThreads);
                        w1 = w2 = true;
                    } else if (sl.axis == 1) {
                        const int depth = sl.minus ? (sl.dref - j) : (j - sl.dref);
                        V4<R> dF;
                        {
                            const V4<R> cj = ld4(sOp + 2 * L::CS + eoj);
                            if (PHASE == 1) dF = {c_c.x - cj.x, c_c.y - cj.y, c_c.z - cj.z, c_c.w - cj.w};
                            else dF = {cj.x - c_c.x, cj.y - c_c.y, cj.z - c_c.z, cj.w - c_c.w};
                        }
                        pml_comp<R, 0>(PFORM, PORDER, tb, p.tmax, depth, 0, sl.inv_d, m, lds_ids4<IDT>(st + L::oId, e), ssrc, PHASE == 1 ? (R)1 : (R)-1, dF, f0, phi,
                                       2 * sl.ostride, slot, 2 * kTmaThreads);
                        {
                            const V4<R> ac = ld4(sOp + eo), aj = ld4(sOp + eoj);
                            if (PHASE == 1) dF = {ac.x - aj.x, ac.y - aj.y, ac.z - aj.z, ac.w - aj.w};
                            else dF = {aj.x - ac.x, aj.y - ac.y, aj.z - ac.z, aj.w - ac.w};
                        }
                        pml_comp<R, 0>(PFORM, PORDER, tb, p.tmax, depth, 0, sl.inv_d, m, lds_ids4<IDT>(st + L::oId, 2 * L::OS + e), ssrc, PHASE == 1 ? (R)-1 : (R)1, dF, f2,
                                       phi + sl.ostride, 2 * sl.ostride, slb, 2 * kTmaThreads);
                        w0 = w2 = true;
                    } else {
                        const int depth = sl.minus ? (sl.dref - k) : (k - sl.dref), ds = sl.minus ? -1 : 1;
                        V4<R> dF;
                        {
                            const R bk = sOp[L::CS + eo + (PHASE == 1 ? -1 : 4)];
                            if (PHASE == 1) dF = {b_c.x - bk, b_c.y - b_c.x, b_c.z - b_c.y, b_c.w - b_c.z};
                            else dF = {b_c.y - b_c.x, b_c.z - b_c.y, b_c.w - b_c.z, bk - b_c.w};
                        }
                        pml_comp<R, 1>(PFORM, PORDER, tb, p.tmax, depth, ds, sl.inv_d, m, lds_ids4<IDT>(st + L::oId, e), ssrc, PHASE == 1 ? (R)-1 : (R)1, dF, f0, phi,
                                       2 * sl.ostride, slot, 2 * kTmaThreads);
                        {
                            const V4<R> ac = ld4(sOp + eo);
                            const R ak = sOp[eo + (PHASE == 1 ? -1 : 4)];
                            if (PHASE == 1) dF = {ac.x - ak, ac.y - ac.x, ac.z - ac.y, ac.w - ac.z};
                            else dF = {ac.y - ac.x, ac.z - ac.y, ac.w - ac.z, ak - ac.w};
                        }
                        pml_comp<R, 1>(PFORM, PORDER, tb, p.tmax, depth, ds, sl.inv_d, m, lds_ids4<IDT>(st + L::oId, L::OS + e), ssrc, PHASE == 1 ? (R)1 : (R)-1, dF, f1,
                                       phi + sl.ostride, 2 * sl.ostride, slb, 2 * kTmaThreads);
                        w0 = w1 = true;
                    }
                    slot = nullptr;   // only the first slab was prefetched
                }
            }
        }
        if (any) {
            const long long off = (long long)pl * p.plane + eoff;
            if (w0) st4(F0 + off, f0);
            if (w1) st4(F1 + off, f1);
            if (w2) st4(F2 + off, f2);
        }
        if (wpml) {
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + (g % kStages));
            if (!PW && tid == 0 && !p_done) produce();
        }
        qb = b_c;
        qc = c_c;
    }
    if (PHASE == 0 && p.progress) {   // this warp's part of the item is in memory: publish (see the E kernel's producer)
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(p.progress + chunkid, 1u);
    }
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
