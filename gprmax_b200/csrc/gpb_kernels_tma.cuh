// gpb_kernels_tma.cuh -- TMA-staged E/H half-step kernels (sm_100a), the bandwidth path.
//
// Why: the register-only vectorised kernels (gpb_kernels_v4.cuh) need ~120 registers per thread to
// keep one plane of operands in flight, which caps occupancy at 16 warps/SM and leaves them
// latency-bound at ~3.4 TB/s (profiles/r1b_*).  Here the operand planes are staged in shared memory
// by the Tensor Memory Accelerator instead:
//   * a CTA owns a TY x TZ tile of (j,k) and marches along x (E phase: +x, H phase: -x);
//   * for every plane one elected thread issues 9 `cp.async.bulk.tensor.3d` loads (6 field tiles, the
//     three operand tiles carrying their one-row / one-column halo, and 3 material-ID tiles) into one
//     stage of a kStages-deep ring; completion is tracked with an mbarrier (expect_tx / complete_tx);
//     out-of-range halo coordinates are zero-filled by the TMA unit, so there is no edge code;
//   * consumers read 128-bit rows from shared memory, so the j+-1 / k+-1 re-reads never touch L2, and
//     kStages planes (~30 KB each) are in flight per CTA regardless of register pressure;
//   * the x-neighbour plane (i-1 for E, i+1 for H) rides in a register queue as before;
//   * results go straight from registers to global memory with 128-bit stores.
// x / y PML slabs are applied in the same pass (vectorised, warp-uniform); z slabs by k_pml_slabs.
#pragma once
#include <cuda.h>

#include "gpb_kernels_v4.cuh"

namespace gpb {


struct TmaMaps9 {
    CUtensorMap opA;   // operand with both halos   (E phase: Hx ; H phase: Ex)  box (TZ+4) x (TY+1)
    CUtensorMap opB;   // operand with the k halo   (E phase: Hy ; H phase: Ey)  box (TZ+4) x TY
    CUtensorMap opC;   // operand with the j halo   (E phase: Hz ; H phase: Ez)  box TZ x (TY+1)
    CUtensorMap own0, own1, own2;  // fields being updated, box TZ x TY
    CUtensorMap id0, id1, id2;     // their material IDs,   box TZ x TY (elements of IDT)
};

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
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// shared-memory layout of one stage (byte offsets), every sub-buffer 128-byte aligned
template <typename R, typename IDT, int TY, int TZ>
struct StageLayout {
    static constexpr int a128(int x) { return (x + 127) / 128 * 128; }
    static constexpr int PA = TZ + 4;  // row pitch (elements) of opA / opB
    static constexpr int szA = a128((TY + 1) * PA * (int)sizeof(R));
    static constexpr int szB = a128(TY * PA * (int)sizeof(R));
    static constexpr int szC = a128((TY + 1) * TZ * (int)sizeof(R));
    static constexpr int szO = a128(TY * TZ * (int)sizeof(R));
    static constexpr int szI = a128(TY * TZ * (int)sizeof(IDT));
    static constexpr int oA = 0, oB = oA + szA, oC = oB + szB, oO0 = oC + szC, oO1 = oO0 + szO, oO2 = oO1 + szO;
    static constexpr int oI0 = oO2 + szO, oI1 = oI0 + szI, oI2 = oI1 + szI;
    static constexpr int bytes = oI2 + szI;
    // bytes the TMA unit delivers per stage (full boxes, out-of-range parts are zero-filled)
    static constexpr int tx = ((TY + 1) * PA + TY * PA + (TY + 1) * TZ + 3 * TY * TZ) * (int)sizeof(R) + 3 * TY * TZ * (int)sizeof(IDT);
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

// ------------------------------------------------------------------------------------------
// PHASE 1: electric half-step (marches +x, operands H, queue = Hy,Hz of plane i-1)
// PHASE 0: magnetic half-step (marches -x, operands E, queue = Ey,Ez of plane i+1)
// Arithmetic identical to k_update_e4 / k_update_h4.
// Shared memory: [kStages mbarriers][coefficient rows][kStages stages]
// ------------------------------------------------------------------------------------------
#ifndef GPB_TMA_CTAS
#define GPB_TMA_CTAS 2
#endif
// Work items = (tile, x-chunk) pairs.  Non-persistent launch: one CTA per item.  Persistent launch (p.persist):
// GPB_TMA_CTAS CTAs per SM pull items from an atomic counter and keep ONE continuous TMA pipeline running across item
// boundaries, so the ring never drains between marches (the ncu source page showed a quarter of all stall samples on
// the `full` mbarrier wait while fresh CTAs filled their pipelines).  Every item starts with a pseudo-plane slot that
// carries only the two x-neighbour operand tiles (register-queue initialisation); a final empty slot carries the
// end-of-work signal.  sched[0] = next item, sched[1] = finished CTAs (the last one resets both for the next launch).
template <typename R, typename IDT, int TY, int TZ, int kStages, int PHASE>
__global__ void __launch_bounds__(TY * TZ / 4, (sizeof(R) == 4 ? GPB_TMA_CTAS * 256 / (TY * TZ / 4) : 1))
k_update_tma(const PhaseParams<R> p, const __grid_constant__ TmaMaps9 maps, int tiles_k, int tiles, int nchunks, int *sched)
{
    constexpr int kTmaThreads = TY * TZ / 4;  // every thread owns 4 consecutive z cells of the tile
    static_assert(kTmaThreads % 32 == 0 && kTmaThreads <= 256, "tile shape");
    using L = StageLayout<R, IDT, TY, TZ>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);   // [kStages] TMA bytes landed
    uint64_t *empty = full + kStages;                           // [kStages] all warps have read the stage
    volatile int *ring = reinterpret_cast<volatile int *>(empty + kStages);  // [4] item ids, producer -> consumers
    Coef4<R> *scoef = reinterpret_cast<Coef4<R> *>(smem_raw + 128);
    R *ssrc = reinterpret_cast<R *>(scoef + p.nmat);
    const int coef_bytes = (int)((p.nmat * (sizeof(Coef4<R>) + sizeof(R)) + 127) / 128 * 128);
    unsigned char *stages = smem_raw + 128 + coef_bytes;

    const int tid = threadIdx.x, lane = tid & 31;
    const int r = tid / (TZ / 4), c = (tid % (TZ / 4)) * 4;
    const int W = tiles * nchunks;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, kTmaThreads / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&maps.opA);
        tma_prefetch_desc(&maps.opB);
        tma_prefetch_desc(&maps.opC);
        tma_prefetch_desc(&maps.own0);
        tma_prefetch_desc(&maps.id0);
    }
    for (int m = tid; m < p.nmat; m += kTmaThreads) {
        scoef[m] = p.coef[m];
        ssrc[m] = p.src[m];
    }
    __syncthreads();

    // ---------------- producer (thread 0): one slot of the CTA's slot sequence per call
    int p_item = -1, p_n = -1, p_q = 0, p_g = 0;   // item being loaded, its next plane (-1 = pseudo-plane), item ordinal, slot
    int p_j0 = 0, p_k0 = 0, p_l0 = 0, p_l1 = 0;
    bool p_first = true, p_done = false;
    auto fetch = [&]() {
        int w;
        if (p.persist) w = atomicAdd(sched, 1);
        else w = p_first ? (int)(blockIdx.y * gridDim.x + blockIdx.x) : W;
        p_first = false;
        p_item = w < W ? w : -1;
        if (p_item >= 0) {
            const int tile = p_item % tiles, chunk = p_item / tiles;
            p_k0 = (tile % tiles_k) * TZ;
            p_j0 = (tile / tiles_k) * TY;
            p_l0 = p.p0 + chunk * p.xchunk;
            p_l1 = min(p_l0 + p.xchunk, p.p1);
        }
        p_n = -1;
    };
    auto produce = [&]() {
        if (p_done) return;
        const int stg = p_g % kStages;
        if (p_g >= kStages) mbar_wait(empty + stg, (uint32_t)(((p_g / kStages) - 1) & 1));
        unsigned char *st = stages + (size_t)stg * L::bytes;
        uint64_t *bar = full + stg;
        if (p_item < 0) {   // end of work: publish the sentinel and complete the phase without bytes
            ring[p_q & 3] = -1;
            __threadfence_block();
            mbar_arrive(bar);
            p_done = true;
            ++p_g;
            return;
        }
        if (p_n < 0) {
            // pseudo-plane: x-neighbour plane (i-1 for E, i+1 for H) of the item's first plane; operand tiles B and C only
            ring[p_q & 3] = p_item;
            __threadfence_block();
            const int pl = PHASE == 1 ? p_l0 : p_l1 + 1;
            mbar_expect_tx(bar, (uint32_t)((TY * L::PA + (TY + 1) * TZ) * sizeof(R)));
            if (PHASE == 1) {
                tma_load_3d(st + L::oB, &maps.opB, bar, p_k0 - 4, p_j0, pl);
                tma_load_3d(st + L::oC, &maps.opC, bar, p_k0, p_j0 - 1, pl);
            } else {
                tma_load_3d(st + L::oB, &maps.opB, bar, p_k0, p_j0, pl);
                tma_load_3d(st + L::oC, &maps.opC, bar, p_k0, p_j0, pl);
            }
        } else {
            const int pl = PHASE == 1 ? (p_l0 + p_n + 1) : (p_l1 - 1 - p_n + 1);
            mbar_expect_tx(bar, (uint32_t)L::tx);
            if (PHASE == 1) {
                tma_load_3d(st + L::oA, &maps.opA, bar, p_k0 - 4, p_j0 - 1, pl);
                tma_load_3d(st + L::oB, &maps.opB, bar, p_k0 - 4, p_j0, pl);
                tma_load_3d(st + L::oC, &maps.opC, bar, p_k0, p_j0 - 1, pl);
            } else {
                tma_load_3d(st + L::oA, &maps.opA, bar, p_k0, p_j0, pl);
                tma_load_3d(st + L::oB, &maps.opB, bar, p_k0, p_j0, pl);
                tma_load_3d(st + L::oC, &maps.opC, bar, p_k0, p_j0, pl);
            }
            tma_load_3d(st + L::oO0, &maps.own0, bar, p_k0, p_j0, pl);
            tma_load_3d(st + L::oO1, &maps.own1, bar, p_k0, p_j0, pl);
            tma_load_3d(st + L::oO2, &maps.own2, bar, p_k0, p_j0, pl);
            tma_load_3d(st + L::oI0, &maps.id0, bar, p_k0, p_j0, pl);
            tma_load_3d(st + L::oI1, &maps.id1, bar, p_k0, p_j0, pl);
            tma_load_3d(st + L::oI2, &maps.id2, bar, p_k0, p_j0, pl);
        }
        ++p_g;
        if (++p_n == p_l1 - p_l0) {
            ++p_q;
            fetch();
        }
    };
    if (tid == 0) {
        fetch();
        for (int s = 0; s < kStages; ++s) produce();
    }

    // fields this phase writes (the operand arrays are read-only in this phase)
    R *__restrict__ F0 = PHASE == 1 ? p.Ex : p.Hx;
    R *__restrict__ F1 = PHASE == 1 ? p.Ey : p.Hy;
    R *__restrict__ F2 = PHASE == 1 ? p.Ez : p.Hz;

    int g = 0;   // consumer position in the slot sequence
    for (int q = 0;; ++q) {
    // ---------------- item prologue: the pseudo-plane slot (or the end-of-work signal)
    mbar_wait(full + (g % kStages), (uint32_t)((g / kStages) & 1));
    const int item = ring[q & 3];
    if (item < 0) break;
    const int tile = item % tiles, chunkid = item / tiles;
    const int k0 = (tile % tiles_k) * TZ, j0 = (tile / tiles_k) * TY;
    const int j = j0 + r, k = k0 + c;
    const int l0 = p.p0 + chunkid * p.xchunk;
    const int l1 = min(l0 + p.xchunk, p.p1);
    const int nl = l1 - l0;
    auto plane_of = [&](int n) { return PHASE == 1 ? (l0 + n + 1) : (l1 - 1 - n + 1); };
    V4<R> qb, qc;   // register queue: operand B / C of the x-neighbour plane at my cells
    {
        const unsigned char *st = stages + (size_t)(g % kStages) * L::bytes;
        const R *sB = reinterpret_cast<const R *>(st + L::oB);
        const R *sC = reinterpret_cast<const R *>(st + L::oC);
        if (PHASE == 1) {
            qb = ld4(sB + r * L::PA + c + 4);
            qc = ld4(sC + (r + 1) * TZ + c);
        } else {
            qb = ld4(sB + r * L::PA + c);
            qc = ld4(sC + r * TZ + c);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + (g % kStages));
        if (tid == 0) produce();
        ++g;
    }

    const bool valid = j <= p.ny && k <= p.nz;
    const JK4 bx = jk4_of(p.box[0].lo, p.box[0].hi, j, k), by = jk4_of(p.box[1].lo, p.box[1].hi, j, k), bz = jk4_of(p.box[2].lo, p.box[2].hi, j, k);
    unsigned smask = 0;
#pragma unroll
    for (int s = 0; s < kMaxSlabs; ++s)
        if (s < p.nslabs && p.slab[s].axis != 2 && valid) smask |= jk4_of(p.slab[s].lo, p.slab[s].hi, j, k).kmask << (4 * s);
    const bool any = valid && ((bx.kmask | by.kmask | bz.kmask) != 0u || smask != 0u);
    // fast path: all 4 cells inside all three update boxes and outside every y-slab footprint; then on the planes
    // between the x slabs (p.fast_i0 <= i < p.fast_i1) the update is straight-line code
    unsigned yfoot = 0;
#pragma unroll
    for (int s = 0; s < kMaxSlabs; ++s)
        if (s < p.nslabs && p.slab[s].axis == 1) yfoot |= (smask >> (4 * s)) & 0xfu;
    const bool fast_jk = valid && bx.kmask == 0xfu && by.kmask == 0xfu && bz.kmask == 0xfu && yfoot == 0u;
    const long long eoff = valid ? ((long long)j * p.pitch + k) : 0;

    for (int n = 0; n < nl; ++n, ++g) {
        const int pl = plane_of(n);
        const int i = p.x_start + pl - 1;
        const unsigned char *st = stages + (size_t)(g % kStages) * L::bytes;
        mbar_wait(full + (g % kStages), (uint32_t)((g / kStages) & 1));
        const R *sA = reinterpret_cast<const R *>(st + L::oA);
        const R *sB = reinterpret_cast<const R *>(st + L::oB);
        const R *sC = reinterpret_cast<const R *>(st + L::oC);
        V4<R> a_c, a_j, b_c, c_c, c_j;
        R a_k, b_k;
        if (PHASE == 1) {
            // opA = Hx rows j0-1.., cols k0-4.. ; opB = Hy cols k0-4.. ; opC = Hz rows j0-1..
            a_c = ld4(sA + (r + 1) * L::PA + c + 4);
            a_j = ld4(sA + r * L::PA + c + 4);
            b_c = ld4(sB + r * L::PA + c + 4);
            c_c = ld4(sC + (r + 1) * TZ + c);
            c_j = ld4(sC + r * TZ + c);
            a_k = __shfl_up_sync(0xffffffffu, a_c.w, 1);
            b_k = __shfl_up_sync(0xffffffffu, b_c.w, 1);
            if (c == 0 || lane == 0) {
                a_k = sA[(r + 1) * L::PA + c + 3];
                b_k = sB[r * L::PA + c + 3];
            }
        } else {
            // opA = Ex rows j0.., cols k0.. (+1 row, +4 cols) ; opB = Ey (+4 cols) ; opC = Ez (+1 row)
            a_c = ld4(sA + r * L::PA + c);
            a_j = ld4(sA + (r + 1) * L::PA + c);
            b_c = ld4(sB + r * L::PA + c);
            c_c = ld4(sC + r * TZ + c);
            c_j = ld4(sC + (r + 1) * TZ + c);
            a_k = __shfl_down_sync(0xffffffffu, a_c.x, 1);
            b_k = __shfl_down_sync(0xffffffffu, b_c.x, 1);
            if (c == TZ - 4 || lane == 31) {
                a_k = sA[r * L::PA + c + 4];
                b_k = sB[r * L::PA + c + 4];
            }
        }
        const int e = r * TZ + c;
        V4<R> f0 = ld4(reinterpret_cast<const R *>(st + L::oO0) + e);
        V4<R> f1 = ld4(reinterpret_cast<const R *>(st + L::oO1) + e);
        V4<R> f2 = ld4(reinterpret_cast<const R *>(st + L::oO2) + e);
        const Ids4 id0 = lds_ids4<IDT>(st + L::oI0, e), id1 = lds_ids4<IDT>(st + L::oI1, e), id2 = lds_ids4<IDT>(st + L::oI2, e);
        // this warp has taken what it needs from the stage.  No CTA-wide barrier: warps drift freely (the
        // first ncu capture showed barrier stalls on top); the refill is issued by thread 0 once all 8
        // warps have arrived on the stage's `empty` mbarrier (after its own compute, below).
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + (g % kStages));

        bool w0 = false, w1 = false, w2 = false;
        // one-sided differences along z (also needed by the z-slab hand-off below)
        V4<R> dA_dz, dB_dz;
        if (PHASE == 1) {
            dA_dz = {a_c.x - a_k, a_c.y - a_c.x, a_c.z - a_c.y, a_c.w - a_c.z};       // dHx/dz
            dB_dz = {b_c.x - b_k, b_c.y - b_c.x, b_c.z - b_c.y, b_c.w - b_c.z};       // dHy/dz
        } else {
            dA_dz = {a_c.y - a_c.x, a_c.z - a_c.y, a_c.w - a_c.z, a_k - a_c.w};       // dEx/dz
            dB_dz = {b_c.y - b_c.x, b_c.z - b_c.y, b_c.w - b_c.z, b_k - b_c.w};       // dEy/dz
        }
        if (fast_jk && i >= p.fast_i0 && i < p.fast_i1) {
            Coef4<R> q0, q1, q2, q3;
            coef4(scoef, id0, q0, q1, q2, q3);
            if (PHASE == 1) {
                f0.x = q0.a * f0.x + q0.by * (c_c.x - c_j.x) - q0.bz * dB_dz.x;
                f0.y = q1.a * f0.y + q1.by * (c_c.y - c_j.y) - q1.bz * dB_dz.y;
                f0.z = q2.a * f0.z + q2.by * (c_c.z - c_j.z) - q2.bz * dB_dz.z;
                f0.w = q3.a * f0.w + q3.by * (c_c.w - c_j.w) - q3.bz * dB_dz.w;
            } else {
                f0.x = q0.a * f0.x - q0.by * (c_j.x - c_c.x) + q0.bz * dB_dz.x;
                f0.y = q1.a * f0.y - q1.by * (c_j.y - c_c.y) + q1.bz * dB_dz.y;
                f0.z = q2.a * f0.z - q2.by * (c_j.z - c_c.z) + q2.bz * dB_dz.z;
                f0.w = q3.a * f0.w - q3.by * (c_j.w - c_c.w) + q3.bz * dB_dz.w;
            }
            coef4(scoef, id1, q0, q1, q2, q3);
            if (PHASE == 1) {
                f1.x = q0.a * f1.x + q0.bz * dA_dz.x - q0.bx * (c_c.x - qc.x);
                f1.y = q1.a * f1.y + q1.bz * dA_dz.y - q1.bx * (c_c.y - qc.y);
                f1.z = q2.a * f1.z + q2.bz * dA_dz.z - q2.bx * (c_c.z - qc.z);
                f1.w = q3.a * f1.w + q3.bz * dA_dz.w - q3.bx * (c_c.w - qc.w);
            } else {
                f1.x = q0.a * f1.x - q0.bz * dA_dz.x + q0.bx * (qc.x - c_c.x);
                f1.y = q1.a * f1.y - q1.bz * dA_dz.y + q1.bx * (qc.y - c_c.y);
                f1.z = q2.a * f1.z - q2.bz * dA_dz.z + q2.bx * (qc.z - c_c.z);
                f1.w = q3.a * f1.w - q3.bz * dA_dz.w + q3.bx * (qc.w - c_c.w);
            }
            coef4(scoef, id2, q0, q1, q2, q3);
            if (PHASE == 1) {
                f2.x = q0.a * f2.x + q0.bx * (b_c.x - qb.x) - q0.by * (a_c.x - a_j.x);
                f2.y = q1.a * f2.y + q1.bx * (b_c.y - qb.y) - q1.by * (a_c.y - a_j.y);
                f2.z = q2.a * f2.z + q2.bx * (b_c.z - qb.z) - q2.by * (a_c.z - a_j.z);
                f2.w = q3.a * f2.w + q3.bx * (b_c.w - qb.w) - q3.by * (a_c.w - a_j.w);
            } else {
                f2.x = q0.a * f2.x - q0.bx * (qb.x - b_c.x) + q0.by * (a_j.x - a_c.x);
                f2.y = q1.a * f2.y - q1.bx * (qb.y - b_c.y) + q1.by * (a_j.y - a_c.y);
                f2.z = q2.a * f2.z - q2.bx * (qb.z - b_c.z) + q2.by * (a_j.z - a_c.z);
                f2.w = q3.a * f2.w - q3.bx * (qb.w - b_c.w) + q3.by * (a_j.w - a_c.w);
            }
            w0 = w1 = w2 = true;
        } else if (any) {
            // E phase: backward differences (c - neighbour); H phase: forward (neighbour - c)
            V4<R> dA_dy, dB_dx, dC_dx, dC_dy;
            if (PHASE == 1) {
                dA_dy = {a_c.x - a_j.x, a_c.y - a_j.y, a_c.z - a_j.z, a_c.w - a_j.w};   // dHx/dy
                dB_dx = {b_c.x - qb.x, b_c.y - qb.y, b_c.z - qb.z, b_c.w - qb.w};         // dHy/dx
                dC_dx = {c_c.x - qc.x, c_c.y - qc.y, c_c.z - qc.z, c_c.w - qc.w};         // dHz/dx
                dC_dy = {c_c.x - c_j.x, c_c.y - c_j.y, c_c.z - c_j.z, c_c.w - c_j.w};   // dHz/dy
            } else {
                dA_dy = {a_j.x - a_c.x, a_j.y - a_c.y, a_j.z - a_c.z, a_j.w - a_c.w};   // dEx/dy
                dB_dx = {qb.x - b_c.x, qb.y - b_c.y, qb.z - b_c.z, qb.w - b_c.w};         // dEy/dx
                dC_dx = {qc.x - c_c.x, qc.y - c_c.y, qc.z - c_c.z, qc.w - c_c.w};         // dEz/dx
                dC_dy = {c_j.x - c_c.x, c_j.y - c_c.y, c_j.z - c_c.z, c_j.w - c_c.w};   // dEz/dy
            }
            const unsigned m0 = (i >= p.box[0].lo[0] && i < p.box[0].hi[0]) ? bx.kmask : 0u;
            const unsigned m1 = (i >= p.box[1].lo[0] && i < p.box[1].hi[0]) ? by.kmask : 0u;
            const unsigned m2 = (i >= p.box[2].lo[0] && i < p.box[2].hi[0]) ? bz.kmask : 0u;
            unsigned pm = 0;
#pragma unroll
            for (int s = 0; s < kMaxSlabs; ++s)
                if (((smask >> (4 * s)) & 0xfu) && i >= p.slab[s].lo[0] && i < p.slab[s].hi[0]) pm |= 1u << s;
            w0 = m0 != 0; w1 = m1 != 0; w2 = m2 != 0;
            // E phase: Ex = CA Ex + CBy dHz/dy - CBz dHy/dz ; Ey = CA Ey + CBz dHx/dz - CBx dHz/dx ; Ez = CA Ez + CBx dHy/dx - CBy dHx/dy
            // H phase: Hx = DA Hx - DBy dEz/dy + DBz dEy/dz ; Hy = DA Hy - DBz dEx/dz + DBx dEz/dx ; Hz = DA Hz - DBx dEy/dx + DBy dEx/dy
            if (m0) {
                Coef4<R> q0, q1, q2, q3;
                coef4(scoef, id0, q0, q1, q2, q3);
                if (PHASE == 1) {
                    f0.x = sel(m0, 0, q0.a * f0.x + q0.by * dC_dy.x - q0.bz * dB_dz.x, f0.x);
                    f0.y = sel(m0, 1, q1.a * f0.y + q1.by * dC_dy.y - q1.bz * dB_dz.y, f0.y);
                    f0.z = sel(m0, 2, q2.a * f0.z + q2.by * dC_dy.z - q2.bz * dB_dz.z, f0.z);
                    f0.w = sel(m0, 3, q3.a * f0.w + q3.by * dC_dy.w - q3.bz * dB_dz.w, f0.w);
                } else {
                    f0.x = sel(m0, 0, q0.a * f0.x - q0.by * dC_dy.x + q0.bz * dB_dz.x, f0.x);
                    f0.y = sel(m0, 1, q1.a * f0.y - q1.by * dC_dy.y + q1.bz * dB_dz.y, f0.y);
                    f0.z = sel(m0, 2, q2.a * f0.z - q2.by * dC_dy.z + q2.bz * dB_dz.z, f0.z);
                    f0.w = sel(m0, 3, q3.a * f0.w - q3.by * dC_dy.w + q3.bz * dB_dz.w, f0.w);
                }
            }
            if (m1) {
                Coef4<R> q0, q1, q2, q3;
                coef4(scoef, id1, q0, q1, q2, q3);
                if (PHASE == 1) {
                    f1.x = sel(m1, 0, q0.a * f1.x + q0.bz * dA_dz.x - q0.bx * dC_dx.x, f1.x);
                    f1.y = sel(m1, 1, q1.a * f1.y + q1.bz * dA_dz.y - q1.bx * dC_dx.y, f1.y);
                    f1.z = sel(m1, 2, q2.a * f1.z + q2.bz * dA_dz.z - q2.bx * dC_dx.z, f1.z);
                    f1.w = sel(m1, 3, q3.a * f1.w + q3.bz * dA_dz.w - q3.bx * dC_dx.w, f1.w);
                } else {
                    f1.x = sel(m1, 0, q0.a * f1.x - q0.bz * dA_dz.x + q0.bx * dC_dx.x, f1.x);
                    f1.y = sel(m1, 1, q1.a * f1.y - q1.bz * dA_dz.y + q1.bx * dC_dx.y, f1.y);
                    f1.z = sel(m1, 2, q2.a * f1.z - q2.bz * dA_dz.z + q2.bx * dC_dx.z, f1.z);
                    f1.w = sel(m1, 3, q3.a * f1.w - q3.bz * dA_dz.w + q3.bx * dC_dx.w, f1.w);
                }
            }
            if (m2) {
                Coef4<R> q0, q1, q2, q3;
                coef4(scoef, id2, q0, q1, q2, q3);
                if (PHASE == 1) {
                    f2.x = sel(m2, 0, q0.a * f2.x + q0.bx * dB_dx.x - q0.by * dA_dy.x, f2.x);
                    f2.y = sel(m2, 1, q1.a * f2.y + q1.bx * dB_dx.y - q1.by * dA_dy.y, f2.y);
                    f2.z = sel(m2, 2, q2.a * f2.z + q2.bx * dB_dx.z - q2.by * dA_dy.z, f2.z);
                    f2.w = sel(m2, 3, q3.a * f2.w + q3.bx * dB_dx.w - q3.by * dA_dy.w, f2.w);
                } else {
                    f2.x = sel(m2, 0, q0.a * f2.x - q0.bx * dB_dx.x + q0.by * dA_dy.x, f2.x);
                    f2.y = sel(m2, 1, q1.a * f2.y - q1.bx * dB_dx.y + q1.by * dA_dy.y, f2.y);
                    f2.z = sel(m2, 2, q2.a * f2.z - q2.bx * dB_dx.z + q2.by * dA_dy.z, f2.z);
                    f2.w = sel(m2, 3, q3.a * f2.w - q3.bx * dB_dx.w + q3.by * dA_dy.w, f2.w);
                }
            }
            if (pm) {
                for (int s = 0; s < p.nslabs; ++s) {
                    if (!((pm >> s) & 1u)) continue;
                    const SlabDev<R> &sl = p.slab[s];
                    const unsigned m = (smask >> (4 * s)) & 0xfu;
                    const int pos = sl.axis == 0 ? i : j;
                    const int depth = sl.minus ? (sl.dref - pos) : (pos - sl.dref);
                    const PmlCo<R> co = pml_load(p.form, p.order, sl, depth);
                    R *phi = sl.phi + ((long long)(i - sl.lo[0]) * sl.n1 + (j - sl.lo[1])) * sl.n2 + (k - sl.lo[2]);
                    if (PHASE == 1) {
                        if (sl.axis == 0) {  // Ey -= , dHz/dx ; Ez += , dHy/dx
                            pml_comp4(p.form, p.order, co, sl, phi, m, id1, ssrc, (R)-1, dC_dx, f1);
                            pml_comp4(p.form, p.order, co, sl, phi + sl.ostride, m, id2, ssrc, (R)1, dB_dx, f2);
                            w1 = w2 = true;
                        } else {  // Ex += , dHz/dy ; Ez -= , dHx/dy
                            pml_comp4(p.form, p.order, co, sl, phi, m, id0, ssrc, (R)1, dC_dy, f0);
                            pml_comp4(p.form, p.order, co, sl, phi + sl.ostride, m, id2, ssrc, (R)-1, dA_dy, f2);
                            w0 = w2 = true;
                        }
                    } else {
                        if (sl.axis == 0) {  // Hy += , dEz/dx ; Hz -= , dEy/dx
                            pml_comp4(p.form, p.order, co, sl, phi, m, id1, ssrc, (R)1, dC_dx, f1);
                            pml_comp4(p.form, p.order, co, sl, phi + sl.ostride, m, id2, ssrc, (R)-1, dB_dx, f2);
                            w1 = w2 = true;
                        } else {  // Hx -= , dEz/dy ; Hz += , dEx/dy
                            pml_comp4(p.form, p.order, co, sl, phi, m, id0, ssrc, (R)-1, dC_dy, f0);
                            pml_comp4(p.form, p.order, co, sl, phi + sl.ostride, m, id2, ssrc, (R)1, dA_dy, f2);
                            w0 = w2 = true;
                        }
                    }
                }
            }
        }

        if (any) {
            const long long off = (long long)pl * p.plane + eoff;
            if (w0) st4(F0 + off, f0);
            if (w1) st4(F1 + off, f1);
            if (w2) st4(F2 + off, f2);
        }
        if (tid == 0) produce();   // refill the ring one slot ahead of the oldest stage (waits for all warps' `empty` arrival)

        qb = b_c;
        qc = c_c;
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
