// gpb_kernels_pair.cuh -- H and E half-steps of ONE iteration in ONE persistent launch (sm_100a).
//
// Why: with a kernel per half-step every field array crosses the HBM pins twice per iteration -- the E half-step reads the H
// planes the H half-step wrote a full sweep (hundreds of MB) earlier, and re-reads the E planes the H half-step read as
// operands.  Here the persistent CTAs of k_update_tma pull work items of BOTH half-steps from one queue, ordered so that the E
// items of an x chunk follow the H items of the chunk behind it by about two chunks of planes:
//
//      H(0) .. H(L)  E(0)  H(L+1)  E(1)  H(L+2)  E(2)  ...  H(C-1)  E(C-1-L) .. E(C-1)     (each group = all tiles of one chunk)
//
// so an E item finds its H operands and its own E planes in the 126 MB L2 (written / read a few tens of MB ago).  It is valid
// because E(i) needs H(i-1), H(i), and the H update of later chunks only reads E planes the E update of this chunk does not
// write.  Correctness does not rest on the order alone: every consumer warp publishes a finished H item (fence + atomic on a
// per-chunk counter) and the producer of a CTA loads an E item of chunk c only after the counters of chunks c-1 and c are
// complete (acquire + fence.proxy.async, because the loads that follow go through the TMA unit).  Items of both kinds share the
// TMA ring: the stage layout is the same for both half-steps.  Any CTA takes any item, so there is no static split of the SMs
// between the half-steps.  Same arithmetic, same bits as the two kernels one after the other: the consumer side of an item is
// the same source (gpb_tma_item.inc), compiled for both phases.
#pragma once
#include <type_traits>

#include "gpb_kernels_tma.cuh"

namespace gpb {

// group g of the unified queue -> (phase, chunk); C = number of x chunks (2 C groups), L = lag: E(c) directly follows H(c + L)
//      H(0) .. H(L)  E(0)  H(L+1)  E(1)  ...  H(C-1)  E(C-1-L)  E(C-L) .. E(C-1)
__device__ __forceinline__ void pair_group(int g, int C, int L, int &phase, int &chunk)
{
    L = min(L, C - 1);
    if (g <= L) { phase = 0; chunk = g; return; }
    const int t = g - (L + 1), rem = C - 1 - L;
    if (t < 2 * rem) {
        phase = (t & 1) ? 0 : 1;
        chunk = (t & 1) ? L + 1 + (t >> 1) : (t >> 1);
    } else {
        phase = 1;
        chunk = rem + (t - 2 * rem);
    }
}

template <typename R, typename IDT, int TY, int TZ, int kStages, int PV, int DISP>
__global__ void __launch_bounds__(TY * TZ / 4 + 32, (sizeof(R) == 4 ? GPB_TMA_CTAS : 1))
k_update_pair(const __grid_constant__ PhaseParams<R> ph, const __grid_constant__ PhaseParams<R> pe, const __grid_constant__ TmaMaps4 mh,
              const __grid_constant__ TmaMaps4 me, int tiles_k, int tiles, int nchunks, int *sched)
{
    constexpr int PW = 1;
    constexpr int kTmaThreads = TY * TZ / 4;
    static_assert(kTmaThreads % 32 == 0 && kTmaThreads + 32 <= 256, "tile shape");
    constexpr int PFORM = PV >> 1, PORDER = (PV & 1) + 1;
    constexpr int KTW = DISP == 1 ? 2 : 1;
    using L = StageLayout<R, IDT, TY, TZ>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *empty = full + kStages;
    volatile int *ring = reinterpret_cast<volatile int *>(empty + kStages);
    // shared memory: barriers | H coefficient rows | E coefficient rows | H PML tables | E PML tables | Phi prefetch slots |
    //                dispersive coefficient triples | T prefetch slots | stage ring
    const int nmat = pe.nmat;
    const int coef_bytes = (int)((nmat * (sizeof(Coef4<R>) + sizeof(R)) + 127) / 128 * 128);
    Coef4<R> *scoefH = reinterpret_cast<Coef4<R> *>(smem_raw + 128);
    R *ssrcH = reinterpret_cast<R *>(scoefH + nmat);
    Coef4<R> *scoefE = reinterpret_cast<Coef4<R> *>(smem_raw + 128 + coef_bytes);
    R *ssrcE = reinterpret_cast<R *>(scoefE + nmat);
    const int tabH_bytes = (int)((ph.nslabs * 4 * PORDER * ph.tmax * sizeof(R) + 127) / 128 * 128);
    const int tabE_bytes = (int)((pe.nslabs * 4 * PORDER * pe.tmax * sizeof(R) + 127) / 128 * 128);
    R *stabH = reinterpret_cast<R *>(smem_raw + 128 + 2 * coef_bytes);
    R *stabE = reinterpret_cast<R *>(smem_raw + 128 + 2 * coef_bytes + tabH_bytes);
    const int off_pf = 128 + 2 * coef_bytes + tabH_bytes + tabE_bytes;
    V4<R> *spf = reinterpret_cast<V4<R> *>(smem_raw + off_pf);
    const int pf_bytes = max(ph.pf_depth, pe.pf_depth) * 2 * PORDER * kTmaThreads * (int)sizeof(V4<R>);
    R *sdc = reinterpret_cast<R *>(smem_raw + off_pf + pf_bytes);
    const int dc_bytes = DISP ? (int)((nmat * pe.maxpoles * 3 * KTW * sizeof(R) + 127) / 128 * 128) : 0;
    V4<R> *stf = reinterpret_cast<V4<R> *>(smem_raw + off_pf + pf_bytes + dc_bytes);
    const int tslot = DISP ? 3 * pe.maxpoles * KTW * kTmaThreads : 0;
    const int tf_bytes = DISP ? pe.t_depth * tslot * (int)sizeof(V4<R>) : 0;
    unsigned char *stages = smem_raw + off_pf + pf_bytes + dc_bytes + tf_bytes;

    const int tid = threadIdx.x, lane = tid & 31;
    const int W = 2 * nchunks * tiles;
    constexpr int nsplit = 0;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, kTmaThreads / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid == 32) {
        tma_prefetch_desc(&mh.op); tma_prefetch_desc(&mh.opx); tma_prefetch_desc(&mh.own); tma_prefetch_desc(&mh.id);
        tma_prefetch_desc(&me.op); tma_prefetch_desc(&me.opx); tma_prefetch_desc(&me.own); tma_prefetch_desc(&me.id);
    }
    for (int m = tid; m < nmat; m += kTmaThreads + 32) {
        scoefH[m] = ph.coef[m]; ssrcH[m] = ph.src[m];
        scoefE[m] = pe.coef[m]; ssrcE[m] = pe.src[m];
    }
    if (DISP) {
        const R *src = reinterpret_cast<const R *>(pe.dcoef);
        for (int m = tid; m < nmat * pe.maxpoles * 3 * KTW; m += kTmaThreads + 32) sdc[m] = src[m];
    }
    for (int phs = 0; phs < 2; ++phs) {
        const PhaseParams<R> &pp = phs ? pe : ph;
        R *stab = phs ? stabE : stabH;
        for (int s = 0; s < pp.nslabs; ++s) {
            const SlabDev<R> &sl = pp.slab[s];
            for (int m = tid; m < 4 * PORDER * sl.t; m += kTmaThreads + 32) {
                const int q = m / (PORDER * sl.t), o = (m / sl.t) % PORDER, dd = m % sl.t;
                const R *src = q == 0 ? sl.RA : (q == 1 ? sl.RB : (q == 2 ? sl.RE : sl.RF));
                stab[((s * 4 + q) * PORDER + o) * pp.tmax + dd] = src[o * sl.t + dd];
            }
        }
    }
    __syncthreads();

    // ---------------- producer (dedicated warp, one lane): walks the CTA's slot sequence through items of both phases
    int p_item = -1, p_n = -1, p_q = 0, p_g = 0, p_phase = 0;
    int p_j0 = 0, p_k0 = 0, p_l0 = 0, p_l1 = 0;
    bool p_done = false;
    auto fetch = [&]() {
        const int w = atomicAdd(sched, 1);
        p_item = -1;
        if (w < W) {
            int chunk;
            pair_group(w / tiles, nchunks, ph.pair_lag, p_phase, chunk);
            const int tile = w % tiles;
            p_item = tile | (p_phase << 19) | (chunk << 20);
            p_k0 = (tile % tiles_k) * TZ;
            p_j0 = (tile / tiles_k) * TY;
            chunk_range(chunk, nchunks, 0, ph.xchunk, ph.p0, ph.p1, p_l0, p_l1, 1);
            if (p_phase == 1) {
                // the H items of this chunk and of the chunk in front of it must be complete (published with fence + atomic by
                // the consumer warps that finished them): acquire, then order the TMA (async proxy) loads behind it
                unsigned long long t0;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                for (int cc = chunk > 0 ? chunk - 1 : chunk; cc <= chunk; ++cc) {
                    unsigned v;
                    for (;;) {
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ph.progress + cc) : "memory");
                        if (v >= ph.prog_need) break;
                        __nanosleep(64);
                        unsigned long long t1;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                        if (t1 - t0 > ph.prog_timeout_ns) { atomicOr(ph.prog_flags, 1u); break; }
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
        if (p_item < 0) {
            ring[p_q & 3] = -1;
            __threadfence_block();
            mbar_arrive(bar);
            p_done = true;
            return;
        }
        const TmaMaps4 &maps = p_phase ? me : mh;
        const int ck = p_phase ? p_k0 - 4 : p_k0, cj = p_phase ? p_j0 - 1 : p_j0;
        if (p_n < 0) {
            ring[p_q & 3] = p_item;
            __threadfence_block();
            mbar_expect_tx(bar, (uint32_t)L::tx_x);
            tma_load_4d(st + L::oOwn, &maps.opx, bar, ck, cj, p_phase ? p_l0 : p_l1 + 1, 1);
        } else {
            const int pl = p_phase ? (p_l0 + p_n + 1) : (p_l1 - 1 - p_n + 1);
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
    if (tid >= kTmaThreads) {
        if (lane == 0) {
            fetch();
            while (!p_done) produce();
        }
        return;
    }

    // ---------------- consumers: one item at a time, whichever phase it belongs to
    const int r = tid / (TZ / 4), c = (tid % (TZ / 4)) * 4;
    const int e = r * TZ + c;
    int g = 0;
    auto consume = [&](auto phase_tag, int item) {
        constexpr int PHASE = decltype(phase_tag)::value;
        const PhaseParams<R> &p = PHASE ? pe : ph;
        Coef4<R> *scoef = PHASE ? scoefE : scoefH;
        R *ssrc = PHASE ? ssrcE : ssrcH;
        R *stab = PHASE ? stabE : stabH;
#include "gpb_tma_item.inc"
    };
    for (int q = 0;; ++q) {
        mbar_wait(full + (g % kStages), (uint32_t)((g / kStages) & 1));
        const int item = ring[q & 3];
        if (item < 0) break;
        if ((item >> 19) & 1) consume(std::integral_constant<int, 1>{}, item);
        else consume(std::integral_constant<int, 0>{}, item);
    }

    if (tid == 0) {   // the last CTA to finish re-arms the scheduler for the next launch
        __threadfence();
        if (atomicAdd(sched + 1, 1) == (int)gridDim.x - 1) {
            sched[0] = 0;
            sched[1] = 0;
            __threadfence();
        }
    }
}

}  // namespace gpb
