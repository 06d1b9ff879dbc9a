// gpb_kernels_coop.cuh -- a whole run of a SMALL grid in one cooperative launch (sm_100a).
//
// Grids that fit L2 (2-D models, 3-D up to ~150^3) are launch-bound on the kernel-per-half-step path: an iteration is 5-7 graph
// nodes of a few microseconds each (cylinder_Ascan_2D: 20 us per iteration for 12.6 k cells).  Here ONE kernel advances n
// iterations: co-resident CTAs (cooperative launch) walk the work items of the magnetic half-step, meet at a grid-wide barrier,
// walk the items of the electric half-step, meet again -- two barriers per iteration and nothing else.  The arithmetic is the
// register-vectorised kernels' (h4_body / e4_body of gpb_kernels_v4.cuh), so the results are bit-identical with every other
// kernel family.  What the separate small kernels did is done by the block that owns the cells, right after its item:
//   * z-slab PML corrections of the item's cells (k_pml_slabs), after the item's x / y-slab corrections;
//   * point sources on the item's cells (k_sources), in list order;
//   * receiver samples for the NEXT iteration (k_step_begin): after the electric half-step of an item the six field values of a
//     receiver cell are final (H since the barrier, E since this item), and so are the H values its currents Ix, Iy, Iz read.
// Not handled here (the solver keeps the graph path): transmission lines (their currents read H of neighbouring cells between
// the half-steps), snapshot iterations (the run is split around them).
#pragma once
#include <cooperative_groups.h>

#include "gpb_kernels_v4.cuh"

namespace gpb {

template <typename R>
struct CoopParams {
    PhaseParams<R> ph, pe;      // p0 = 0, p1 = nplanes, xchunk = planes per item
    PointParams<R> pp;
    int nrx;
    const int *rxc;
    R *rxs;
    int nsrc;
    const SrcDev<R> *srcs;
    unsigned zs_h, zs_e;        // z slabs (bit per slab) of each phase
    int gx, gy;                 // items: gx blocks of 4-cell vectors per plane x gy x-chunks
    int it0, n_iters;           // first iteration and count of this launch
    int *d_iter;                // [0] current, [1] next -- left as the kernel-per-step path expects them
};

// z-slab corrections of the cells of item (bx, by): one thread per cell, same routine as k_pml_slabs
template <typename R, typename IDT>
__device__ __forceinline__ void coop_zslabs(const PhaseParams<R> &p, int phase, unsigned zs, int bx, int by, int tid)
{
    const int l0 = p.p0 + by * p.xchunk, l1 = min(l0 + p.xchunk, p.p1);
    const long long e0 = (long long)bx * kThreadsV4 * 4;
    for (int s = 0; s < p.nslabs; ++s) {
        if (!((zs >> s) & 1u)) continue;
        const SlabDev<R> &sl = p.slab[s];
        for (int l = l0; l < l1; ++l) {
            const int i = p.x_start + l;
            if (i < sl.lo[0] || i >= sl.hi[0]) continue;
            for (int q = tid; q < kThreadsV4 * 4; q += kThreadsV4) {
                const long long e = e0 + q;
                if (e >= p.plane) break;
                const int j = (int)(e / p.pitch), k = (int)(e - (long long)j * p.pitch);
                if (j >= sl.lo[1] && j < sl.hi[1] && k >= sl.lo[2] && k < sl.hi[2]) pml_slab_cell<R, IDT>(p, phase, sl, i, j, k);
            }
        }
    }
}

template <typename R>
__device__ __forceinline__ bool coop_owns(const PhaseParams<R> &p, int bx, int by, int i, int j, int k)
{
    const int l = i - p.x_start, l0 = p.p0 + by * p.xchunk, l1 = min(l0 + p.xchunk, p.p1);
    const long long e = (long long)j * p.pitch + k, e0 = (long long)bx * kThreadsV4 * 4;
    return l >= l0 && l < l1 && e >= e0 && e < e0 + kThreadsV4 * 4;
}

// point sources of one phase on the cells of item (bx, by), applied by one thread in list order (k_sources, without lines)
template <typename R, typename IDT>
__device__ __forceinline__ void coop_sources(const CoopParams<R> &cp, const PhaseParams<R> &p, int phase, int bx, int by, int it)
{
    const PointParams<R> &pp = cp.pp;
    for (int pass = 0; pass < (phase == 0 ? 1 : 2); ++pass) {
        const int kind = phase == 0 ? 1 : (pass == 0 ? 2 : 0);   // magnetic dipoles | voltage sources, then Hertzian dipoles
        for (int s = 0; s < cp.nsrc; ++s) {
            const SrcDev<R> &sc = cp.srcs[s];
            if (sc.kind != kind || it < sc.it_first || it > sc.it_last || !coop_owns(p, bx, by, sc.i, sc.j, sc.k)) continue;
            const long long o = pt_off(pp, sc.i, sc.j, sc.k);
            if (kind == 1) {
                const unsigned m = ld_id<IDT>(pp.ID[3 + sc.pol], o);
                pp.F[3 + sc.pol][o] -= pp.srcH[m] * sc.wave[it] * sc.f2;
            } else if (kind == 2) {
                if (sc.hard) {
                    pp.F[sc.pol][o] = -sc.wave[it] / sc.f2;
                } else {
                    const unsigned m = ld_id<IDT>(pp.ID[sc.pol], o);
                    pp.F[sc.pol][o] -= pp.srcE[m] * sc.wave[it] * sc.f2;
                }
            } else {
                const unsigned m = ld_id<IDT>(pp.ID[sc.pol], o);
                pp.F[sc.pol][o] -= pp.srcE[m] * sc.wave[it] * sc.f1 * sc.f2;
            }
        }
    }
}

// receiver rows of iteration `it` for the receivers on the cells of item (bx, by) of the electric decomposition (bx < 0: all)
template <typename R>
__device__ __forceinline__ void coop_receivers(const CoopParams<R> &cp, int bx, int by, int it, int tid)
{
    const PointParams<R> &pp = cp.pp;
    if (it >= pp.iterations) return;
    for (int r = tid; r < cp.nrx; r += kThreadsV4) {
        const int i = cp.rxc[3 * r], j = cp.rxc[3 * r + 1], k = cp.rxc[3 * r + 2];
        if (!pt_owned(pp, i)) continue;
        if (bx >= 0 && !coop_owns(cp.pe, bx, by, i, j, k)) continue;
        const long long o = pt_off(pp, i, j, k);
        for (int c = 0; c < 6; ++c) cp.rxs[((long long)c * pp.iterations + it) * cp.nrx + r] = pp.F[c][o];
        for (int c = 0; c < 3; ++c) cp.rxs[((long long)(6 + c) * pp.iterations + it) * cp.nrx + r] = current_at(pp, c, i, j, k);
    }
}

// The two half-step bodies are CALLED, not inlined: inlined into one kernel they spilled 1.5 KB per thread even at 168
// registers; as functions each keeps the register allocation it has in its own kernel.
template <typename R, typename IDT>
__device__ __noinline__ void coop_h(const PhaseParams<R> &p, const Coef4<R> *coef, const R *src, int bx, int by, int tid)
{
    h4_body<R, IDT, false>(p, coef, src, bx, by, tid);
}
template <typename R, typename IDT, bool DISP>
__device__ __noinline__ void coop_e(const PhaseParams<R> &p, const Coef4<R> *coef, const R *src, int bx, int by, int tid)
{
    e4_body<R, IDT, DISP, false>(p, coef, src, bx, by, tid);
}

#ifndef GPB_COOP_MINBLOCKS
#define GPB_COOP_MINBLOCKS 4
#endif
template <typename R, typename IDT, bool DISP>
__global__ void __launch_bounds__(kThreadsV4, GPB_COOP_MINBLOCKS) k_run_coop(const __grid_constant__ CoopParams<R> cp)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nmat = cp.pe.nmat, tid = threadIdx.x;
    Coef4<R> *coefH = reinterpret_cast<Coef4<R> *>(smem_raw);
    Coef4<R> *coefE = coefH + nmat;
    R *srcH = reinterpret_cast<R *>(coefE + nmat);
    R *srcE = srcH + nmat;
    for (int m = tid; m < nmat; m += kThreadsV4) {
        coefH[m] = cp.ph.coef[m]; srcH[m] = cp.ph.src[m];
        coefE[m] = cp.pe.coef[m]; srcE[m] = cp.pe.src[m];
    }
    __syncthreads();
    const int items = cp.gx * cp.gy;
    // receiver samples of the first iteration of this launch: whatever ran before (the previous launch) is complete
    if (blockIdx.x == 0) coop_receivers(cp, -1, 0, cp.it0, tid);
    grid.sync();   // nobody changes H before the samples are taken
    for (int s = 0; s < cp.n_iters; ++s) {
        const int it = cp.it0 + s;
        for (int w = blockIdx.x; w < items; w += gridDim.x) {
            const int bx = w % cp.gx, by = w / cp.gx;
            coop_h<R, IDT>(cp.ph, coefH, srcH, bx, by, tid);
            if (cp.zs_h) {
                __syncthreads();
                coop_zslabs<R, IDT>(cp.ph, 0, cp.zs_h, bx, by, tid);
            }
            if (cp.nsrc) {
                __syncthreads();
                if (tid == 0) coop_sources<R, IDT>(cp, cp.ph, 0, bx, by, it);
            }
        }
        grid.sync();
        for (int w = blockIdx.x; w < items; w += gridDim.x) {
            const int bx = w % cp.gx, by = w / cp.gx;
            coop_e<R, IDT, DISP>(cp.pe, coefE, srcE, bx, by, tid);
            if (cp.zs_e) {
                __syncthreads();
                coop_zslabs<R, IDT>(cp.pe, 1, cp.zs_e, bx, by, tid);
            }
            __syncthreads();
            if (cp.nsrc && tid == 0) coop_sources<R, IDT>(cp, cp.pe, 1, bx, by, it);
            if (cp.nrx && s + 1 < cp.n_iters) {   // the next iteration's samples of the receivers on these cells
                __syncthreads();
                coop_receivers(cp, bx, by, it + 1, tid);
            }
        }
        grid.sync();
    }
    if (blockIdx.x == 0 && tid == 0) {
        cp.d_iter[0] = cp.it0 + cp.n_iters - 1;
        cp.d_iter[1] = cp.it0 + cp.n_iters;
    }
}

}  // namespace gpb
