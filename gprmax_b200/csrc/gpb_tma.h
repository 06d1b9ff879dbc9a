// gpb_tma.h -- interface between the host side (gpb_core.cu) and the translation units that hold the TMA-staged
// E/H kernels (gpb_tma_inst.cu, compiled once per (float type, PML variant) -- see gprmax_b200/build.py).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <string>

#include "gpb_kernels.cuh"

namespace gpb {

// The three fields a phase reads (operands), the three it updates (own) and their three ID arrays are each ONE
// allocation [3][planes][rows][pitch], so a 4-D box fetches a whole triple with a single TMA instruction.
struct TmaMaps4 {
    CUtensorMap op;    // operands, box (TZ+4) x (TY+1) x 1 x 3 (E phase: Hx,Hy,Hz from (k0-4, j0-1); H phase: Ex,Ey,Ez from (k0, j0))
    CUtensorMap opx;   // x-neighbour plane at the start of an item: components B and C only, same box x 2
    CUtensorMap own;   // fields being updated, box TZ x TY x 1 x 3
    CUtensorMap id;    // their material IDs,   box TZ x TY x 1 x 3 (elements of IDT)
};

template <typename R>
struct TmaLaunch {
    PhaseParams<R> p;         // phase parameters; p0/p1/xchunk/persist/fast_i*/zfused set by the caller, tmax/pf_depth by the launcher
    const TmaMaps4 *maps;
    int phase;                // 0 magnetic, 1 electric
    int ty, tz, stages, pw;   // tile, ring depth, dedicated producer warp
    int idbytes;              // 1, 2, 4
    int pf_max;               // upper bound on the Phi prefetch distance (GPB_TMA_PF)
    int disp;                 // electric half-step of a dispersive model: 1 complex T, 2 real T (0: none)
    int t_max;                // upper bound on the T prefetch distance (GPB_TMA_TPF)
    int nosplit;              // no half-size items at the end of a persistent launch (GPB_TMA_NOSPLIT)
    int sm_count;
    int *sched;               // [2] work-item scheduler state
    cudaStream_t stream;
    int pdl;                  // launch with the programmatic-stream-serialization attribute (gpb_kernels.cuh: launch_pdl)
};

// PV = 2 * formulation + order - 1.  Returns 0, or 1 with *err filled.
template <typename R, int PV>
int tma_launch(const TmaLaunch<R> &a, std::string *err);

// Both half-steps of one iteration in one launch (gpb_kernels_pair.cuh): default 14 x 64 tile with the producer warp only.
template <typename R>
struct TmaLaunchPair {
    PhaseParams<R> ph, pe;    // magnetic / electric phase parameters (p0/p1/xchunk/fast_i*/zfused/znocoop/progress set by the caller)
    const TmaMaps4 *maps_h, *maps_e;
    int idbytes, pf_max, t_max;
    int disp;                 // 0 none, 1 complex T, 2 real T
    int sm_count;
    int *sched;
    cudaStream_t stream;
};
template <typename R, int PV>
int tma_launch_pair(const TmaLaunchPair<R> &a, std::string *err);

}  // namespace gpb
