// gpb_kernels_v4.cuh -- vectorised E/H half-step kernels (non-dispersive fast path).
//
// Same arithmetic as k_update_e / k_update_h in gpb_kernels.cuh (which remain the generic path for
// dispersive materials), restructured after the first ncu capture (profiles/r1a_*): the scalar
// kernels executed ~400 instructions per cell (index decode, 9 box tests and 64-bit addressing per
// cell) and were issue-bound at 29 % of DRAM bandwidth.  Here
//   * one thread owns 4 consecutive z cells: every field/ID/Phi access is one 128-bit (fp32) load,
//     and the index / box / slab logic is paid once per 4 cells;
//   * the k+-1 operands come from the neighbouring lane by warp shuffle (one scalar load on the
//     warp-edge lane only);
//   * the march along x keeps the i+1 (H phase) / i-1 (E phase) operand planes in a register queue;
//   * x- and y-slab PML corrections are warp-uniform (depth depends on i or j only) and vectorised;
//     z-slab corrections, which would diverge in every warp (10 cells at both ends of every row), run
//     in their own small kernel (k_pml_slabs) right after the main one.
#pragma once
#include "gpb_kernels.cuh"

namespace gpb {

#ifndef GPB_V4_THREADS
#define GPB_V4_THREADS 128
#endif
#ifndef GPB_V4_MINBLOCKS
#define GPB_V4_MINBLOCKS 4
#endif
#ifndef GPB_V4_UNROLL
#define GPB_V4_UNROLL 1
#endif
constexpr int kThreadsV4 = GPB_V4_THREADS;
constexpr int kV4Unroll = GPB_V4_UNROLL;

// T * or T *__restrict__
template <bool RESTRICT, typename T>
struct MaybeRestrict {
    typedef T *type;
};
template <typename T>
struct MaybeRestrict<true, T> {
    typedef T *__restrict__ type;
};

template <typename R>
struct alignas(4 * sizeof(R)) V4 {
    R x, y, z, w;
};

template <typename R>
__device__ __forceinline__ V4<R> ld4(const R *p)
{
    return *reinterpret_cast<const V4<R> *>(p);
}
template <typename R>
__device__ __forceinline__ void st4(R *p, const V4<R> &v)
{
    *reinterpret_cast<V4<R> *>(p) = v;
}

struct Ids4 {
    unsigned a, b, c, d;
};
template <typename IDT>
__device__ __forceinline__ Ids4 ld_ids4(const void *base, long long off);
template <>
__device__ __forceinline__ Ids4 ld_ids4<uint8_t>(const void *base, long long off)
{
    const unsigned v = __ldg(reinterpret_cast<const unsigned *>(reinterpret_cast<const uint8_t *>(base) + off));
    return {v & 0xffu, (v >> 8) & 0xffu, (v >> 16) & 0xffu, v >> 24};
}
template <>
__device__ __forceinline__ Ids4 ld_ids4<uint16_t>(const void *base, long long off)
{
    const uint2 v = __ldg(reinterpret_cast<const uint2 *>(reinterpret_cast<const uint16_t *>(base) + off));
    return {v.x & 0xffffu, v.x >> 16, v.y & 0xffffu, v.y >> 16};
}
template <>
__device__ __forceinline__ Ids4 ld_ids4<uint32_t>(const void *base, long long off)
{
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<const uint32_t *>(base) + off));
    return {v.x, v.y, v.z, v.w};
}

// per-thread, loop-invariant description of where its 4 cells sit relative to a box
struct JK4 {
    bool j_in;
    unsigned kmask;  // bit q set: cell k+q inside the box's k range
};
__device__ __forceinline__ JK4 jk4_of(const int *lo, const int *hi, int j, int k)
{
    JK4 r;
    r.j_in = j >= lo[1] && j < hi[1];
    // cells k + q, q = 0..3, inside [lo2, hi2): bits [max(lo2 - k, 0), min(hi2 - k, 4))
    const int q0 = max(lo[2] - k, 0), q1 = min(hi[2] - k, 4);
    r.kmask = (r.j_in && q1 > q0) ? (((1u << q1) - 1u) & ~((1u << q0) - 1u)) : 0u;
    return r;
}

template <typename R>
__device__ __forceinline__ R sel(unsigned mask, int q, R a, R b)
{
    return (mask >> q) & 1u ? a : b;
}

// One vectorised PML component: F (4 cells) -= / += src[id] * term(dF / d), Phi advanced.  dF / d is formed as dF * (1/d): the
// IEEE division's slow path is taken for denormal numerators, which is what most of a PML holds (SlabDev::inv_d); the generic
// scalar kernels (gpb_kernels.cuh) keep the reference's division.
// sign = +1 / -1.  `m` = lanes (cells) inside the slab.
template <typename R>
__device__ __forceinline__ void pml_comp4(int form, int order, const PmlCo<R> &co, const SlabDev<R> &sl, R *phi, unsigned m,
                                          const Ids4 &id, const R *src, R sign, const V4<R> &dF, V4<R> &F)
{
    V4<R> p0 = ld4(phi), p1 = p0;
    if (order == 2) p1 = ld4(phi + 2 * sl.ostride);
    V4<R> q0 = p0, q1 = p1;
    const R tx = pml_apply(form, order, co, mul_(dF.x, sl.inv_d), q0.x, q1.x);
    const R ty = pml_apply(form, order, co, mul_(dF.y, sl.inv_d), q0.y, q1.y);
    const R tz = pml_apply(form, order, co, mul_(dF.z, sl.inv_d), q0.z, q1.z);
    const R tw = pml_apply(form, order, co, mul_(dF.w, sl.inv_d), q0.w, q1.w);
    if (m & 1u) { F.x = fma_(sign, mul_(src[id.a], tx), F.x); p0.x = q0.x; p1.x = q1.x; }
    if (m & 2u) { F.y = fma_(sign, mul_(src[id.b], ty), F.y); p0.y = q0.y; p1.y = q1.y; }
    if (m & 4u) { F.z = fma_(sign, mul_(src[id.c], tz), F.z); p0.z = q0.z; p1.z = q1.z; }
    if (m & 8u) { F.w = fma_(sign, mul_(src[id.d], tw), F.w); p0.w = q0.w; p1.w = q1.w; }
    st4(phi, p0);
    if (order == 2) st4(phi + 2 * sl.ostride, p1);
}

template <typename R>
__device__ __forceinline__ void coef4(const Coef4<R> *coef, const Ids4 &id, Coef4<R> &c0, Coef4<R> &c1, Coef4<R> &c2, Coef4<R> &c3)
{
    c0 = coef[id.a];
    if (id.a == id.b && id.a == id.c && id.a == id.d) {
        c1 = c2 = c3 = c0;
    } else {
        c1 = coef[id.b];
        c2 = coef[id.c];
        c3 = coef[id.d];
    }
}

// ------------------------------------------------------------------------------------------
// Magnetic half-step, 4 z cells per thread.  Arithmetic: fields_updates_ext.pyx:352-412 and
// pml_updates_magnetic_*_ext.pyx (x and y slabs).
// ------------------------------------------------------------------------------------------
// (blk, chunk) = block of 4-cell vectors within a plane / x chunk; NC: the operand fields are read-only for the whole kernel, so
// their pointers may be restrict-qualified (loads hoisted above stores, non-coherent path).  The cooperative whole-run kernel
// (gpb_kernels_coop.cuh) runs both half-steps in one launch and instantiates NC = false.
template <typename R, typename IDT, bool NC>
__device__ __forceinline__ void h4_body(const PhaseParams<R> &p, const Coef4<R> *coef, const R *srcm, int blk, int chunk, int tid)
{
    const long long idx4 = (long long)blk * kThreadsV4 + tid;
    const long long e0 = idx4 * 4;
    const bool valid = e0 < p.plane;
    const long long eoff = valid ? e0 : 0;
    const int j = (int)(eoff / p.pitch);
    const int k = (int)(eoff - (long long)j * p.pitch);
    const int lane = tid & 31;
    const int l0 = p.p0 + chunk * p.xchunk;
    const int l1 = min(l0 + p.xchunk, p.p1);
    if (l0 >= l1) return;
    const JK4 bx = jk4_of(p.box[0].lo, p.box[0].hi, j, k), by = jk4_of(p.box[1].lo, p.box[1].hi, j, k), bz = jk4_of(p.box[2].lo, p.box[2].hi, j, k);
    // x / y slabs: which of my cells lie in the slab's (j,k) footprint (z slabs are handled by k_pml_slabs)
    unsigned smask = 0;  // 4 bits per slab
#pragma unroll
    for (int s = 0; s < kMaxSlabs; ++s)
        if (s < p.nslabs && p.slab[s].axis != 2 && valid) smask |= jk4_of(p.slab[s].lo, p.slab[s].hi, j, k).kmask << (4 * s);
    const bool any = valid && ((bx.kmask | by.kmask | bz.kmask) != 0u || smask != 0u);

    // E is read-only in this phase and never aliases H: tell the compiler so that the loads of the
    // next plane can be hoisted above the stores of this one
    typename MaybeRestrict<NC, const R>::type Ex = p.Ex, Ey = p.Ey, Ez = p.Ez;
    typename MaybeRestrict<NC, R>::type Hx = p.Hx, Hy = p.Hy, Hz = p.Hz;
    long long off = (long long)(l0 + 1) * p.plane + eoff;
    V4<R> ey_c = ld4(Ey + off), ez_c = ld4(Ez + off);
#pragma unroll kV4Unroll
    for (int l = l0; l < l1; ++l, off += p.plane) {
        const int i = p.x_start + l;
        const V4<R> ey_n = ld4(Ey + off + p.plane), ez_n = ld4(Ez + off + p.plane);
        const V4<R> ex_c = ld4(Ex + off);
        const V4<R> ex_j = ld4(Ex + off + p.pitch), ez_j = ld4(Ez + off + p.pitch);
        // k+1 operands: next lane's .x (scalar load on the warp-edge lane)
        R ex_k4 = __shfl_down_sync(0xffffffffu, ex_c.x, 1), ey_k4 = __shfl_down_sync(0xffffffffu, ey_c.x, 1);
        if (lane == 31) {
            ex_k4 = Ex[off + 4];
            ey_k4 = Ey[off + 4];
        }
        if (any) {
            const V4<R> dEz_dy = {ez_j.x - ez_c.x, ez_j.y - ez_c.y, ez_j.z - ez_c.z, ez_j.w - ez_c.w};
            const V4<R> dEy_dz = {ey_c.y - ey_c.x, ey_c.z - ey_c.y, ey_c.w - ey_c.z, ey_k4 - ey_c.w};
            const V4<R> dEx_dz = {ex_c.y - ex_c.x, ex_c.z - ex_c.y, ex_c.w - ex_c.z, ex_k4 - ex_c.w};
            const V4<R> dEz_dx = {ez_n.x - ez_c.x, ez_n.y - ez_c.y, ez_n.z - ez_c.z, ez_n.w - ez_c.w};
            const V4<R> dEy_dx = {ey_n.x - ey_c.x, ey_n.y - ey_c.y, ey_n.z - ey_c.z, ey_n.w - ey_c.w};
            const V4<R> dEx_dy = {ex_j.x - ex_c.x, ex_j.y - ex_c.y, ex_j.z - ex_c.z, ex_j.w - ex_c.w};
            const unsigned mx = (i >= p.box[0].lo[0] && i < p.box[0].hi[0]) ? bx.kmask : 0u;
            const unsigned my = (i >= p.box[1].lo[0] && i < p.box[1].hi[0]) ? by.kmask : 0u;
            const unsigned mz = (i >= p.box[2].lo[0] && i < p.box[2].hi[0]) ? bz.kmask : 0u;
            unsigned pm = 0;  // slabs active on this plane for this thread
#pragma unroll
            for (int s = 0; s < kMaxSlabs; ++s)
                if (((smask >> (4 * s)) & 0xfu) && i >= p.slab[s].lo[0] && i < p.slab[s].hi[0]) pm |= 1u << s;
            bool wx = mx != 0, wy = my != 0, wz = mz != 0;
            V4<R> hx, hy, hz;
            Ids4 idx_, idy_, idz_;
            const bool needx = wx || (pm != 0), needy = wy || (pm != 0), needz = wz || (pm != 0);
            if (needx) {
                hx = ld4(Hx + off);
                idx_ = ld_ids4<IDT>(p.ID[0], off);
            }
            if (needy) {
                hy = ld4(Hy + off);
                idy_ = ld_ids4<IDT>(p.ID[1], off);
            }
            if (needz) {
                hz = ld4(Hz + off);
                idz_ = ld_ids4<IDT>(p.ID[2], off);
            }
            if (mx) {
                Coef4<R> c0, c1, c2, c3;
                coef4(coef, idx_, c0, c1, c2, c3);
                hx.x = sel(mx, 0, upd3(c0.a, hx.x, -c0.by, dEz_dy.x, c0.bz, dEy_dz.x), hx.x);
                hx.y = sel(mx, 1, upd3(c1.a, hx.y, -c1.by, dEz_dy.y, c1.bz, dEy_dz.y), hx.y);
                hx.z = sel(mx, 2, upd3(c2.a, hx.z, -c2.by, dEz_dy.z, c2.bz, dEy_dz.z), hx.z);
                hx.w = sel(mx, 3, upd3(c3.a, hx.w, -c3.by, dEz_dy.w, c3.bz, dEy_dz.w), hx.w);
            }
            if (my) {
                Coef4<R> c0, c1, c2, c3;
                coef4(coef, idy_, c0, c1, c2, c3);
                hy.x = sel(my, 0, upd3(c0.a, hy.x, -c0.bz, dEx_dz.x, c0.bx, dEz_dx.x), hy.x);
                hy.y = sel(my, 1, upd3(c1.a, hy.y, -c1.bz, dEx_dz.y, c1.bx, dEz_dx.y), hy.y);
                hy.z = sel(my, 2, upd3(c2.a, hy.z, -c2.bz, dEx_dz.z, c2.bx, dEz_dx.z), hy.z);
                hy.w = sel(my, 3, upd3(c3.a, hy.w, -c3.bz, dEx_dz.w, c3.bx, dEz_dx.w), hy.w);
            }
            if (mz) {
                Coef4<R> c0, c1, c2, c3;
                coef4(coef, idz_, c0, c1, c2, c3);
                hz.x = sel(mz, 0, upd3(c0.a, hz.x, -c0.bx, dEy_dx.x, c0.by, dEx_dy.x), hz.x);
                hz.y = sel(mz, 1, upd3(c1.a, hz.y, -c1.bx, dEy_dx.y, c1.by, dEx_dy.y), hz.y);
                hz.z = sel(mz, 2, upd3(c2.a, hz.z, -c2.bx, dEy_dx.z, c2.by, dEx_dy.z), hz.z);
                hz.w = sel(mz, 3, upd3(c3.a, hz.w, -c3.bx, dEy_dx.w, c3.by, dEx_dy.w), hz.w);
            }
            if (pm) {
                for (int s = 0; s < p.nslabs; ++s) {
                    if (!((pm >> s) & 1u)) continue;
                    const SlabDev<R> &sl = p.slab[s];
                    const int pos = sl.axis == 0 ? i : j;
                    const int depth = sl.minus ? (sl.dref - pos) : (pos - sl.dref);
                    const PmlCo<R> co = pml_load(p.form, p.order, sl, depth);
                    R *phi = sl.phi + ((long long)(i - sl.lo[0]) * sl.n1 + (j - sl.lo[1])) * sl.n2 + (k - sl.ko);
                    const unsigned m = (smask >> (4 * s)) & 0xfu;
                    if (sl.axis == 0) {  // Hy += , dEz/dx ; Hz -= , dEy/dx
                        pml_comp4(p.form, p.order, co, sl, phi, m, idy_, srcm, (R)1, dEz_dx, hy);
                        pml_comp4(p.form, p.order, co, sl, phi + sl.ostride, m, idz_, srcm, (R)-1, dEy_dx, hz);
                        wy = wz = true;
                    } else {  // Hx -= , dEz/dy ; Hz += , dEx/dy
                        pml_comp4(p.form, p.order, co, sl, phi, m, idx_, srcm, (R)-1, dEz_dy, hx);
                        pml_comp4(p.form, p.order, co, sl, phi + sl.ostride, m, idz_, srcm, (R)1, dEx_dy, hz);
                        wx = wz = true;
                    }
                }
            }
            if (wx) st4(Hx + off, hx);
            if (wy) st4(Hy + off, hy);
            if (wz) st4(Hz + off, hz);
        }
        ey_c = ey_n;
        ez_c = ez_n;
    }
}

template <typename R, typename IDT, bool TABSMEM>
__global__ void __launch_bounds__(kThreadsV4, GPB_V4_MINBLOCKS) k_update_h4(const PhaseParams<R> p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pdl_launch_dependents();
    const Coef4<R> *coef = p.coef;
    const R *srcm = p.src;
    if (TABSMEM) {
        Coef4<R> *scoef = reinterpret_cast<Coef4<R> *>(smem_raw);
        R *ssrc = reinterpret_cast<R *>(scoef + p.nmat);
        stage_coefs(p, scoef, ssrc);
        coef = scoef;
        srcm = ssrc;
    }
    pdl_wait();   // the coefficient tables staged above are constant during a run
    h4_body<R, IDT, true>(p, coef, srcm, (int)blockIdx.x, (int)blockIdx.y, (int)threadIdx.x);
}

// Dispersive sum for 4 cells of one component: part B of the previous step folded into part A of this one, the per-cell
// arithmetic of gpb_kernels.cuh (disp_cell_c / disp_cell_r).  Complex T: complex[pole][cells], 4 cells = two 128-bit (fp32)
// accesses; real T (Debye media, p.treal): R[pole][cells], one access.  Cells outside the update box keep their T.
template <typename R>
__device__ __forceinline__ void disp4(const PhaseParams<R> &p, int comp, const Ids4 &id, long long off, unsigned m, const V4<R> &e, float (&ph)[4])
{
    ph[0] = ph[1] = ph[2] = ph[3] = 0;
    if (p.treal) {
        R *T = reinterpret_cast<R *>(p.T[comp]) + off;
        const R *dc = reinterpret_cast<const R *>(p.dcoef);
        for (int q = 0; q < p.maxpoles; ++q, T += p.tstride) {
            V4<R> t = ld4(T);
            if (m & 1u) disp_cell_r(dc + ((long long)id.a * p.maxpoles + q) * 3, e.x, t.x, ph[0]);
            if (m & 2u) disp_cell_r(dc + ((long long)id.b * p.maxpoles + q) * 3, e.y, t.y, ph[1]);
            if (m & 4u) disp_cell_r(dc + ((long long)id.c * p.maxpoles + q) * 3, e.z, t.z, ph[2]);
            if (m & 8u) disp_cell_r(dc + ((long long)id.d * p.maxpoles + q) * 3, e.w, t.w, ph[3]);
            st4(T, t);
        }
    } else {
        R *T = reinterpret_cast<R *>(p.T[comp] + off);
        for (int q = 0; q < p.maxpoles; ++q, T += 2 * p.tstride) {
            V4<R> t01 = ld4(T), t23 = ld4(T + 4);
            if (m & 1u) disp_cell_c(p.dcoef + ((long long)id.a * p.maxpoles + q) * 3, e.x, t01.x, t01.y, ph[0]);
            if (m & 2u) disp_cell_c(p.dcoef + ((long long)id.b * p.maxpoles + q) * 3, e.y, t01.z, t01.w, ph[1]);
            if (m & 4u) disp_cell_c(p.dcoef + ((long long)id.c * p.maxpoles + q) * 3, e.z, t23.x, t23.y, ph[2]);
            if (m & 8u) disp_cell_c(p.dcoef + ((long long)id.d * p.maxpoles + q) * 3, e.w, t23.z, t23.w, ph[3]);
            st4(T, t01);
            st4(T + 4, t23);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Electric half-step, 4 z cells per thread.  Arithmetic: fields_updates_ext.pyx:30-107 and
// pml_updates_electric_*_ext.pyx (x and y slabs).
// ------------------------------------------------------------------------------------------
template <typename R, typename IDT, bool DISP, bool NC>
__device__ __forceinline__ void e4_body(const PhaseParams<R> &p, const Coef4<R> *coef, const R *srce, int blk, int chunk, int tid)
{
    const long long idx4 = (long long)blk * kThreadsV4 + tid;
    const long long e0 = idx4 * 4;
    const bool valid = e0 < p.plane;
    const long long eoff = valid ? e0 : 4;  // 4: keeps the k-1 load of lane 0 inside the allocation
    const int j = (int)(eoff / p.pitch);
    const int k = (int)(eoff - (long long)j * p.pitch);
    const int lane = tid & 31;
    const int l0 = p.p0 + chunk * p.xchunk;
    const int l1 = min(l0 + p.xchunk, p.p1);
    if (l0 >= l1) return;
    const JK4 bx = jk4_of(p.box[0].lo, p.box[0].hi, j, k), by = jk4_of(p.box[1].lo, p.box[1].hi, j, k), bz = jk4_of(p.box[2].lo, p.box[2].hi, j, k);
    unsigned smask = 0;  // 4 bits per slab
#pragma unroll
    for (int s = 0; s < kMaxSlabs; ++s)
        if (s < p.nslabs && p.slab[s].axis != 2 && valid) smask |= jk4_of(p.slab[s].lo, p.slab[s].hi, j, k).kmask << (4 * s);
    const bool any = valid && ((bx.kmask | by.kmask | bz.kmask) != 0u || smask != 0u);

    typename MaybeRestrict<NC, const R>::type Hx = p.Hx, Hy = p.Hy, Hz = p.Hz;
    typename MaybeRestrict<NC, R>::type Ex = p.Ex, Ey = p.Ey, Ez = p.Ez;
    long long off = (long long)(l0 + 1) * p.plane + eoff;
    V4<R> hy_p = ld4(Hy + off - p.plane), hz_p = ld4(Hz + off - p.plane);
#pragma unroll kV4Unroll
    for (int l = l0; l < l1; ++l, off += p.plane) {
        const int i = p.x_start + l;
        const V4<R> hx_c = ld4(Hx + off), hy_c = ld4(Hy + off), hz_c = ld4(Hz + off);
        const V4<R> hx_j = ld4(Hx + off - p.pitch), hz_j = ld4(Hz + off - p.pitch);
        // k-1 operands: previous lane's .w (scalar load on the warp-edge lane)
        R hx_km = __shfl_up_sync(0xffffffffu, hx_c.w, 1), hy_km = __shfl_up_sync(0xffffffffu, hy_c.w, 1);
        if (lane == 0) {
            hx_km = Hx[off - 1];
            hy_km = Hy[off - 1];
        }
        if (any) {
            const V4<R> dHz_dy = {hz_c.x - hz_j.x, hz_c.y - hz_j.y, hz_c.z - hz_j.z, hz_c.w - hz_j.w};
            const V4<R> dHy_dz = {hy_c.x - hy_km, hy_c.y - hy_c.x, hy_c.z - hy_c.y, hy_c.w - hy_c.z};
            const V4<R> dHx_dz = {hx_c.x - hx_km, hx_c.y - hx_c.x, hx_c.z - hx_c.y, hx_c.w - hx_c.z};
            const V4<R> dHz_dx = {hz_c.x - hz_p.x, hz_c.y - hz_p.y, hz_c.z - hz_p.z, hz_c.w - hz_p.w};
            const V4<R> dHy_dx = {hy_c.x - hy_p.x, hy_c.y - hy_p.y, hy_c.z - hy_p.z, hy_c.w - hy_p.w};
            const V4<R> dHx_dy = {hx_c.x - hx_j.x, hx_c.y - hx_j.y, hx_c.z - hx_j.z, hx_c.w - hx_j.w};
            const unsigned mx = (i >= p.box[0].lo[0] && i < p.box[0].hi[0]) ? bx.kmask : 0u;
            const unsigned my = (i >= p.box[1].lo[0] && i < p.box[1].hi[0]) ? by.kmask : 0u;
            const unsigned mz = (i >= p.box[2].lo[0] && i < p.box[2].hi[0]) ? bz.kmask : 0u;
            unsigned pm = 0;
#pragma unroll
            for (int s = 0; s < kMaxSlabs; ++s)
                if (((smask >> (4 * s)) & 0xfu) && i >= p.slab[s].lo[0] && i < p.slab[s].hi[0]) pm |= 1u << s;
            bool wx = mx != 0, wy = my != 0, wz = mz != 0;
            V4<R> ex, ey, ez;
            Ids4 idx_, idy_, idz_;
            const bool needx = wx || (pm != 0), needy = wy || (pm != 0), needz = wz || (pm != 0);
            if (needx) {
                ex = ld4(Ex + off);
                idx_ = ld_ids4<IDT>(p.ID[0], off);
            }
            if (needy) {
                ey = ld4(Ey + off);
                idy_ = ld_ids4<IDT>(p.ID[1], off);
            }
            if (needz) {
                ez = ld4(Ez + off);
                idz_ = ld_ids4<IDT>(p.ID[2], off);
            }
            if (mx) {
                Coef4<R> c0, c1, c2, c3;
                coef4(coef, idx_, c0, c1, c2, c3);
                if (DISP) {
                    float ph[4];
                    disp4(p, 0, idx_, off, mx, ex, ph);
                    ex.x = sel(mx, 0, disp_sub(upd3(c0.a, ex.x, c0.by, dHz_dy.x, -c0.bz, dHy_dz.x), srce[idx_.a], ph[0]), ex.x);
                    ex.y = sel(mx, 1, disp_sub(upd3(c1.a, ex.y, c1.by, dHz_dy.y, -c1.bz, dHy_dz.y), srce[idx_.b], ph[1]), ex.y);
                    ex.z = sel(mx, 2, disp_sub(upd3(c2.a, ex.z, c2.by, dHz_dy.z, -c2.bz, dHy_dz.z), srce[idx_.c], ph[2]), ex.z);
                    ex.w = sel(mx, 3, disp_sub(upd3(c3.a, ex.w, c3.by, dHz_dy.w, -c3.bz, dHy_dz.w), srce[idx_.d], ph[3]), ex.w);
                } else {
                ex.x = sel(mx, 0, upd3(c0.a, ex.x, c0.by, dHz_dy.x, -c0.bz, dHy_dz.x), ex.x);
                ex.y = sel(mx, 1, upd3(c1.a, ex.y, c1.by, dHz_dy.y, -c1.bz, dHy_dz.y), ex.y);
                ex.z = sel(mx, 2, upd3(c2.a, ex.z, c2.by, dHz_dy.z, -c2.bz, dHy_dz.z), ex.z);
                ex.w = sel(mx, 3, upd3(c3.a, ex.w, c3.by, dHz_dy.w, -c3.bz, dHy_dz.w), ex.w);
                }
            }
            if (my) {
                Coef4<R> c0, c1, c2, c3;
                coef4(coef, idy_, c0, c1, c2, c3);
                if (DISP) {
                    float ph[4];
                    disp4(p, 1, idy_, off, my, ey, ph);
                    ey.x = sel(my, 0, disp_sub(upd3(c0.a, ey.x, c0.bz, dHx_dz.x, -c0.bx, dHz_dx.x), srce[idy_.a], ph[0]), ey.x);
                    ey.y = sel(my, 1, disp_sub(upd3(c1.a, ey.y, c1.bz, dHx_dz.y, -c1.bx, dHz_dx.y), srce[idy_.b], ph[1]), ey.y);
                    ey.z = sel(my, 2, disp_sub(upd3(c2.a, ey.z, c2.bz, dHx_dz.z, -c2.bx, dHz_dx.z), srce[idy_.c], ph[2]), ey.z);
                    ey.w = sel(my, 3, disp_sub(upd3(c3.a, ey.w, c3.bz, dHx_dz.w, -c3.bx, dHz_dx.w), srce[idy_.d], ph[3]), ey.w);
                } else {
                ey.x = sel(my, 0, upd3(c0.a, ey.x, c0.bz, dHx_dz.x, -c0.bx, dHz_dx.x), ey.x);
                ey.y = sel(my, 1, upd3(c1.a, ey.y, c1.bz, dHx_dz.y, -c1.bx, dHz_dx.y), ey.y);
                ey.z = sel(my, 2, upd3(c2.a, ey.z, c2.bz, dHx_dz.z, -c2.bx, dHz_dx.z), ey.z);
                ey.w = sel(my, 3, upd3(c3.a, ey.w, c3.bz, dHx_dz.w, -c3.bx, dHz_dx.w), ey.w);
                }
            }
            if (mz) {
                Coef4<R> c0, c1, c2, c3;
                coef4(coef, idz_, c0, c1, c2, c3);
                if (DISP) {
                    float ph[4];
                    disp4(p, 2, idz_, off, mz, ez, ph);
                    ez.x = sel(mz, 0, disp_sub(upd3(c0.a, ez.x, c0.bx, dHy_dx.x, -c0.by, dHx_dy.x), srce[idz_.a], ph[0]), ez.x);
                    ez.y = sel(mz, 1, disp_sub(upd3(c1.a, ez.y, c1.bx, dHy_dx.y, -c1.by, dHx_dy.y), srce[idz_.b], ph[1]), ez.y);
                    ez.z = sel(mz, 2, disp_sub(upd3(c2.a, ez.z, c2.bx, dHy_dx.z, -c2.by, dHx_dy.z), srce[idz_.c], ph[2]), ez.z);
                    ez.w = sel(mz, 3, disp_sub(upd3(c3.a, ez.w, c3.bx, dHy_dx.w, -c3.by, dHx_dy.w), srce[idz_.d], ph[3]), ez.w);
                } else {
                ez.x = sel(mz, 0, upd3(c0.a, ez.x, c0.bx, dHy_dx.x, -c0.by, dHx_dy.x), ez.x);
                ez.y = sel(mz, 1, upd3(c1.a, ez.y, c1.bx, dHy_dx.y, -c1.by, dHx_dy.y), ez.y);
                ez.z = sel(mz, 2, upd3(c2.a, ez.z, c2.bx, dHy_dx.z, -c2.by, dHx_dy.z), ez.z);
                ez.w = sel(mz, 3, upd3(c3.a, ez.w, c3.bx, dHy_dx.w, -c3.by, dHx_dy.w), ez.w);
                }
            }
            if (pm) {
                for (int s = 0; s < p.nslabs; ++s) {
                    if (!((pm >> s) & 1u)) continue;
                    const SlabDev<R> &sl = p.slab[s];
                    const int pos = sl.axis == 0 ? i : j;
                    const int depth = sl.minus ? (sl.dref - pos) : (pos - sl.dref);
                    const PmlCo<R> co = pml_load(p.form, p.order, sl, depth);
                    R *phi = sl.phi + ((long long)(i - sl.lo[0]) * sl.n1 + (j - sl.lo[1])) * sl.n2 + (k - sl.ko);
                    const unsigned m = (smask >> (4 * s)) & 0xfu;
                    if (sl.axis == 0) {  // Ey -= , dHz/dx ; Ez += , dHy/dx
                        pml_comp4(p.form, p.order, co, sl, phi, m, idy_, srce, (R)-1, dHz_dx, ey);
                        pml_comp4(p.form, p.order, co, sl, phi + sl.ostride, m, idz_, srce, (R)1, dHy_dx, ez);
                        wy = wz = true;
                    } else {  // Ex += , dHz/dy ; Ez -= , dHx/dy
                        pml_comp4(p.form, p.order, co, sl, phi, m, idx_, srce, (R)1, dHz_dy, ex);
                        pml_comp4(p.form, p.order, co, sl, phi + sl.ostride, m, idz_, srce, (R)-1, dHx_dy, ez);
                        wx = wz = true;
                    }
                }
            }
            if (wx) st4(Ex + off, ex);
            if (wy) st4(Ey + off, ey);
            if (wz) st4(Ez + off, ez);
        }
        hy_p = hy_c;
        hz_p = hz_c;
    }
}

template <typename R, typename IDT, bool TABSMEM, bool DISP>
__global__ void __launch_bounds__(kThreadsV4, GPB_V4_MINBLOCKS) k_update_e4(const PhaseParams<R> p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pdl_launch_dependents();
    const Coef4<R> *coef = p.coef;
    const R *srce = p.src;
    if (TABSMEM) {
        Coef4<R> *scoef = reinterpret_cast<Coef4<R> *>(smem_raw);
        R *ssrc = reinterpret_cast<R *>(scoef + p.nmat);
        stage_coefs(p, scoef, ssrc);
        coef = scoef;
        srce = ssrc;
    }
    pdl_wait();   // the coefficient tables staged above are constant during a run
    e4_body<R, IDT, DISP, true>(p, coef, srce, (int)blockIdx.x, (int)blockIdx.y, (int)threadIdx.x);
}

// ------------------------------------------------------------------------------------------
// Slab-wise PML correction, one thread per slab cell (both components).  Used for the z slabs of
// the vectorised path: the 10 cells at either end of every z row would diverge in every warp of the
// main kernel.  `phase` 0 = magnetic, 1 = electric.  Grid: x = cells of one plane of the slab box,
// y = planes.
// ------------------------------------------------------------------------------------------
// one slab cell, both components; `phase` 0 = magnetic, 1 = electric
template <typename R, typename IDT>
__device__ __forceinline__ void pml_slab_cell(const PhaseParams<R> &p, int phase, const SlabDev<R> &sl, int i, int j, int k)
{
    const int a = sl.axis;
    const int pos = a == 0 ? i : (a == 1 ? j : k);
    const int depth = sl.minus ? (sl.dref - pos) : (pos - sl.dref);
    const long long off = (long long)(i - p.x_start + 1) * p.plane + (long long)j * p.pitch + k;
    R *phi = sl.phi + ((long long)(i - sl.lo[0]) * sl.n1 + (j - sl.lo[1])) * sl.n2 + (k - sl.ko);
    const long long st = a == 0 ? p.plane : (a == 1 ? p.pitch : 1);
    const R *src = p.src;
    R *Fa, *Fb;
    const R *Ga, *Gb;
    R sa, sb, dA, dB;
    int ca, cb;
    if (phase == 1) {
        // electric: backward differences of H along the slab axis
        if (a == 0) { Fa = p.Ey; Ga = p.Hz; sa = -1; ca = 1; Fb = p.Ez; Gb = p.Hy; sb = 1; cb = 2; }
        else if (a == 1) { Fa = p.Ex; Ga = p.Hz; sa = 1; ca = 0; Fb = p.Ez; Gb = p.Hx; sb = -1; cb = 2; }
        else { Fa = p.Ex; Ga = p.Hy; sa = -1; ca = 0; Fb = p.Ey; Gb = p.Hx; sb = 1; cb = 1; }
        dA = (Ga[off] - Ga[off - st]) * sl.inv_d;
        dB = (Gb[off] - Gb[off - st]) * sl.inv_d;
    } else {
        // magnetic: forward differences of E along the slab axis
        if (a == 0) { Fa = p.Hy; Ga = p.Ez; sa = 1; ca = 1; Fb = p.Hz; Gb = p.Ey; sb = -1; cb = 2; }
        else if (a == 1) { Fa = p.Hx; Ga = p.Ez; sa = -1; ca = 0; Fb = p.Hz; Gb = p.Ex; sb = 1; cb = 2; }
        else { Fa = p.Hx; Ga = p.Ey; sa = 1; ca = 0; Fb = p.Hy; Gb = p.Ex; sb = -1; cb = 1; }
        dA = (Ga[off + st] - Ga[off]) * sl.inv_d;
        dB = (Gb[off + st] - Gb[off]) * sl.inv_d;
    }
    const unsigned ma = ld_id<IDT>(p.ID[ca], off), mb = ld_id<IDT>(p.ID[cb], off);
    Fa[off] = fma_(sa, mul_(src[ma], pml_term(p.form, p.order, sl, depth, dA, phi, sl.ostride)), Fa[off]);
    Fb[off] = fma_(sb, mul_(src[mb], pml_term(p.form, p.order, sl, depth, dB, phi + sl.ostride, sl.ostride)), Fb[off]);
}

// ------------------------------------------------------------------------------------------
// Slab-wise PML correction, one thread per slab cell (both components).  Used for the z slabs of
// the vectorised path: the 10 cells at either end of every z row would diverge in every warp of the
// main kernel.  Grid: x = cells of one plane of the slab box, y = planes.
// ------------------------------------------------------------------------------------------
template <typename R, typename IDT>
__global__ void __launch_bounds__(256) k_pml_slabs(const PhaseParams<R> p, int phase, unsigned slabsel, int p0, int p1)
{
    pdl_launch_dependents();
    pdl_wait();
    // blockIdx.z selects the n-th slab of `slabsel`: the slabs run side by side instead of back to back
    int s = -1;
    for (int q = 0, n = 0; q < p.nslabs; ++q)
        if ((slabsel >> q) & 1u) {
            if (n == (int)blockIdx.z) s = q;
            ++n;
        }
    if (s < 0) return;
    const SlabDev<R> &sl = p.slab[s];
    const int n1 = sl.hi[1] - sl.lo[1], n2 = sl.hi[2] - sl.lo[2];
    const int i = sl.lo[0] + blockIdx.y;
    if (i >= sl.hi[0] || i < p.x_start + p0 || i >= p.x_start + p1) return;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n1 * n2) return;
    pml_slab_cell<R, IDT>(p, phase, sl, i, sl.lo[1] + q / n2, sl.lo[2] + q % n2);
}

}  // namespace gpb
