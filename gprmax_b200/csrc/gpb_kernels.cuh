// gpb_kernels.cuh -- device side of the B200 FDTD core (sm_100a).
//
// Replaces the reference's CUDA-in-Python kernels:
//   update_e / update_h / update_e_dispersive_A,B   gprMax/fields_updates_gpu.py:40-240
//   order{1,2}_{x,y,z}{minus,plus} x 4 files         gprMax/pml_updates/*_gpu.py
//   update_hertzian_dipole / magnetic_dipole / voltage_source   gprMax/source_updates_gpu.py:38-200
//   store_outputs                                    gprMax/fields_outputs.py:81-105
//   store_snapshot                                   gprMax/snapshots_gpu.py:31-77
// with the index ranges and arithmetic of the CPU solver the traces are judged against
// (fields_updates_ext.pyx, pml_updates/*_ext.pyx, sources.py, snapshots_ext.pyx).
//
// Design:
//   * private device layout: z pitch padded to a multiple of 32 elements, one ghost plane on each
//     x side (halo for x-slab sharding), material IDs narrowed to u8/u16 when they fit;
//   * one thread per (j,k) column of a plane, marching XCHUNK planes along x with a register queue
//     for the i-1 (E phase) / i+1 (H phase) operands -> every operand array is read once per phase;
//   * the CFS-PML slab corrections are fused into the same pass (no extra sweeps over E/H/ID; the
//     only extra traffic is the Phi state); on edge / corner cells the x and y slabs are applied in
//     G.pmls order and the z slabs after them -- one order for every kernel family, a last-ulp
//     difference from the reference's x0,y0,z0,xmax,ymax,zmax sequence;
//   * dispersive part B of step n is folded into part A of step n+1 (E is untouched in between),
//     removing one full pass over E, ID and T per step;
//   * coefficient rows live in shared memory (no constant-cache serialisation with many materials).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gpb {

// Programmatic dependent launch (GPB_PDL=0 switches it off): a kernel launched with the programmatic-stream-serialization attribute may
// become resident while its predecessor in the stream is still draining -- every CTA of the predecessor has passed
// pdl_launch_dependents() or exited -- and runs its prologue (shared-memory tables, barrier set-up, descriptor prefetch) there;
// pdl_wait() returns once the predecessor has completed and its writes are visible.  EVERY kernel of the time step calls
// pdl_wait() in every CTA before its first access to anything a predecessor writes (fields, Phi, T, iteration counter, work
// queue) and before it exits, so that "kernel N complete" still implies "kernels < N complete".  Both instructions are no-ops
// in a kernel that was launched without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Launch `kern` with or without the programmatic-stream-serialization attribute (host side of the above).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

constexpr int kThreads = 256;
constexpr int kXChunk = 16;
constexpr int kMaxSlabs = 6;

template <typename R>
struct alignas(4 * sizeof(R)) Coef4 {
    R a, bx, by, bz;  // E: CA, CBx, CBy, CBz ; H: DA, DBx, DBy, DBz  (materials.py:200-201)
};

template <typename R>
struct Cplx {
    R re, im;
};

// One PML slab as seen by one phase (electric or magnetic).
template <typename R>
struct SlabDev {
    int lo[3], hi[3];  // global FIELD-index box [lo, hi) touched in this phase
    int axis;          // 0 x, 1 y, 2 z
    int minus;         // 1: depth = dref - idx ; 0: depth = idx - dref
    int dref;
    int t;             // thickness = row length of the R tables
    int n1, n2;        // extents of the box along y and z (Phi is [order][2][n0][n1][n2])
    int ko;            // z index of Phi column 0 (x/y slabs: lo[2]; z slabs: lo[2] rounded down to a multiple of 4, rows padded
                       // to a multiple of 4 so that the 4-cell threads of the vectorised kernels use aligned 128-bit accesses)
    long long ostride; // n0*n1*n2 : distance between Phi components / orders
    R *phi;
    const R *RA, *RB, *RE, *RF;
    R d;               // spacing, rounded through float first like the reference's `float d`
    R inv_d;           // 1 / d (TMA kernels multiply: the IEEE division's slow path, taken for the denormal field tails that
                       // fill most of a PML, made up 17 % of all executed instructions -- profiles/README.md r1h)
};

struct Box {
    int lo[3], hi[3];  // global index box in which a component's base update runs
};

template <typename R>
struct PhaseParams {
    // geometry
    int nx, ny, nz;        // global cells
    int x_start, nplanes;  // owned node planes
    int pitch;             // z pitch (elements)
    long long plane;       // (ny+1)*pitch
    // fields: pointers to local plane 0 (= ghost plane x_start-1)
    R *Ex, *Ey, *Ez, *Hx, *Hy, *Hz;
    const void *ID[3];     // the three ID components of this phase (same layout as fields)
    const Coef4<R> *coef;  // [nmat]
    const R *src;          // [nmat] srce / srcm
    int nmat;
    Box box[3];
    int nslabs;
    int form;              // 0 HORIPML, 1 MRIPML
    int order;             // 1, 2
    SlabDev<R> slab[kMaxSlabs];
    // dispersive (E phase only)
    int maxpoles;
    Cplx<R> *T[3];            // [maxpoles][nplanes+2][ny+1][pitch]   (treal: the same memory holds R instead of complex)
    const Cplx<R> *dcoef;     // [nmat][maxpoles][3]                  (treal: R[nmat][maxpoles][3], the real parts)
    long long tstride;        // elements between poles
    int treal;                // every dispersive coefficient is real (Debye media): T is stored and advanced as a real array --
                              // half the bytes of the part that dominates a dispersive half-step; same E bits as the complex form
    // sub-range of owned planes processed by this launch (for halo overlap): [p0, p1)
    int p0, p1;
    int xchunk;               // planes marched by one CTA of the TMA kernels
    int fast_i0, fast_i1;     // TMA kernels: planes between the x slabs and inside all three update boxes
    int zfused;               // TMA kernels: z-slab PML in the same pass (else k_pml_slabs afterwards)
    int znocoop;              // TMA kernels: z-slab corrections per thread (4 cells) instead of one cell per lane
    int tmax;                 // TMA kernels: row length of the shared-memory copies of the PML R tables (max thickness)
    int pf_depth;             // TMA kernels: Phi prefetch distance in planes (cp.async ring per thread), 0 = direct loads
    int persist;              // TMA kernels: persistent CTAs pulling (tile, x-chunk) items from an atomic counter
    int t_depth;              // TMA kernels, dispersive E half-step: T prefetch distance in planes (cp.async ring per thread), 0 = direct loads
    // k_update_pair (both half-steps of an iteration in one launch): x chunks are handed out in increasing x; finished H items are
    // counted per chunk in progress[], the producer loads an E item only when the H items of its chunk and of the one before it
    // are complete -- so E reads H (and re-reads E) out of L2
    // TMA kernels, linked x-slab shards: results of array plane `peer_plane` (components 1 and 2 of the phase: Ey,Ez / Hy,Hz)
    // are stored to the neighbour's ghost plane as well; peer1 / peer2 point at that plane in the neighbour's arrays
    R *peer1, *peer2;
    int peer_plane;
    // ... and the warp that completes the last item holding that plane publishes *peer_flag = *peer_iter + peer_add in the
    // neighbour's memory (null: the host launches k_flag_signal after the half-step instead)
    unsigned *peer_flag, *peer_counter;
    unsigned *peer_flag2;     // second flag published with the same value (E half-step: the neighbour's H_FREE credit), or null
    unsigned peer_need;       // tiles * consumer warps
    const int *peer_iter;
    int peer_add;
    int monotone;
    int pair_lag;             // k_update_pair: the E items of chunk c follow the H items of chunk c + pair_lag in the queue
    unsigned *progress;       // [nchunks] finished H warps per chunk (null: kernels run one after the other)
    unsigned prog_need;       // tiles * consumer warps
    unsigned *prog_flags;     // [0] |= 1 when a wait timed out
    unsigned long long prog_timeout_ns;
};

// ------------------------------------------------------------------------------------------
// Explicitly rounded building blocks.  Every kernel family (scalar, register-vectorised, TMA-staged) forms the field
// update and the PML terms from these, so the placement of the fused multiply-adds is fixed by the source and not by
// how the compiler happens to contract a*b + c*d in each kernel: a sharded run may put a shard on another kernel
// family than the single-GPU run it must reproduce bit for bit (an inlined a*f + b*d - c*e came out with two
// different contractions in k_update_tma and k_update_e4: 1-ulp differences next to the source after 3 iterations).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_(double a, double b, double c) { return __fma_rn(a, b, c); }
// a*f + b1*d1 + b2*d2 (pass -b for a subtracted term): fma(b2, d2, fma(b1, d1, a*f))
template <typename R>
__device__ __forceinline__ R upd3(R a, R f, R b1, R d1, R b2, R d2)
{
    return fma_(b2, d2, fma_(b1, d1, mul_(a, f)));
}

// ------------------------------------------------------------------------------------------
// PML recursive integration.  Arithmetic follows pml_updates_electric_HORIPML_ext.pyx:69-87 (order 1), :132-157
// (order 2) and pml_updates_electric_MRIPML_ext.pyx:68-89, :133-162 (identical in the magnetic files).
// PmlCo = the coefficient set of one depth; pml_apply returns the bracket that multiplies srce/srcm and advances Phi.
// ------------------------------------------------------------------------------------------
template <typename R>
struct PmlCo {
    R a, b, c, d, e, f, g, h, r;
};
// from the R-table values of one depth (index 0 / 1 = first / second CFS)
template <typename R>
__device__ __forceinline__ PmlCo<R> pml_co(int form, int order, R RA0, R RA1, R RB0, R RB1, R RE0, R RE1, R RF0, R RF1)
{
    PmlCo<R> c;
    const R one = (R)1;
    if (form == 0) {
        if (order == 1) {
            c.a = RA0 - one;  // RA01
            c.b = RB0;
            c.e = RE0;
            c.f = RF0;
        } else {
            c.a = fma_(RA0, RA1, -one);  // RA01
            c.b = RB0;
            c.e = RE0;
            c.f = RF0;
            c.c = RA0;
            c.d = RA1;
            c.g = RB1;
            c.h = RE1;
            c.r = RF1;
        }
    } else {
        if (order == 1) {
            const R IRA = one / RA0;
            c.a = IRA;
            c.b = IRA - one;
            c.c = mul_(mul_(IRA, RB0), RF0);  // RC0
            c.e = RE0;
        } else {
            const R IRA = one / (RA0 + RA1);
            c.a = IRA;
            c.b = IRA - one;
            c.c = mul_(IRA, RF0);
            c.d = mul_(IRA, RF1);
            c.e = RE0;
            c.h = RE1;
            c.f = RB0;
            c.g = RB1;
        }
    }
    return c;
}
template <typename R>
__device__ __forceinline__ PmlCo<R> pml_load(int form, int order, const SlabDev<R> &s, int depth)
{
    const R RA0 = __ldg(s.RA + depth), RB0 = __ldg(s.RB + depth), RE0 = __ldg(s.RE + depth), RF0 = __ldg(s.RF + depth);
    R RA1 = RA0, RB1 = RB0, RE1 = RE0, RF1 = RF0;
    if (order == 2) {
        RA1 = __ldg(s.RA + s.t + depth); RB1 = __ldg(s.RB + s.t + depth); RE1 = __ldg(s.RE + s.t + depth); RF1 = __ldg(s.RF + s.t + depth);
    }
    return pml_co(form, order, RA0, RA1, RB0, RB1, RE0, RE1, RF0, RF1);
}
// phi0 / phi1 are updated in registers
template <typename R>
__device__ __forceinline__ R pml_apply(int form, int order, const PmlCo<R> &c, R dF, R &phi0, R &phi1)
{
    R term;
    if (form == 0) {
        if (order == 1) {
            term = fma_(c.a, dF, mul_(c.b, phi0));                       // RA01 dF + RB0 Phi0
            phi0 = fma_(c.e, phi0, -mul_(c.f, dF));                      // RE0 Phi0 - RF0 dF
        } else {
            term = fma_(c.g, phi1, fma_(c.a, dF, mul_(mul_(c.d, c.b), phi0)));   // RA01 dF + RA1 RB0 Phi0 + RB1 Phi1
            phi1 = fma_(c.h, phi1, -mul_(c.r, fma_(c.c, dF, mul_(c.b, phi0))));  // RE1 Phi1 - RF1 (RA0 dF + RB0 Phi0)
            phi0 = fma_(c.e, phi0, -mul_(c.f, dF));
        }
    } else {
        if (order == 1) {
            term = fma_(c.b, dF, -mul_(c.a, phi0));                      // (IRA - 1) dF - IRA Phi0
            phi0 = fma_(-c.c, phi0, fma_(c.c, dF, mul_(c.e, phi0)));     // RE0 Phi0 + RC0 dF - RC0 Phi0
        } else {
            const R psi = fma_(c.f, phi0, mul_(c.g, phi1));              // RB0 Phi0 + RB1 Phi1
            term = fma_(c.b, dF, -mul_(c.a, psi));
            const R dd = dF - psi;
            phi1 = fma_(c.h, phi1, mul_(c.d, dd));
            phi0 = fma_(c.e, phi0, mul_(c.c, dd));
        }
    }
    return term;
}
// one cell, Phi in memory (scalar kernels, k_pml_slabs)
template <typename R>
__device__ __forceinline__ R pml_term(int form, int order, const SlabDev<R> &s, int depth, R dF, R *phi0p, long long ostride)
{
    const PmlCo<R> co = pml_load(form, order, s, depth);
    R *phi1p = phi0p + 2 * ostride;
    R phi0 = *phi0p, phi1 = order == 2 ? *phi1p : phi0;
    const R term = pml_apply(form, order, co, dF, phi0, phi1);
    *phi0p = phi0;
    if (order == 2) *phi1p = phi1;
    return term;
}

template <typename R>
__device__ __forceinline__ bool in_box(const int *lo, const int *hi, int i, int j, int k)
{
    return i >= lo[0] && i < hi[0] && j >= lo[1] && j < hi[1] && k >= lo[2] && k < hi[2];
}

template <typename IDT>
__device__ __forceinline__ unsigned ld_id(const void *base, long long off)
{
    return (unsigned)__ldg(reinterpret_cast<const IDT *>(base) + off);
}

// Shared-memory staging of the coefficient rows of one phase.
template <typename R>
__device__ __forceinline__ void stage_coefs(const PhaseParams<R> &p, Coef4<R> *scoef, R *ssrc)
{
    for (int m = threadIdx.x; m < p.nmat; m += blockDim.x) {
        scoef[m] = p.coef[m];
        ssrc[m] = p.src[m];
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// Magnetic half-step: base update (fields_updates_ext.pyx:352-412) + fused H-PML corrections
// (pml_updates_magnetic_*_ext.pyx).  Marches +x; queue holds Ey,Ez of the next plane.
// ------------------------------------------------------------------------------------------
template <typename R, typename IDT, bool TABSMEM>
__global__ void __launch_bounds__(kThreads) k_update_h(const PhaseParams<R> p)
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
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int j = (int)(idx / p.pitch);
    const int k = (int)(idx - (long long)j * p.pitch);
    if (j > p.ny || k > p.nz) return;
    const bool hasj = j < p.ny, hask = k < p.nz;  // rows j+1 / k+1 exist
    const int l0 = p.p0 + blockIdx.y * kXChunk;
    const int l1 = min(l0 + kXChunk, p.p1);
    if (l0 >= l1) return;
    long long off = (long long)(l0 + 1) * p.plane + idx;  // +1: ghost plane in front
    R ey_c = p.Ey[off], ez_c = p.Ez[off];
#pragma unroll 2
    for (int l = l0; l < l1; ++l, off += p.plane) {
        const int i = p.x_start + l;
        const R ey_n = p.Ey[off + p.plane], ez_n = p.Ez[off + p.plane];
        const R ex_c = p.Ex[off];
        const R ex_j = hasj ? p.Ex[off + p.pitch] : (R)0;
        const R ex_k = hask ? p.Ex[off + 1] : (R)0;
        const R ey_k = hask ? p.Ey[off + 1] : (R)0;
        const R ez_j = hasj ? p.Ez[off + p.pitch] : (R)0;
        // the six one-sided differences every update / correction below is built from
        const R dEz_dy = ez_j - ez_c, dEy_dz = ey_k - ey_c;
        const R dEx_dz = ex_k - ex_c, dEz_dx = ez_n - ez_c;
        const R dEy_dx = ey_n - ey_c, dEx_dy = ex_j - ex_c;
        R hx = p.Hx[off], hy = p.Hy[off], hz = p.Hz[off];
        bool wx = false, wy = false, wz = false;
        unsigned mx = 0, my = 0, mz = 0;
        if (in_box<R>(p.box[0].lo, p.box[0].hi, i, j, k)) {
            mx = ld_id<IDT>(p.ID[0], off);
            const Coef4<R> c = coef[mx];
            hx = upd3(c.a, hx, -c.by, dEz_dy, c.bz, dEy_dz);
            wx = true;
        }
        if (in_box<R>(p.box[1].lo, p.box[1].hi, i, j, k)) {
            my = ld_id<IDT>(p.ID[1], off);
            const Coef4<R> c = coef[my];
            hy = upd3(c.a, hy, -c.bz, dEx_dz, c.bx, dEz_dx);
            wy = true;
        }
        if (in_box<R>(p.box[2].lo, p.box[2].hi, i, j, k)) {
            mz = ld_id<IDT>(p.ID[2], off);
            const Coef4<R> c = coef[mz];
            hz = upd3(c.a, hz, -c.bx, dEy_dx, c.by, dEx_dy);
            wz = true;
        }
        for (int s2 = 0; s2 < 2 * p.nslabs; ++s2) {   // x / y slabs in G.pmls order, then the z slabs (as in every kernel family)
            const int s = s2 < p.nslabs ? s2 : s2 - p.nslabs;
            const SlabDev<R> &sl = p.slab[s];
            if ((sl.axis == 2) != (s2 >= p.nslabs)) continue;
            if (!in_box<R>(sl.lo, sl.hi, i, j, k)) continue;
            const int a = sl.axis;
            const int pos = (a == 0) ? i : (a == 1) ? j : k;
            const int depth = sl.minus ? (sl.dref - pos) : (pos - sl.dref);
            R *phi = sl.phi + ((long long)(i - sl.lo[0]) * sl.n1 + (j - sl.lo[1])) * sl.n2 + (k - sl.ko);
            // magnetic x: Hy += , dEz/dx ; Hz -= , dEy/dx   (pml_updates_magnetic_HORIPML_ext.pyx:79-87)
            // magnetic y: Hx -= , dEz/dy ; Hz += , dEx/dy   (:347-355)
            // magnetic z: Hx += , dEy/dz ; Hy -= , dEx/dz   (:615-623)
            if (a == 0) {
                if (!wy) my = ld_id<IDT>(p.ID[1], off);
                if (!wz) mz = ld_id<IDT>(p.ID[2], off);
                hy = fma_((R)1, mul_(srcm[my], pml_term(p.form, p.order, sl, depth, dEz_dx / sl.d, phi, sl.ostride)), hy);
                hz = fma_((R)-1, mul_(srcm[mz], pml_term(p.form, p.order, sl, depth, dEy_dx / sl.d, phi + sl.ostride, sl.ostride)), hz);
                wy = wz = true;
            } else if (a == 1) {
                if (!wx) mx = ld_id<IDT>(p.ID[0], off);
                if (!wz) mz = ld_id<IDT>(p.ID[2], off);
                hx = fma_((R)-1, mul_(srcm[mx], pml_term(p.form, p.order, sl, depth, dEz_dy / sl.d, phi, sl.ostride)), hx);
                hz = fma_((R)1, mul_(srcm[mz], pml_term(p.form, p.order, sl, depth, dEx_dy / sl.d, phi + sl.ostride, sl.ostride)), hz);
                wx = wz = true;
            } else {
                if (!wx) mx = ld_id<IDT>(p.ID[0], off);
                if (!wy) my = ld_id<IDT>(p.ID[1], off);
                hx = fma_((R)1, mul_(srcm[mx], pml_term(p.form, p.order, sl, depth, dEy_dz / sl.d, phi, sl.ostride)), hx);
                hy = fma_((R)-1, mul_(srcm[my], pml_term(p.form, p.order, sl, depth, dEx_dz / sl.d, phi + sl.ostride, sl.ostride)), hy);
                wx = wy = true;
            }
        }
        if (wx) p.Hx[off] = hx;
        if (wy) p.Hy[off] = hy;
        if (wz) p.Hz[off] = hz;
        ey_c = ey_n;
        ez_c = ez_n;
    }
}

// ------------------------------------------------------------------------------------------
// Dispersive media.  Part B of the previous step (fields_updates_ext.pyx:183-235: T -= c2 E, after the PML and the sources)
// is folded into part A of this one (:113-179), because E is not touched in between: one pass over E, ID and T per step
// disappears.  Per cell and pole, with c0 = Re(e0 eqt2), c1 = eqt, c2 = zt (materials.py:204-208):
//     t    = T - c2 E_old                 (part B, deferred)
//     phi += c0 Re(t)                     (`phi` is a C float in the reference even in its float64 build, :143)
//     T    = c1 t + c2 E_old              (part A)
// and E_new = (base update) - srce phi.  The products and sums are placed explicitly (mul_ / fma_), identically in every
// kernel family, like the field update itself.  When every coefficient of the model is real (Debye poles), Im(T) stays zero
// and T is kept as a real array (PhaseParams::treal): the real form below gives the same bits as the complex one.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float phi_acc(float phi, float c0, float re) { return fma_(c0, re, phi); }
__device__ __forceinline__ float phi_acc(float phi, double c0, double re) { return (float)fma_(c0, re, (double)phi); }

template <typename R>
__device__ __forceinline__ void disp_cell_c(const Cplx<R> *dc, R e_old, R &tre, R &tim, float &phi)
{
    const Cplx<R> c0 = dc[0], c1 = dc[1], c2 = dc[2];
    const R re = fma_(-c2.re, e_old, tre), im = fma_(-c2.im, e_old, tim);
    phi = phi_acc(phi, c0.re, re);
    tre = fma_(c2.re, e_old, fma_(-c1.im, im, mul_(c1.re, re)));
    tim = fma_(c2.im, e_old, fma_(c1.im, re, mul_(c1.re, im)));
}
template <typename R>
__device__ __forceinline__ void disp_cell_r(const R *dc, R e_old, R &t, float &phi)
{
    const R c0 = dc[0], c1 = dc[1], c2 = dc[2];
    const R re = fma_(-c2, e_old, t);
    phi = phi_acc(phi, c0, re);
    t = fma_(c2, e_old, mul_(c1, re));
}
// E_new = base - srce phi
template <typename R>
__device__ __forceinline__ R disp_sub(R base, R srce, float phi) { return fma_(-srce, (R)phi, base); }

// one cell, all poles, T in global memory (scalar kernels)
template <typename R>
__device__ __forceinline__ float dispersive_AB(const PhaseParams<R> &p, int comp, unsigned m, long long off, R e_old)
{
    float phi = 0;
    if (p.treal) {
        const R *dc = reinterpret_cast<const R *>(p.dcoef) + (long long)m * p.maxpoles * 3;
        R *T = reinterpret_cast<R *>(p.T[comp]) + off;
        for (int q = 0; q < p.maxpoles; ++q, T += p.tstride, dc += 3) {
            R t = *T;
            disp_cell_r(dc, e_old, t, phi);
            *T = t;
        }
    } else {
        const Cplx<R> *dc = p.dcoef + (long long)m * p.maxpoles * 3;
        Cplx<R> *T = p.T[comp] + off;
        for (int q = 0; q < p.maxpoles; ++q, T += p.tstride, dc += 3) {
            Cplx<R> t = *T;
            disp_cell_c(dc, e_old, t.re, t.im, phi);
            *T = t;
        }
    }
    return phi;
}

// ------------------------------------------------------------------------------------------
// Electric half-step: base update (fields_updates_ext.pyx:30-107, dispersive :113-179) + fused
// E-PML corrections (pml_updates_electric_*_ext.pyx).  Marches +x; queue holds Hy,Hz of plane i-1.
// ------------------------------------------------------------------------------------------
template <typename R, typename IDT, bool TABSMEM, bool DISP>
__global__ void __launch_bounds__(kThreads) k_update_e(const PhaseParams<R> p)
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
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int j = (int)(idx / p.pitch);
    const int k = (int)(idx - (long long)j * p.pitch);
    if (j > p.ny || k > p.nz) return;
    const bool hasj = j > 0, hask = k > 0;  // rows j-1 / k-1 exist
    const int l0 = p.p0 + blockIdx.y * kXChunk;
    const int l1 = min(l0 + kXChunk, p.p1);
    if (l0 >= l1) return;
    long long off = (long long)(l0 + 1) * p.plane + idx;
    R hy_p = p.Hy[off - p.plane], hz_p = p.Hz[off - p.plane];
#pragma unroll 2
    for (int l = l0; l < l1; ++l, off += p.plane) {
        const int i = p.x_start + l;
        const R hx_c = p.Hx[off], hy_c = p.Hy[off], hz_c = p.Hz[off];
        const R hx_j = hasj ? p.Hx[off - p.pitch] : (R)0;
        const R hx_k = hask ? p.Hx[off - 1] : (R)0;
        const R hy_k = hask ? p.Hy[off - 1] : (R)0;
        const R hz_j = hasj ? p.Hz[off - p.pitch] : (R)0;
        const R dHz_dy = hz_c - hz_j, dHy_dz = hy_c - hy_k;
        const R dHx_dz = hx_c - hx_k, dHz_dx = hz_c - hz_p;
        const R dHy_dx = hy_c - hy_p, dHx_dy = hx_c - hx_j;
        R ex = p.Ex[off], ey = p.Ey[off], ez = p.Ez[off];
        bool wx = false, wy = false, wz = false;
        unsigned mx = 0, my = 0, mz = 0;
        if (in_box<R>(p.box[0].lo, p.box[0].hi, i, j, k)) {
            mx = ld_id<IDT>(p.ID[0], off);
            const Coef4<R> c = coef[mx];
            if (DISP) {
                const float phi = dispersive_AB(p, 0, mx, off, ex);
                ex = disp_sub(upd3(c.a, ex, c.by, dHz_dy, -c.bz, dHy_dz), srce[mx], phi);
            } else {
                ex = upd3(c.a, ex, c.by, dHz_dy, -c.bz, dHy_dz);
            }
            wx = true;
        }
        if (in_box<R>(p.box[1].lo, p.box[1].hi, i, j, k)) {
            my = ld_id<IDT>(p.ID[1], off);
            const Coef4<R> c = coef[my];
            if (DISP) {
                const float phi = dispersive_AB(p, 1, my, off, ey);
                ey = disp_sub(upd3(c.a, ey, c.bz, dHx_dz, -c.bx, dHz_dx), srce[my], phi);
            } else {
                ey = upd3(c.a, ey, c.bz, dHx_dz, -c.bx, dHz_dx);
            }
            wy = true;
        }
        if (in_box<R>(p.box[2].lo, p.box[2].hi, i, j, k)) {
            mz = ld_id<IDT>(p.ID[2], off);
            const Coef4<R> c = coef[mz];
            if (DISP) {
                const float phi = dispersive_AB(p, 2, mz, off, ez);
                ez = disp_sub(upd3(c.a, ez, c.bx, dHy_dx, -c.by, dHx_dy), srce[mz], phi);
            } else {
                ez = upd3(c.a, ez, c.bx, dHy_dx, -c.by, dHx_dy);
            }
            wz = true;
        }
        for (int s2 = 0; s2 < 2 * p.nslabs; ++s2) {   // x / y slabs in G.pmls order, then the z slabs (as in every kernel family)
            const int s = s2 < p.nslabs ? s2 : s2 - p.nslabs;
            const SlabDev<R> &sl = p.slab[s];
            if ((sl.axis == 2) != (s2 >= p.nslabs)) continue;
            if (!in_box<R>(sl.lo, sl.hi, i, j, k)) continue;
            const int a = sl.axis;
            const int pos = (a == 0) ? i : (a == 1) ? j : k;
            const int depth = sl.minus ? (sl.dref - pos) : (pos - sl.dref);
            R *phi = sl.phi + ((long long)(i - sl.lo[0]) * sl.n1 + (j - sl.lo[1])) * sl.n2 + (k - sl.ko);
            // electric x: Ey -= , dHz/dx ; Ez += , dHy/dx   (pml_updates_electric_HORIPML_ext.pyx:79-87)
            // electric y: Ex += , dHz/dy ; Ez -= , dHx/dy   (:347-355)
            // electric z: Ex -= , dHy/dz ; Ey += , dHx/dz   (:615-623)
            if (a == 0) {
                if (!wy) my = ld_id<IDT>(p.ID[1], off);
                if (!wz) mz = ld_id<IDT>(p.ID[2], off);
                ey = fma_((R)-1, mul_(srce[my], pml_term(p.form, p.order, sl, depth, dHz_dx / sl.d, phi, sl.ostride)), ey);
                ez = fma_((R)1, mul_(srce[mz], pml_term(p.form, p.order, sl, depth, dHy_dx / sl.d, phi + sl.ostride, sl.ostride)), ez);
                wy = wz = true;
            } else if (a == 1) {
                if (!wx) mx = ld_id<IDT>(p.ID[0], off);
                if (!wz) mz = ld_id<IDT>(p.ID[2], off);
                ex = fma_((R)1, mul_(srce[mx], pml_term(p.form, p.order, sl, depth, dHz_dy / sl.d, phi, sl.ostride)), ex);
                ez = fma_((R)-1, mul_(srce[mz], pml_term(p.form, p.order, sl, depth, dHx_dy / sl.d, phi + sl.ostride, sl.ostride)), ez);
                wx = wz = true;
            } else {
                if (!wx) mx = ld_id<IDT>(p.ID[0], off);
                if (!wy) my = ld_id<IDT>(p.ID[1], off);
                ex = fma_((R)-1, mul_(srce[mx], pml_term(p.form, p.order, sl, depth, dHy_dz / sl.d, phi, sl.ostride)), ex);
                ey = fma_((R)1, mul_(srce[my], pml_term(p.form, p.order, sl, depth, dHx_dz / sl.d, phi + sl.ostride, sl.ostride)), ey);
                wx = wy = true;
            }
        }
        if (wx) p.Ex[off] = ex;
        if (wy) p.Ey[off] = ey;
        if (wz) p.Ez[off] = ez;
        hy_p = hy_c;
        hz_p = hz_c;
    }
}

// ------------------------------------------------------------------------------------------
// Point sources, receivers, transmission lines, snapshots.
// ------------------------------------------------------------------------------------------
template <typename R>
struct SrcDev {
    int kind;        // 0 hertzian, 1 magnetic, 2 voltage
    int i, j, k;     // global
    int pol;
    int it_first, it_last;
    int hard;        // voltage source with zero resistance
    R f1, f2;        // hertzian: dl, 1/(dx dy dz) ; magnetic: f2 = 1/(dx dy dz) ; voltage: f2 = 1/(R d1 d2) or (hard) d_pol
    const R *wave;   // [iterations]
};

template <typename R>
struct TLDev {
    int i, j, k, pol;
    int it_first, it_last;
    int nl, srcpos, antpos;
    double coefV;     // resistance * (c dt / dl)
    double coefI;     // (1/resistance) * (c dt / dl)
    double cdtdl;     // c dt / dl
    double h;         // (c dt - dl) / (c dt + dl)
    R dpol;           // spacing along the polarisation (E assignment, sources.py:415-424)
    R *voltage, *current;  // [nl]
    R *abcv;          // [2]
    const R *wave_whole, *wave_half;
    R *Vtotal, *Itotal;    // [iterations]
};

template <typename R>
struct PointParams {
    int x_start, nplanes, ny, nz, pitch;
    long long plane;
    R *F[6];               // Ex,Ey,Ez,Hx,Hy,Hz (local plane 0 = ghost)
    const void *ID[6];
    const R *srcE, *srcH;  // [nmat]
    const int *iter;       // device iteration counter (current)
    int iterations;
    R dx, dy, dz;
    // linked x-slab shards: field arrays of the RIGHT neighbour (peer-mapped, same plane / pitch; local plane 0 = its ghost
    // plane) for the snapshot cells whose +x operands lie on the neighbour's planes; null when there is none
    const R *Fr[6];
    int xr_start, xr_planes;
};

template <typename R>
__device__ __forceinline__ long long pt_off(const PointParams<R> &p, int i, int j, int k)
{
    return (long long)(i - p.x_start + 1) * p.plane + (long long)j * p.pitch + k;
}
template <typename R>
__device__ __forceinline__ bool pt_owned(const PointParams<R> &p, int i)
{
    return i >= p.x_start && i < p.x_start + p.nplanes;
}

// I_x, I_y, I_z of grid.py:413-461 at node (i,j,k); needs H one plane back (ghost is valid).
template <typename R>
__device__ __forceinline__ R current_at(const PointParams<R> &p, int comp, int i, int j, int k)
{
    const long long o = pt_off(p, i, j, k);
    const R *Hx = p.F[3], *Hy = p.F[4], *Hz = p.F[5];
    if (comp == 0) {
        if (j == 0 || k == 0) return (R)0;
        return p.dy * (Hy[o - 1] - Hy[o]) + p.dz * (Hz[o] - Hz[o - p.pitch]);
    } else if (comp == 1) {
        if (i == 0 || k == 0) return (R)0;
        return p.dx * (Hx[o] - Hx[o - 1]) + p.dz * (Hz[o - p.plane] - Hz[o]);
    } else {
        if (i == 0 || j == 0) return (R)0;
        return p.dx * (Hx[o - p.pitch] - Hx[o]) + p.dy * (Hy[o] - Hy[o - p.plane]);
    }
}

// Step prologue: receiver gather (fields_outputs.py:40-64 / :81-105), transmission-line totals
// (:62-64) and the device iteration counter.  One block; `next` is bumped after every thread read it.
template <typename R>
__device__ __forceinline__ void step_begin_body(const PointParams<R> &p, int *iter_cur, int *iter_next, int nrx, const int *rxc, R *rxs,
                                                int ntl, const TLDev<R> *tls)
{
    const int it = *iter_next;
    if (it < p.iterations) {
        for (int r = threadIdx.x; r < nrx; r += blockDim.x) {
            const int i = rxc[3 * r], j = rxc[3 * r + 1], k = rxc[3 * r + 2];
            if (!pt_owned(p, i)) continue;
            const long long o = pt_off(p, i, j, k);
            for (int c = 0; c < 6; ++c) rxs[((long long)c * p.iterations + it) * nrx + r] = p.F[c][o];
            for (int c = 0; c < 3; ++c) rxs[((long long)(6 + c) * p.iterations + it) * nrx + r] = current_at(p, c, i, j, k);
        }
        for (int t = threadIdx.x; t < ntl; t += blockDim.x) {
            if (!pt_owned(p, tls[t].i)) continue;   // sharded runs sum the slabs' outputs: only the owner records
            tls[t].Vtotal[it] = tls[t].voltage[tls[t].antpos];
            tls[t].Itotal[it] = tls[t].current[tls[t].antpos];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *iter_cur = it;
        *iter_next = it + 1;
    }
}

template <typename R>
__global__ void k_step_begin(PointParams<R> p, int *iter_cur, int *iter_next, int nrx, const int *rxc, R *rxs,
                             int ntl, const TLDev<R> *tls)
{
    pdl_launch_dependents();
    pdl_wait();
    step_begin_body(p, iter_cur, iter_next, nrx, rxc, rxs, ntl, tls);
}

// Sources of one phase, applied in array order by a single thread each (order between different
// sources only matters when two sit on the same edge; then the serial loop keeps the CPU order).
// phase 0 (after H update + H-PML): transmission lines (current), magnetic dipoles   model_build_run.py:440-442
// phase 1 (after E update + E-PML): voltage sources, transmission lines (voltage), Hertzian dipoles  :458-461
template <typename R, typename IDT>
__device__ __forceinline__ void sources_body(const PointParams<R> &p, int phase, int nsrc, const SrcDev<R> *srcs, int ntl, const TLDev<R> *tls,
                                             int i_lo, int i_hi, int tl_lo, int tl_hi)
{
    // (one thread) [i_lo, i_hi): global planes whose point sources this launch applies (a sharded half-step applies the sources
    // of its boundary plane before that plane is sent to the neighbour); [tl_lo, tl_hi): planes whose transmission lines it
    // advances (a line belongs to exactly one plane, so every line is advanced by exactly one launch per phase)
    const int it = *p.iter;
    if (phase == 0) {
        for (int t = 0; t < ntl; ++t) {
            const TLDev<R> &tl = tls[t];
            if (it < tl.it_first || it > tl.it_last || !pt_owned(p, tl.i) || tl.i < tl_lo || tl.i >= tl_hi) continue;
            // update_current, sources.py:379-393 (float64 arithmetic, stored as R)
            for (int n = 0; n < tl.nl - 1; ++n) {
                const R dv = tl.voltage[n + 1] - tl.voltage[n];
                tl.current[n] = (R)((double)tl.current[n] - tl.coefI * (double)dv);
            }
            tl.current[tl.srcpos - 1] = (R)((double)tl.current[tl.srcpos - 1] + tl.coefI * (double)tl.wave_half[it]);
            tl.current[tl.antpos] = current_at(p, tl.pol, tl.i, tl.j, tl.k);  // :444-452
        }
        for (int s = 0; s < nsrc; ++s) {
            const SrcDev<R> &sc = srcs[s];
            if (sc.kind != 1 || it < sc.it_first || it > sc.it_last || !pt_owned(p, sc.i) || sc.i < i_lo || sc.i >= i_hi) continue;
            const long long o = pt_off(p, sc.i, sc.j, sc.k);
            const unsigned m = ld_id<IDT>(p.ID[3 + sc.pol], o);
            // sources.py:220-232
            p.F[3 + sc.pol][o] -= p.srcH[m] * sc.wave[it] * sc.f2;
        }
    } else {
        for (int s = 0; s < nsrc; ++s) {
            const SrcDev<R> &sc = srcs[s];
            if (sc.kind != 2 || it < sc.it_first || it > sc.it_last || !pt_owned(p, sc.i) || sc.i < i_lo || sc.i >= i_hi) continue;
            const long long o = pt_off(p, sc.i, sc.j, sc.k);
            if (sc.hard) {
                p.F[sc.pol][o] = -sc.wave[it] / sc.f2;  // sources.py:103, 110, 117
            } else {
                const unsigned m = ld_id<IDT>(p.ID[sc.pol], o);
                p.F[sc.pol][o] -= p.srcE[m] * sc.wave[it] * sc.f2;  // sources.py:99-101
            }
        }
        for (int t = 0; t < ntl; ++t) {
            const TLDev<R> &tl = tls[t];
            if (it < tl.it_first || it > tl.it_last || !pt_owned(p, tl.i) || tl.i < tl_lo || tl.i >= tl_hi) continue;
            // update_voltage, sources.py:360-377
            for (int n = tl.nl - 1; n >= 1; --n) {
                const R dc = tl.current[n] - tl.current[n - 1];
                tl.voltage[n] = (R)((double)tl.voltage[n] - tl.coefV * (double)dc);
            }
            tl.voltage[tl.srcpos] = (R)((double)tl.voltage[tl.srcpos] + tl.cdtdl * (double)tl.wave_whole[it]);
            // update_abc, :348-358
            const R v0 = (R)(tl.h * (double)(R)(tl.voltage[1] - tl.abcv[0]) + (double)tl.abcv[1]);
            tl.voltage[0] = v0;
            tl.abcv[0] = v0;
            tl.abcv[1] = tl.voltage[1];
            p.F[tl.pol][pt_off(p, tl.i, tl.j, tl.k)] = -tl.voltage[tl.antpos] / tl.dpol;  // :415-424
        }
        for (int s = 0; s < nsrc; ++s) {
            const SrcDev<R> &sc = srcs[s];
            if (sc.kind != 0 || it < sc.it_first || it > sc.it_last || !pt_owned(p, sc.i) || sc.i < i_lo || sc.i >= i_hi) continue;
            const long long o = pt_off(p, sc.i, sc.j, sc.k);
            const unsigned m = ld_id<IDT>(p.ID[sc.pol], o);
            // sources.py:181-193
            p.F[sc.pol][o] -= p.srcE[m] * sc.wave[it] * sc.f1 * sc.f2;
        }
    }
}

template <typename R, typename IDT>
__global__ void k_sources(PointParams<R> p, int phase, int nsrc, const SrcDev<R> *srcs, int ntl, const TLDev<R> *tls, int i_lo, int i_hi, int tl_lo, int tl_hi)
{
    pdl_launch_dependents();
    pdl_wait();
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    sources_body<R, IDT>(p, phase, nsrc, srcs, ntl, tls, i_lo, i_hi, tl_lo, tl_hi);
}

// The electric-phase sources of iteration n and the step prologue of iteration n+1 in one launch (one block): between two
// iterations inside a multi-iteration graph nothing else happens, and a kernel boundary costs 2-3 us.
template <typename R, typename IDT>
__global__ void k_sources_begin(PointParams<R> p, int nsrc, const SrcDev<R> *srcs, int ntl_src, const TLDev<R> *tls, int i_lo, int i_hi, int tl_lo, int tl_hi,
                                int *iter_cur, int *iter_next, int nrx, const int *rxc, R *rxs, int ntl)
{
    pdl_launch_dependents();
    pdl_wait();
    if (threadIdx.x == 0) sources_body<R, IDT>(p, 1, nsrc, srcs, ntl_src, tls, i_lo, i_hi, tl_lo, tl_hi);
    __syncthreads();   // the receivers below may sample the edges the sources just wrote
    step_begin_body(p, iter_cur, iter_next, nrx, rxc, rxs, ntl, tls);
}

// Snapshot: strided cell-centred averages (snapshots.py:87-130, snapshots_ext.pyx:56-80).
template <typename R>
struct SnapDev {
    int xs, ys, zs, dx, dy, dz, nx, ny, nz;
    R *out[6];  // each [nx][ny][nz]
};

// value of component c at (global plane gi, in-plane offset jk): own planes, or the right neighbour's (linked shards)
template <typename R>
__device__ __forceinline__ R snap_at(const PointParams<R> &p, int c, int gi, long long jk)
{
    if (gi < p.x_start + p.nplanes) return p.F[c][(long long)(gi - p.x_start + 1) * p.plane + jk];
    return p.Fr[c][(long long)(gi - p.xr_start + 1) * p.plane + jk];
}

template <typename R>
__global__ void k_snapshot(PointParams<R> p, SnapDev<R> s)
{
    const long long n = (long long)s.nx * s.ny * s.nz;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(q % s.nz), j = (int)((q / s.nz) % s.ny), i = (int)(q / ((long long)s.nz * s.ny));
        const int gi = s.xs + i * s.dx, gj = s.ys + j * s.dy, gk = s.zs + k * s.dz;
        if (!pt_owned(p, gi)) continue;   // a snapshot cell belongs to the shard that owns its plane gi
        const long long o = (long long)gj * p.pitch + gk;
        const long long sj = (long long)s.dy * p.pitch, sk = s.dz;
        const int gn = gi + s.dx;
        s.out[0][q] = (snap_at(p, 0, gi, o) + snap_at(p, 0, gi, o + sj) + snap_at(p, 0, gi, o + sk) + snap_at(p, 0, gi, o + sj + sk)) / 4;
        s.out[1][q] = (snap_at(p, 1, gi, o) + snap_at(p, 1, gn, o) + snap_at(p, 1, gi, o + sk) + snap_at(p, 1, gn, o + sk)) / 4;
        s.out[2][q] = (snap_at(p, 2, gi, o) + snap_at(p, 2, gn, o) + snap_at(p, 2, gi, o + sj) + snap_at(p, 2, gn, o + sj)) / 4;
        s.out[3][q] = (snap_at(p, 3, gi, o) + snap_at(p, 3, gn, o)) / 2;
        s.out[4][q] = (snap_at(p, 4, gi, o) + snap_at(p, 4, gi, o + sj)) / 2;
        s.out[5][q] = (snap_at(p, 5, gi, o) + snap_at(p, 5, gi, o + sk)) / 2;
    }
}

// ------------------------------------------------------------------------------------------
// Linked x-slab shards: halo planes are PUSHED into the neighbour's ghost plane with peer stores over NVLink and announced
// with monotonic flags (value = iteration + 1) in the neighbour's memory; the neighbour's stream holds a one-thread kernel
// that waits for the flag.  No host round trip and no collective on the data path, so a shard's whole iteration is a fixed
// kernel sequence (the expected flag values come from the device iteration counter) that is captured into one CUDA graph.
// ------------------------------------------------------------------------------------------
enum { GPB_FLAG_H_READY = 0, GPB_FLAG_E_READY = 1, GPB_FLAG_H_FREE = 2, GPB_FLAG_SNAP_READY = 3, GPB_FLAG_SNAP_DONE = 4,
       GPB_FLAG_TIMEOUT = 8, GPB_FLAG_PUSH_COUNT = 9, GPB_NFLAGS = 16 };

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// wait until *flag >= *iter + add (one thread).  A wait that lasts longer than `timeout_ns` gives up and records itself in
// my_flags[GPB_FLAG_TIMEOUT] (checked by the host after the run): a lost neighbour must not hang the device.
static __global__ void k_flag_wait(const unsigned *flag, const int *iter, int add, unsigned *my_flags, int which, unsigned long long timeout_ns)
{
    const unsigned want = (unsigned)(*iter + add);
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (ld_acquire_sys(flag) < want) {
        __nanosleep(200);
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) {
            atomicOr(my_flags + GPB_FLAG_TIMEOUT, 1u << which);
            break;
        }
    }
}

// *flag = *iter + add in (peer) memory, after everything this stream did before
static __global__ void k_flag_signal(unsigned *flag, const int *iter, int add)
{
    __threadfence_system();
    st_release_sys(flag, (unsigned)(*iter + add));
}

// copy two planes (n elements each, 16-byte aligned) into the neighbour's ghost planes with 128-bit peer stores; the last
// block to finish publishes *flag = *iter + add
template <typename R>
__global__ void __launch_bounds__(256) k_halo_push(const R *__restrict__ src_a, const R *__restrict__ src_b, R *__restrict__ dst_a, R *__restrict__ dst_b,
                                                   long long n, unsigned *flag, const int *iter, int add, unsigned *counter)
{
    const long long nv = n * (long long)sizeof(R) / 16;
    const uint4 *sa = reinterpret_cast<const uint4 *>(src_a), *sb = reinterpret_cast<const uint4 *>(src_b);
    uint4 *da = reinterpret_cast<uint4 *>(dst_a), *db = reinterpret_cast<uint4 *>(dst_b);
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nv; q += (long long)gridDim.x * blockDim.x) {
        da[q] = sa[q];
        db[q] = sb[q];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(counter, 1u) == gridDim.x - 1) {
            *counter = 0;
            __threadfence_system();
            st_release_sys(flag, (unsigned)(*iter + add));
        }
    }
}

// uint32 host IDs -> narrow device IDs with z pitch; records the largest ID seen.
template <typename IDT>
__global__ void k_narrow_ids(const uint32_t *src, IDT *dst, long long rows, int nzp1, int pitch, unsigned *maxid)
{
    unsigned mx = 0;
    const long long n = rows * nzp1;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const long long row = q / nzp1;
        const int k = (int)(q - row * nzp1);
        const uint32_t v = src[q];
        dst[row * pitch + k] = (IDT)v;
        mx = max(mx, v);
    }
    if (mx) atomicMax(maxid, mx);
}

// IDs narrowed on the host (dense rows of nzp1 elements) -> device layout with z pitch
template <typename IDT>
__global__ void k_place_ids(const IDT *src, IDT *dst, long long rows, int nzp1, int pitch)
{
    const long long n = rows * nzp1;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const long long row = q / nzp1;
        dst[row * pitch + (int)(q - row * nzp1)] = src[q];
    }
}

template <typename IDT>
__global__ void k_fill_ids(IDT *dst, long long n, IDT v)
{
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) dst[q] = v;
}

}  // namespace gpb
