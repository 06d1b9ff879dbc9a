/*
 * gprmax_b200.h -- C ABI of the B200-native FDTD time-stepping core for gprMax.
 *
 * This library replaces the reference's PyCUDA solver path and nothing else:
 *
 *   gprMax/model_build_run.py:477-716   solve_gpu()            -> gpb_create / gpb_run / gpb_get_* / gpb_destroy
 *   gprMax/fields_updates_gpu.py:40-240 update_e / update_h / update_e_dispersive_A,B
 *   gprMax/pml_updates/pml_updates_{electric,magnetic}_{HORIPML,MRIPML}_gpu.py   (48 slab kernels)
 *   gprMax/source_updates_gpu.py:38-200 Hertzian dipole / magnetic dipole / voltage source
 *   gprMax/sources.py:286-452           TransmissionLine (CPU-only in the reference)
 *   gprMax/fields_outputs.py:67-108     store_outputs (receiver gather, + Ix/Iy/Iz of grid.py:413-461)
 *   gprMax/snapshots_gpu.py:31-77       store_snapshot (semantics of snapshots.py:87-130 / snapshots_ext.pyx)
 *   gprMax/grid.py:247-272              FDTDGrid.gpu_* (device arrays, launch geometry)
 *   gprMax/utilities.py:341-413         GPU / detect_check_gpus -> gpb_device_count / gpb_device_info
 *
 * The reference has no FFI for this path (its seam is the Python call
 * `solve_gpu(currentmodelrun, modelend, G)`); the host side stays Python and binds these entry
 * points with ctypes (gprmax_b200/_lib.py; INTEGRATION.md shows the two-line patch to the
 * reference).  Only plain pointers and sizes cross the boundary.
 *
 * Conventions
 *   - All host arrays are C-contiguous and laid out as the reference holds them
 *     (fields/ID: [nx+1][ny+1][nz+1], z contiguous; grid.py:168-190).  The library copies what it
 *     needs during gpb_create and keeps no host pointer afterwards.
 *   - `dtype` selects the floating type R of every `const void*` table and of all outputs:
 *     GPB_F32 -> float / complex64, GPB_F64 -> double / complex128 (constants.py:36-49).
 *   - Every function returns 0 on success, non-zero on failure; gpb_last_error() gives the
 *     message (thread-local).  The library never exits or aborts the process.
 *   - There is no CPU fallback: without a CUDA device gpb_create fails.
 *   - x-slab sharding: a handle owns the node planes i in [x_start, x_start + nx_planes) of a
 *     global grid of nx_global cells; all array arguments then hold ONLY those planes (coordinates
 *     of sources/receivers/PML slabs stay global).  Single GPU: x_start = 0, nx_planes = nx + 1.
 */
#ifndef GPRMAX_B200_H
#define GPRMAX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPB_ABI_VERSION 3

enum { GPB_F32 = 0, GPB_F64 = 1 };
enum { GPB_HORIPML = 0, GPB_MRIPML = 1 };                       /* pml.py:155 */
enum { GPB_XMINUS = 0, GPB_YMINUS, GPB_ZMINUS, GPB_XPLUS, GPB_YPLUS, GPB_ZPLUS }; /* pml.py:163 */
enum { GPB_SRC_HERTZIAN = 0, GPB_SRC_MAGNETIC = 1, GPB_SRC_VOLTAGE = 2 };
/* receiver output rows, receivers.py:29 */
enum { GPB_EX = 0, GPB_EY, GPB_EZ, GPB_HX, GPB_HY, GPB_HZ, GPB_IX, GPB_IY, GPB_IZ, GPB_NRXOUT };

/* utilities.py:341-366 (class GPU) */
typedef struct gpb_device_info_t {
    int32_t device_id;
    char name[256];
    char pci_bus_id[32];
    uint64_t total_mem;      /* bytes */
    uint64_t const_mem;      /* bytes */
    int32_t sm_count;
    int32_t cc_major, cc_minor;
} gpb_device_info_t;

/* One PML slab: pml.py:149-274.  R tables are R[order][thickness]. */
typedef struct gpb_pml_t {
    int32_t direction;                 /* GPB_XMINUS .. GPB_ZPLUS */
    int32_t xs, xf, ys, yf, zs, zf;    /* global cell coordinates of the slab */
    int32_t thickness;
    double d;                          /* spacing along the slab axis */
    const void *ERA, *ERB, *ERE, *ERF;
    const void *HRA, *HRB, *HRE, *HRF;
} gpb_pml_t;

/* Point source: source_updates_gpu.py / sources.py:71-232, packed like sources.py:235-283. */
typedef struct gpb_source_t {
    int32_t kind;                      /* GPB_SRC_* */
    int32_t i, j, k;                   /* global node coordinates */
    int32_t polarisation;              /* 0 x, 1 y, 2 z */
    int32_t it_first, it_last;         /* inclusive iteration range in which  start <= it*dt <= stop */
    double param;                      /* Hertzian: dl ; voltage: resistance (0 = hard source) ; magnetic: unused */
    const void *waveform;              /* R[iterations]: whole-step values (Hertzian, resistive voltage) or
                                          half-step values (magnetic dipole, hard voltage source) */
} gpb_source_t;

/* Transmission line: sources.py:286-452, state as left by calculate_incident_V_I (:326-346). */
typedef struct gpb_tline_t {
    int32_t i, j, k, polarisation;
    int32_t it_first, it_last;
    int32_t nl, srcpos, antpos;
    double resistance, dl;
    double abcv0, abcv1;
    const void *voltage0, *current0;   /* R[nl] initial line state */
    const void *wave_whole, *wave_half;/* R[iterations] */
} gpb_tline_t;

/* Snapshot: snapshots.py:28-84.  Output arrays are R[nx][ny][nz] per component. */
typedef struct gpb_snapshot_t {
    int32_t xs, ys, zs, xf, yf, zf;
    int32_t dx, dy, dz;
    int32_t nx, ny, nz;
    int32_t time;                      /* taken when time == iteration + 1 (model_build_run.py:430) */
} gpb_snapshot_t;

typedef struct gpb_model_t {
    int32_t abi_version;               /* GPB_ABI_VERSION */
    int32_t dtype;                     /* GPB_F32 / GPB_F64 */
    int32_t nx, ny, nz;                /* global cells */
    int32_t x_start, nx_planes;        /* owned node planes (see sharding note above) */
    double dx, dy, dz, dt;
    int32_t iterations;
    int32_t nmaterials;
    const uint32_t *ID;                /* [6][nx_planes][ny+1][nz+1]; NULL = homogeneous: every edge is `uniform_id` */
    int64_t id_comp_stride;            /* elements between the six components of ID; 0 = dense (nx_planes*(ny+1)*(nz+1)).  A shard
                                          may point into the caller's global G.ID: ID = &G.ID[0][x_start][0][0] with the global
                                          component stride (nx+1)*(ny+1)*(nz+1) -- no second host copy of the slab */
    int32_t uniform_id;                /* used only when ID == NULL (synthetic multi-billion-cell domains) */
    const void *updatecoeffsE;         /* R[nmaterials][5]  materials.py:200 */
    const void *updatecoeffsH;         /* R[nmaterials][5]  materials.py:201 */
    int32_t maxpoles;                  /* Material.maxpoles */
    const void *updatecoeffsdispersive;/* C[nmaterials][3*maxpoles] materials.py:204-208, or NULL */
    int32_t pml_formulation;           /* GPB_HORIPML / GPB_MRIPML */
    int32_t pml_order;                 /* len(G.cfs): 1 or 2 */
    int32_t npml;
    const gpb_pml_t *pmls;             /* in G.pmls order (x0,y0,z0,xmax,ymax,zmax; grid.py:132-136) */
    int32_t nsources;
    const gpb_source_t *sources;       /* applied in array order within each kind */
    int32_t ntlines;
    const gpb_tline_t *tlines;
    int32_t nrx;
    const int32_t *rxcoords;           /* int32[nrx][3] global node coordinates */
    int32_t nsnapshots;
    const gpb_snapshot_t *snapshots;
} gpb_model_t;

typedef struct gpb_solver *gpb_handle;

/* ---- device discovery (utilities.py:369-413) ---- */
int gpb_device_count(int *count);
int gpb_device_info(int device_id, gpb_device_info_t *out);

/* ---- lifetime ---- */
int gpb_create(const gpb_model_t *model, int device_id, gpb_handle *out);
int gpb_destroy(gpb_handle h);

/* ---- time loop (model_build_run.py:590-696) ----
 * gpb_run advances `n_iters` full iterations (rx store, snapshots, H half, E half) starting at the
 * handle's current iteration; the loop time accumulates in gpb_elapsed_seconds (CUDA events,
 * as the reference times it, :586-588/:708-710). */
int gpb_run(gpb_handle h, int n_iters);
int gpb_iteration(gpb_handle h, int *iteration);
int gpb_elapsed_seconds(gpb_handle h, double *seconds);
int gpb_mem_used(gpb_handle h, uint64_t *bytes);        /* device bytes owned by this handle */
int gpb_kernel_launches(gpb_handle h, uint64_t *count); /* kernels launched by gpb_run/gpb_half_step so far */
int gpb_reset(gpb_handle h);                            /* zero fields / PML / T / rx, iteration = 0 */
/* Next trace of a B-scan on a FIXED geometry (the reference's --geometry-fixed: the same FDTDGrid is solved again with
 * stepped sources / receivers, model_build_run.py:294-330): take the sources, transmission lines, receivers and snapshots
 * of `model` (grid, iteration count and material table must be the resident ones; ID / coefficient / PML arguments are
 * ignored), clear all field state, keep the ID and coefficient arrays on the device. */
int gpb_set_points(gpb_handle h, const gpb_model_t *model);
/* Measurement aid: advance n_iters iterations with plain launches and CUDA events between the
 * kernels; ms4 = device milliseconds {step prologue, H update, E update, source kernels} summed. */
int gpb_profile(gpb_handle h, int n_iters, double *ms4);
/* Which kernel each half-step runs on, e.g. "H:k_update_tma<14x64,PHASE=0> E:k_update_tma<14x64,PHASE=1>" (the
 * library picks the TMA-staged, register-vectorised or scalar family from the grid size). */
int gpb_kernel_path(gpb_handle h, char *buf, size_t buf_bytes);

/* ---- sharded stepping (one handle per x-slab; the host moves the halo planes between calls) ----
 * phase 0: rx store + snapshots + H half-step (needs the Ey,Ez ghost plane at x_start+nx_planes)
 * phase 1: E half-step                        (needs the Hy,Hz ghost plane at x_start-1)
 * part  0: step prologue (phase 0 only) and the ONE plane whose result the neighbour needs
 *          (phase 0: last owned plane, phase 1: first owned plane) -- after this call the send halo
 *          of gpb_halo() is final (unless a point source sits on that plane; then send after part 1)
 * part  1: all remaining planes and the point sources of the phase
 * part -1: both.  All launches go to the handle's stream (gpb_stream); nothing here synchronises. */
int gpb_half_step(gpb_handle h, int phase, int part);
/* Device pointers (and byte size) of the contiguous halo planes:
 * which = 0: send  E (Ey then Ez, first owned plane)      -> left neighbour's recv E
 * which = 1: recv  E (Ey then Ez, ghost plane after last)
 * which = 2: send  H (Hy then Hz, last owned plane)       -> right neighbour's recv H
 * which = 3: recv  H (Hy then Hz, ghost plane before first) */
int gpb_halo(gpb_handle h, int which, void **dptr_a, void **dptr_b, size_t *bytes_each);
int gpb_stream(gpb_handle h, void **cuda_stream);
int gpb_synchronize(gpb_handle h);

/* ---- one domain over several GPUs of ONE process (the reference rejects models larger than one GPU, grid.py:239-241) ----
 * gpb_create_sharded takes the GLOBAL model (x_start = 0, nx_planes = nx + 1, ID = the whole G.ID or NULL) and the
 * device list of `-gpu 0 1 ...`; it cuts the nx + 1 node planes into contiguous x-slabs (one per device, slab by slab
 * straight out of the caller's ID array), links neighbouring slabs over peer memory and returns ONE handle: gpb_run,
 * gpb_get_receivers / _snapshot / _tline / _field, gpb_reset, gpb_mem_used ... then act on the whole domain.  Results are
 * bit-identical to a single-GPU run of the same model. */
int gpb_create_sharded(const gpb_model_t *global_model, const int *device_ids, int ndevices, gpb_handle *out);

/* ---- linking shards that live in DIFFERENT processes (one process per GPU, e.g. under torchrun) ----
 * Every rank creates its own slab handle (gpb_create with x_start / nx_planes), publishes gpb_link_info() to its two
 * neighbours (any byte transport: the struct is plain data and carries CUDA IPC handles) and calls gpb_link with what it
 * received.  From then on gpb_run advances the shard with the halo planes pushed into the neighbours' ghost planes by peer
 * stores over NVLink and announced by flags in peer memory: no host round trip, no collective, the whole iteration is one
 * CUDA graph.  gpb_link(h, NULL, NULL) unlinks.  Snapshots and transmission lines on a cut plane are supported in linked
 * mode only (they need the neighbour's planes). */
typedef struct gpb_link_t {
    uint64_t process_id;               /* owner process */
    int32_t device_id;                 /* CUDA runtime ordinal in the owner process */
    int32_t dtype;
    int32_t x_start, nx_planes, ny, nz;
    uint64_t plane_elems, array_elems; /* elements per (padded) plane / per component array incl. the two ghost planes */
    uint64_t fields_ptr, flags_ptr;    /* device addresses, valid inside the owner process */
    unsigned char fields_ipc[64];      /* cudaIpcMemHandle_t of the driver allocation that holds the field arrays ... */
    unsigned char flags_ipc[64];       /* ... and of the one that holds the flag words (small allocations share a block) */
    uint64_t fields_ipc_offset, flags_ipc_offset;   /* byte offsets of the arrays inside those allocations */
} gpb_link_t;
int gpb_link_info(gpb_handle h, gpb_link_t *out);
int gpb_link(gpb_handle h, const gpb_link_t *left, const gpb_link_t *right);

/* ---- results ---- */
/* out: R[GPB_NRXOUT][iterations][nrx] (rows of fields_outputs.py:81-105 + Ix,Iy,Iz).  For a shard,
 * receivers outside the owned planes are left zero. */
int gpb_get_receivers(gpb_handle h, void *out, size_t out_bytes);
/* out: 6 arrays R[nx][ny][nz] (Ex,Ey,Ez,Hx,Hy,Hz cell-centred averages) of snapshot `index`. */
int gpb_get_snapshot(gpb_handle h, int index, void *out6[6], size_t bytes_each);
/* out: R[iterations] total voltage / current of transmission line `index` (fields_outputs.py:62-64) */
int gpb_get_tline(gpb_handle h, int index, void *vtotal, void *itotal, size_t bytes_each);
/* Debug/parity: copy one field component (GPB_EX..GPB_HZ) back in the host layout
 * R[nx_planes][ny+1][nz+1]; gpb_set_field is the inverse (tests seed random fields). */
int gpb_get_field(gpb_handle h, int component, void *out, size_t out_bytes);
int gpb_set_field(gpb_handle h, int component, const void *in, size_t in_bytes);

/* Device memory freed by gpb_destroy is kept per device and reused by the next gpb_create with arrays of the same size (a
 * B-scan creates one solver per trace; the reference pays a fresh CUDA context and allocations per model run,
 * model_build_run.py:492-497, 713-714).  gpb_release_cached returns it to the driver; GPB_NO_POOL=1 disables the cache. */
int gpb_release_cached(void);

/* ---- host-side build of the per-edge ID array (yee_cell_build_ext.pyx:110-257), all host cores ----
 * Arrays in the reference's layout: solid uint32[nx][ny][nz], rigidE int8[12][nx][ny][nz], rigidH int8[6][nx][ny][nz]
 * (grid.py:157-169), ID uint32[6][nx+1][ny+1][nz+1] in/out (edges marked rigid keep what the geometry commands wrote).
 * gpb_ids_scan writes every non-rigid edge of the planes [x0, x1) whose surrounding cells agree and returns the DISTINCT
 * disagreeing combinations of cell materials, each with the first edge that shows it, sorted in the reference's scan order
 * (component, i, j, k).  The caller resolves them in that order with create_electric_average / create_magnetic_average
 * (:31-107) -- which is what fixes the numbering of the averaged materials -- and gpb_ids_apply writes the result on every
 * disagreeing edge.  Return: 0 ok, 1 bad argument, 2 more than max_combos combinations (*ncombos = the number needed),
 * 3 a combination met in gpb_ids_apply was not in the table. */
typedef struct gpb_idcombo_t {
    uint32_t id[4];                    /* numID1..4 as the reference passes them (magnetic components: 2 used) */
    int32_t comp, i, j, k;             /* component (0 Ex .. 5 Hz) and first edge in scan order */
} gpb_idcombo_t;
int gpb_ids_scan(const uint32_t *solid, const int8_t *rigidE, const int8_t *rigidH, uint32_t *ID, int nx, int ny, int nz, int x0, int x1,
                 gpb_idcombo_t *combos, int max_combos, int *ncombos);
int gpb_ids_apply(const uint32_t *solid, const int8_t *rigidE, const int8_t *rigidH, uint32_t *ID, int nx, int ny, int nz, int x0, int x1,
                  const gpb_idcombo_t *combos, const uint32_t *numid, int ncombos);
/* The same two passes for a caller that holds only a slab of the arrays (one rank of a sharded build): solid / rigidE /
 * rigidH carry the cell planes [solid_x0, solid_x0 + solid_nx) and ID the node planes [id_x0, id_x0 + id_nx) of the nx x ny x nz
 * domain; x0, x1 and the edges returned in `combos` are global plane indices.  The arrays must cover the node planes [x0, x1)
 * and the cell planes [x0 - 1, x1) (an edge looks at the cells on both sides of its plane), else 1 is returned. */
int gpb_ids_scan_slab(const uint32_t *solid, const int8_t *rigidE, const int8_t *rigidH, uint32_t *ID, int nx, int ny, int nz,
                      int solid_x0, int solid_nx, int id_x0, int id_nx, int x0, int x1, gpb_idcombo_t *combos, int max_combos, int *ncombos);
int gpb_ids_apply_slab(const uint32_t *solid, const int8_t *rigidE, const int8_t *rigidH, uint32_t *ID, int nx, int ny, int nz,
                       int solid_x0, int solid_nx, int id_x0, int id_nx, int x0, int x1, const gpb_idcombo_t *combos, const uint32_t *numid, int ncombos);

/* ---- host-side re-layout for the streaming VTK writers (snapshots.py:128-167, geometry_outputs.py:119-205,
 * geometry_outputs_ext.pyx:81-110), all host cores ----
 * ParaView order (x fastest, z slowest, components interleaved) out of z-fastest arrays:
 *   out[((k * count[1] + j) * count[0] + i) * ncomp + c] =
 *       src[c][(start[0] + i*step[0]) * stride[0] + (start[1] + j*step[1]) * stride[1] + (start[2] + k*step[2]) * stride[2]]
 * for i < count[0], j < count[1], k < count[2]; strides in elements, elements of 1, 2, 4 or 8 bytes (moved as bit patterns).
 * A caller streams a file by asking for one range of k at a time.  Return: 0 ok, 1 bad argument. */
int gpb_vtk_transpose(const void *const *src, int ncomp, int elem_bytes, const int64_t stride[3], const int32_t start[3],
                      const int32_t count[3], const int32_t step[3], void *out);

const char *gpb_last_error(void);
const char *gpb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GPRMAX_B200_H */
