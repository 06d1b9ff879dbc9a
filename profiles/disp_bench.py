"""Dispersive half-step at a size that leaves L2: BASELINE.json configs[2] (heterogeneous_soil.in, 50-bin Peplinski soil = 1-pole
Debye everywhere, rough surface, 6 PML slabs) scaled up by tiling the reference-built 150 x 150 x 100 ID array of the
full-size fixture (material / dispersive / PML tables are the reference's own; dt only depends on the cell size).

    python profiles/disp_bench.py [tiles_x tiles_y tiles_z] [iterations]      default 2 2 3 -> 300 x 300 x 300, 200 iterations
Environment: GPB_DISP_V4=1 (register-vectorised E kernel), GPB_DISP_COMPLEX=1 (complex T), GPB_TMA_TPF=0/1/2 (T prefetch distance).
Prints one JSON line: Mcells/s, per-kernel ms, algorithmic roofline fraction (96 + 48 P + PML bytes per cell-step)."""
import json, os, sys
sys.path.insert(0, ".")
import numpy as np
from gprmax_b200 import Solver, load_model

args = [int(a) for a in sys.argv[1:]]
tx, ty, tz = (args + [2, 2, 3])[:3] if len(args) >= 3 else (2, 2, 3)
its = args[3] if len(args) > 3 else 200
G, _ = load_model('tests/golden/heterogeneous_soil_full_f32.npz')
nx0, ny0, nz0 = G.nx, G.ny, G.nz
ID = np.asarray(G.ID)
cellpart = ID[:, :nx0, :ny0, :nz0]
big = np.tile(cellpart, (1, tx, ty, tz))
big = np.pad(big, ((0, 0), (0, 1), (0, 1), (0, 1)), mode='wrap')
G.ID = np.ascontiguousarray(big)
G.nx, G.ny, G.nz = nx0 * tx, ny0 * ty, nz0 * tz
for p in G.pmls:                                   # same thickness and tables, new extents
    t = p.thickness
    ext = {'xminus': (0, t, 0, G.ny, 0, G.nz), 'yminus': (0, G.nx, 0, t, 0, G.nz), 'zminus': (0, G.nx, 0, G.ny, 0, t),
           'xplus': (G.nx - t, G.nx, 0, G.ny, 0, G.nz), 'yplus': (0, G.nx, G.ny - t, G.ny, 0, G.nz), 'zplus': (0, G.nx, 0, G.ny, G.nz - t, G.nz)}[p.direction]
    p.xs, p.xf, p.ys, p.yf, p.zs, p.zf = ext
G.iterations = its
for s in G.hertziandipoles:
    s.waveformvalues_wholestep = s.waveformvalues_wholestep[:its]
    s.waveformvalues_halfstep = s.waveformvalues_halfstep[:its]
    s.xcoord, s.ycoord, s.zcoord = G.nx // 3, G.ny // 2, int(G.nz * 0.85)
for r in G.rxs:
    r.xcoord, r.ycoord, r.zcoord = 2 * G.nx // 3, G.ny // 2, int(G.nz * 0.85)
cells = G.nx * G.ny * G.nz
P = int(G.maxpoles)
S = sum(p.thickness * {'x': G.ny * G.nz, 'y': G.nx * G.nz, 'z': G.nx * G.ny}[p.direction[0]] for p in G.pmls)
b_alg = 96.0 + 48.0 * P + 32.0 * len(G.cfs) * S / cells
with Solver(G, device_id=0) as sv:
    path = sv.kernel_path
    sv.run(); sv.reset(); sv.run()
    t = sv.elapsed
    rx = sv.receivers()
    sv.reset()
    sv.profile(20)
    prof = sv.profile(its - 20)
mc = cells * its / t / 1e6
peak = 6456.8
try:
    peak = float(json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'])
except Exception:
    pass
print(json.dumps({'model': 'heterogeneous_soil tiled {}x{}x{} -> {} x {} x {} cells, {} materials, maxpoles {}'.format(tx, ty, tz, G.nx, G.ny, G.nz, G.updatecoeffsE.shape[0], P),
                  'iterations': its, 'kernels': path, 'mcells_per_s': mc, 'us_per_iteration': t / its * 1e6,
                  'kernel_ms_per_iteration': {k: v / (its - 20) for k, v in prof.items()},
                  'alg_bytes_per_cell_step': b_alg, 'alg_gbs': mc * 1e6 * b_alg / 1e9, 'frac_of_peak': mc * 1e6 * b_alg / 1e9 / peak,
                  'e_kernel_alg_gbs': cells * (b_alg / 2 + 24.0 * P) / (prof['update_e'] / (its - 20) * 1e-3) / 1e9,
                  'rx_checksum': float(np.abs(rx).sum()), 'env': {k: v for k, v in os.environ.items() if k.startswith('GPB_')}}))
