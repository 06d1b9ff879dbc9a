#!/usr/bin/env python
"""Generate the parity fixtures in tests/golden/ by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference and `python oracle/build_ref.py
--full`).  For each model below it runs the reference end to end on its CPU solver
(`gprMax.run(<file>.in)` -> run_model -> solve_cpu, model_build_run.py:84-474), hooks
`solve_cpu` to capture the fully built FDTDGrid `G`, and stores
    * everything the time loop reads from G (gprmax_b200.model_io.save_model), and
    * the reference's outputs: every receiver trace, transmission-line V/I, snapshots
as one compressed .npz per model.  The GPU tests load these on a box where the
reference does not exist.

Usage:  python tests/golden/make_golden.py [f32|f64] [name ...]
One precision per process (the reference's precision is a module-level constant).
"""
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

REF = os.environ.get('GPRMAX_REFERENCE', '/root/reference')

# name -> (source .in relative to the reference or inline text, extra lines, precisions)
# Inline models are small variants that exercise code paths no shipped small model covers.
MODELS = {}


def ref_file(name, rel, precisions=('f32', 'f64'), append='', replace=None, trace_only=(), suffix=None):
    """trace_only: precisions for which only the reference's traces are stored (no second copy of a big grid);
    suffix: file name suffix per precision when it is not the precision itself."""
    MODELS[name] = dict(rel=rel, text=None, precisions=precisions, append=append, replace=replace or {},
                        trace_only=tuple(trace_only), suffix=suffix or {})


def inline(name, text, precisions=('f32', 'f64')):
    MODELS[name] = dict(rel=None, text=text, precisions=precisions, append='', replace={}, trace_only=(), suffix={})


# --- the reference's own basic test models (tests/test_models.py:47) ----------------------
ref_file('cylinder_Ascan_2D', 'user_models/cylinder_Ascan_2D.in')
ref_file('2D_ExHyHz', 'tests/models_basic/2D_ExHyHz/2D_ExHyHz.in')
ref_file('2D_EyHxHz', 'tests/models_basic/2D_EyHxHz/2D_EyHxHz.in')
ref_file('2D_EzHxHy', 'tests/models_basic/2D_EzHxHy/2D_EzHxHy.in')
ref_file('hertzian_dipole_fs', 'tests/models_basic/hertzian_dipole_fs/hertzian_dipole_fs.in')
ref_file('hertzian_dipole_hs', 'tests/models_basic/hertzian_dipole_hs/hertzian_dipole_hs.in')
ref_file('hertzian_dipole_dispersive', 'tests/models_basic/hertzian_dipole_dispersive/hertzian_dipole_dispersive.in')
ref_file('magnetic_dipole_fs', 'tests/models_basic/magnetic_dipole_fs/magnetic_dipole_fs.in')

_BOX = """#title: {title}
#domain: 0.040 0.036 0.032
#dx_dy_dz: 0.001 0.001 0.001
#time_window: 300
"""

# --- PML variants on a small non-cubic box with a dielectric block crossing the PML ------
for form in ('HORIPML', 'MRIPML'):
    for order in (1, 2):
        cfs = '#pml_cfs: constant forward 0 0 constant forward 1 1 quartic forward 0 None\n'
        if order == 2:
            cfs = ('#pml_cfs: constant forward 0.05 0.05 quartic forward 1 7 quartic forward 0 None\n'
                   '#pml_cfs: constant forward 0 0 constant forward 1 1 sextic forward 0 0.5\n')
        inline('pml_{}_{}'.format(form, order), _BOX.format(title='PML ' + form + ' order ' + str(order)) + """#pml_formulation: {form}
#pml_cells: 6 7 8 9 6 5
{cfs}#material: 4 0.01 1 0 slab
#box: 0 0 0 0.040 0.036 0.012 slab
#waveform: gaussiandotnorm 1 8e9 w
#hertzian_dipole: y 0.024 0.020 0.016 w
#rx: 0.012 0.010 0.009
#rx: 0.030 0.028 0.020
#rx: 0.003 0.004 0.003
""".format(form=form, cfs=cfs))

# --- the reference's own PML test models (tests/models_pmls/pml_3D_pec_plate: the elongated thin PEC plate of the PML
#     literature, 51 x 126 x 26 cells, receiver 3 cells from the PML): CFS with alpha and kappa > 1, 'reverse' scaling profiles,
#     two CFS per PML, a reduced time step (#time_step_stability_factor), #python blocks.  Shortened from 2100 to 700 iterations
#     (the wave runs along the 100-cell plate and back in ~350).
for _v, _f in (('CFS', 'CFS-PML'), ('HORIPML_2', 'HORIPML-2'), ('MRIPML_2', 'MRIPML-2')):
    ref_file('pec_plate_' + _v, 'tests/models_pmls/pml_3D_pec_plate/pml_3D_pec_plate_{}.in'.format(_f), replace={'#time_window: 2100': '#time_window: 700'})

# --- sources: resistive + hard voltage source, magnetic dipole, start/stop gating, Ix/Iy/Iz outputs
inline('sources_mixed', _BOX.format(title='mixed sources') + """#material: 3 0.005 1.5 0.1 stuff
#box: 0.010 0.008 0.006 0.030 0.026 0.020 stuff
#waveform: gaussian 1 6e9 wg
#waveform: ricker 2 8e9 wr
#voltage_source: z 0.020 0.018 0.010 50 wg
#voltage_source: x 0.024 0.020 0.022 0 wr 5e-11 3.0e-10
#magnetic_dipole: y 0.014 0.018 0.016 wr
#hertzian_dipole: x 0.012 0.010 0.024 wg 2e-11 4.0e-10
#rx: 0.020 0.018 0.016 rxa Ex Ey Ez Hx Hy Hz Ix Iy Iz
#rx: 0.026 0.012 0.008 rxb Ez Hx Iy
""")

# --- transmission line feeding a short PEC dipole (CPU-only feature in the reference) -----
inline('transmission_line', _BOX.format(title='transmission line') + """#waveform: gaussian 1 6e9 wg
#transmission_line: z 0.020 0.018 0.016 73 wg
#edge: 0.020 0.018 0.006 0.020 0.018 0.016 pec
#edge: 0.020 0.018 0.017 0.020 0.018 0.026 pec
#rx: 0.028 0.018 0.016
""")

# --- dispersive: 2-pole Debye + Lorentz + Drude in one model (multipole path), PML on ----
inline('dispersive_multipole', _BOX.format(title='multi-pole dispersive') + """#material: 3 0.001 1 0 deb
#add_dispersion_debye: 2 5 1e-10 3 4e-11 deb
#material: 2 0 1 0 lor
#add_dispersion_lorentz: 1 4 1.2e10 2e9 lor
#material: 1 0 1 0 dru
#add_dispersion_drude: 1 8e9 1e9 0 dru
#box: 0 0 0 0.040 0.036 0.010 deb
#box: 0.005 0.005 0.012 0.018 0.030 0.024 lor
#sphere: 0.028 0.020 0.020 0.006 dru
#waveform: gaussiandotnorm 1 8e9 w
#hertzian_dipole: z 0.020 0.018 0.016 w
#rx: 0.012 0.010 0.006
#rx: 0.010 0.012 0.018
#rx: 0.028 0.020 0.020
""")

# --- snapshots (strided and offset sub-volumes) -------------------------------------------
inline('snapshots', _BOX.format(title='snapshots') + """#waveform: ricker 1 8e9 w
#hertzian_dipole: z 0.020 0.018 0.016 w
#rx: 0.012 0.010 0.009
#snapshot: 0 0 0 0.040 0.036 0.032 0.001 0.001 0.001 120 snapA
#snapshot: 0.004 0.006 0.002 0.036 0.030 0.030 0.002 0.003 0.004 200 snapB
""")

# --- the published benchmark model, smallest size (tests/benchmarking/bench_100x100x100.in),
#     shortened: pins the synthetic free-space builder used by bench.py
ref_file('bench_100', 'tests/benchmarking/bench_100x100x100.in', replace={'#time_window: 3e-9': '#time_window: 120'})

# --- config 3 scaled down: heterogeneous_soil.in with explicit seeds, smaller box, fewer steps
ref_file('heterogeneous_soil_small', 'user_models/heterogeneous_soil.in', replace={
    '#domain: 0.15 0.15 0.1': '#domain: 0.06 0.06 0.05',
    '#time_window: 6e-9': '#time_window: 250',
    '#rx: 0.105 0.075 0.085': '#rx: 0.040 0.030 0.040',
    '#hertzian_dipole: y 0.045 0.075 0.085 my_ricker': '#hertzian_dipole: y 0.020 0.030 0.040 my_ricker',
    '#fractal_box: 0 0 0 0.15 0.15 0.070 1.5 1 1 1 50 my_soil my_soil_box': '#fractal_box: 0 0 0 0.06 0.06 0.035 1.5 1 1 1 50 my_soil my_soil_box 7',
    '#add_surface_roughness: 0 0 0.070 0.15 0.15 0.070 1.5 1 1 0.065 0.080 my_soil_box': '#add_surface_roughness: 0 0 0.035 0.06 0.06 0.035 1.5 1 1 0.030 0.040 my_soil_box 3',
    '#geometry_view': '##geometry_view',
})

# --- FULL-SIZE fixtures of BASELINE.json configs 2-4 (minutes each on 8 cores; used by dedicated GPU tests) ----------
# config 2: the headline benchmark model itself, all 1559 iterations -- only the reference's receiver trace is stored
# (gprmax_b200.synthetic.bench_model rebuilds the grid bit-identically, tests/test_synthetic.py)
ref_file('bench_300_trace', 'tests/benchmarking/bench_300x300x300.in', precisions=('f32', 'f64'), trace_only=('f32', 'f64'))
# config 3 at full size: heterogeneous_soil.in with explicit fractal seeds (the shipped file has none, so its geometry
# differs from run to run); float64 run stored as traces only ("truth" for compare_f32_with_truth)
ref_file('heterogeneous_soil_full', 'user_models/heterogeneous_soil.in', replace={
    '#fractal_box: 0 0 0 0.15 0.15 0.070 1.5 1 1 1 50 my_soil my_soil_box': '#fractal_box: 0 0 0 0.15 0.15 0.070 1.5 1 1 1 50 my_soil my_soil_box 7',
    '#add_surface_roughness: 0 0 0.070 0.15 0.15 0.070 1.5 1 1 0.065 0.080 my_soil_box': '#add_surface_roughness: 0 0 0.070 0.15 0.15 0.070 1.5 1 1 0.065 0.080 my_soil_box 3',
    '#geometry_view': '##geometry_view',
}, trace_only=('f64',), suffix={'f64': 'f64_truth'})
# config 4: trace 1 of the GSSI 1.5 GHz B-scan (current_model_run = 1), the reference's own input file unchanged
ref_file('bscan_gssi_trace1', 'user_models/cylinder_Bscan_GSSI_1500.in', precisions=('f32',))


def model_text(spec):
    if spec['text'] is not None:
        return spec['text']
    with open(os.path.join(REF, spec['rel'])) as f:
        text = f.read()
    for a, b in spec['replace'].items():
        if a not in text:
            raise RuntimeError('pattern not found in {}: {}'.format(spec['rel'], a))
        text = text.replace(a, b)
    return text + '\n' + spec['append']


def generate(name, variant, gprMax, outdir=HERE):
    import gprMax.model_build_run as mbr
    from gprmax_b200.model_io import save_model
    spec = MODELS[name]
    work = tempfile.mkdtemp(prefix='golden_')
    cap = {}
    orig = mbr.solve_cpu

    def hooked(cur, end, G):
        from gprMax.materials import Material
        G.maxpoles = Material.maxpoles
        if os.environ.get('GOLDEN_GEOMETRY_ONLY'):
            cap['G'], cap['tl0'] = G, []
            return 0.0
        # TL start state must be captured before the loop mutates it
        cap['tl0'] = [(t.voltage[:t.nl].copy(), t.current[:t.nl].copy(), t.abcv0, t.abcv1) for t in G.transmissionlines]
        cap['G'] = G
        return orig(cur, end, G)

    try:
        infile = os.path.join(work, name + '.in')
        with open(infile, 'w') as f:
            f.write(model_text(spec))
        mbr.solve_cpu = hooked
        mbr.write_hdf5_outputfile = lambda f, G: None
        for s in ('write_vtk_imagedata',):
            pass
        import gprMax.snapshots as snapmod
        snapmod.Snapshot.write_vtk_imagedata = lambda self, pbar, G: None
        gprMax.run(infile)
        G = cap['G']
        golden = {}
        for n, rx in enumerate(G.rxs):
            for k, v in rx.outputs.items():
                golden['rx{}_{}'.format(n, k)] = np.asarray(v)
        for n, tl in enumerate(G.transmissionlines):
            golden['tl{}_Vtotal'.format(n)] = tl.Vtotal
            golden['tl{}_Itotal'.format(n)] = tl.Itotal
        for n, snap in enumerate(G.snapshots):
            golden['snap{}_electric'.format(n)] = snap.electric
            golden['snap{}_magnetic'.format(n)] = snap.magnetic
        # restore TL start state so the fixture describes the model at loop entry
        for t, (v, c, a0, a1) in zip(G.transmissionlines, cap['tl0']):
            t.voltage = v
            t.current = c
            t.abcv0, t.abcv1 = a0, a1
        from gprMax._version import __version__
        meta = dict(reference_version=__version__, source=spec['rel'] or 'inline (tests/golden/make_golden.py)',
                    variant=variant, numpy=np.__version__)
        path = os.path.join(outdir, '{}_{}.npz'.format(name, spec['suffix'].get(variant, variant)))
        if os.environ.get('GOLDEN_GEOMETRY_ONLY'):
            # regenerate only the built grid and compare it with the committed fixture (bit-exact ID / tables)
            from gprmax_b200.model_io import load_model
            Gold, _ = load_model(path)
            same = all(np.array_equal(np.asarray(getattr(G, a)), np.asarray(getattr(Gold, a))) for a in ('ID', 'updatecoeffsE', 'updatecoeffsH'))
            print('wrote nothing: rebuilt grid {} the committed {}'.format('MATCHES' if same else 'DIFFERS FROM', os.path.basename(path)))
            return
        if variant in spec['trace_only']:
            np.savez_compressed(path, **{'golden_' + k: v for k, v in golden.items()})
        else:
            save_model(G, path, golden=golden, meta=meta)
        print('wrote {} ({:.1f} kB)'.format(path, os.path.getsize(path) / 1e3))
    finally:
        mbr.solve_cpu = orig
        shutil.rmtree(work, ignore_errors=True)


def main(argv):
    """One fresh interpreter per model: the reference keeps process-global state between models
    (`Material.maxpoles` is a class attribute that is never reset, materials.py:28), so a model
    run after a dispersive one in the same process would silently take the dispersive code path."""
    import subprocess
    variant = 'f32'
    names = []
    one = '--one' in argv
    for a in argv:
        if a in ('f32', 'f64'):
            variant = a
        elif not a.startswith('--'):
            names.append(a)
    if one:
        from ref_import import import_reference
        gprMax = import_reference(variant)
        generate(names[0], variant, gprMax)
        return
    for name in (names or list(MODELS)):
        if variant in MODELS[name]['precisions'] or names:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), '--one', variant, name],
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
            tail = [l for l in r.stdout.splitlines() if l.startswith('wrote ')]
            print(tail[-1] if tail and r.returncode == 0 else 'FAILED {} {}\n{}'.format(name, variant, r.stdout[-2000:]))


if __name__ == '__main__':
    main(sys.argv[1:])
