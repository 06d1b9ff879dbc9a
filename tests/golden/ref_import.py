"""Import the UNMODIFIED reference (gprMax v3.1.7, /root/reference) in this container.

Only used by the golden-vector generator (make_golden.py) and by the
`needs_reference` tests; it never runs on the GPU box (no /root/reference there).

The reference tree is read-only and ships no compiled extensions, so the `gprMax`
package is assembled from two search locations:
    oracle/_ref/<f32|f64>/gprMax   compiled Cython kernels (oracle/build_ref.py --full)
    /root/reference/gprMax         the reference's own Python sources
Three pure-Python dependencies that are not installed here (colorama,
terminaltables, h5py) are replaced by inert stand-ins; none of them is on the
solver path (they colour stdout, draw a table and write the .out file).
"""
import importlib.machinery
import importlib.util
import os
import sys
import types

REF = os.environ.get('GPRMAX_REFERENCE', '/root/reference')
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _install_standins():
    if 'colorama' not in sys.modules:
        m = types.ModuleType('colorama')

        class _Blank(object):
            def __getattr__(self, name):
                return ''
        m.init = lambda *a, **k: None
        m.Fore = _Blank()
        m.Style = _Blank()
        m.Back = _Blank()
        sys.modules['colorama'] = m
    if 'terminaltables' not in sys.modules:
        m = types.ModuleType('terminaltables')

        class AsciiTable(object):
            def __init__(self, data, title=None):
                self.table_data = data
                self.outer_border = True
                self.justify_columns = {}

            @property
            def table(self):
                return '\n'.join(' | '.join(str(c) for c in row) for row in self.table_data)
        m.AsciiTable = AsciiTable
        m.SingleTable = AsciiTable
        sys.modules['terminaltables'] = m
    if 'h5py' not in sys.modules:
        m = types.ModuleType('h5py')

        def _nofile(*a, **k):
            raise RuntimeError('h5py stand-in: the golden generator never writes .out files')
        m.File = _nofile
        sys.modules['h5py'] = m


def import_reference(variant='f32'):
    """Returns the reference's top-level `gprMax` package (float32 or float64 build)."""
    if 'gprMax' in sys.modules:
        have = getattr(sys.modules['gprMax'], '_graft_variant', None)
        if have != variant:
            raise RuntimeError('reference already imported as {}; one precision per process'.format(have))
        return sys.modules['gprMax']
    extdir = os.path.join(ROOT, 'oracle', '_ref', variant, 'gprMax')
    if not os.path.isdir(REF):
        raise RuntimeError('reference tree not present at ' + REF)
    if not os.path.isdir(extdir):
        raise RuntimeError('run `python oracle/build_ref.py --full` first')
    _install_standins()
    # user_libs (antenna macros) are imported by some .in files
    if REF not in sys.path:
        sys.path.append(REF)
    spec = importlib.util.spec_from_file_location(
        'gprMax', os.path.join(REF, 'gprMax', '__init__.py'),
        submodule_search_locations=[extdir, os.path.join(REF, 'gprMax')])
    mod = importlib.util.module_from_spec(spec)
    mod._graft_variant = variant
    sys.modules['gprMax'] = mod
    # the sub-package pml_updates needs the same two-location treatment
    pspec = importlib.util.spec_from_file_location(
        'gprMax.pml_updates', os.path.join(REF, 'gprMax', 'pml_updates', '__init__.py'),
        submodule_search_locations=[os.path.join(extdir, 'pml_updates'), os.path.join(REF, 'gprMax', 'pml_updates')])
    pmod = importlib.util.module_from_spec(pspec)
    sys.modules['gprMax.pml_updates'] = pmod
    pspec.loader.exec_module(pmod)
    mod.pml_updates = pmod
    spec.loader.exec_module(mod)
    return mod
