"""Inert stand-ins for three pure-Python dependencies of the reference that are not installed in this image
(colorama, terminaltables, h5py).  None of them is on the solver path: they colour stdout, draw a table and write the
.out file.  On a machine that has the real packages these stand-ins are never installed (`install()` only fills gaps).

The h5py stand-in implements the small write-side subset `fields_outputs.write_hdf5_outputfile` (fields_outputs.py:111-161)
uses -- `File(name, 'w')`, `.attrs[...]`, `create_group`, item assignment of arrays -- and stores the same tree (same
group / dataset / attribute names) as a NumPy .npz container under the requested file name, so a B-scan run through the
reference's front end leaves one readable output per trace; `read_out(path)` returns {'attrs': {...}, 'data': {...}}.
"""
import sys
import types

import numpy as np


class _Attrs(dict):
    pass


class _Group(object):
    def __init__(self, root, path):
        self._root, self._path = root, path.rstrip('/')
        self.attrs = root._attrs.setdefault(self._path or '/', _Attrs())

    def create_group(self, name):
        path = name if name.startswith('/') else self._path + '/' + name
        return _Group(self._root, path)

    def __setitem__(self, name, value):
        path = name if name.startswith('/') else self._path + '/' + name
        self._root._data[path] = np.asarray(value)

    def __getitem__(self, name):
        path = name if name.startswith('/') else self._path + '/' + name
        if path in self._root._data:
            return self._root._data[path]
        return _Group(self._root, path)


class File(_Group):
    def __init__(self, name, mode='r'):
        self._name, self._mode = str(name), mode
        self._attrs, self._data = {}, {}
        if mode.startswith('r') or mode == 'a':
            try:
                loaded = read_out(self._name)
                self._attrs = {k: _Attrs(v) for k, v in loaded['attrs'].items()}
                self._data = dict(loaded['data'])
            except FileNotFoundError:
                if mode.startswith('r'):
                    raise
        _Group.__init__(self, self, '')

    def close(self):
        if self._mode.startswith('r'):
            return
        out = {}
        for g, attrs in self._attrs.items():
            for k, v in attrs.items():
                out['attr:' + g + ':' + k] = np.asarray(v)
        for k, v in self._data.items():
            out['data:' + k] = v
        with open(self._name, 'wb') as f:
            np.savez(f, **out)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            if not self._mode.startswith('r') and (self._attrs or self._data):
                self.close()
        except Exception:
            pass


def read_out(path):
    """Contents of an output file written through the h5py stand-in."""
    z = np.load(path, allow_pickle=False)
    attrs, data = {}, {}
    for k in z.files:
        if k.startswith('attr:'):
            _, g, name = k.split(':', 2)
            attrs.setdefault(g, {})[name] = z[k]
        elif k.startswith('data:'):
            data[k[5:]] = z[k]
    return {'attrs': attrs, 'data': data}


def install():
    """Register stand-ins for whichever of colorama / terminaltables / h5py cannot be imported."""
    import importlib
    for name in ('colorama', 'terminaltables', 'h5py'):
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
            continue
        except ImportError:
            pass
        m = types.ModuleType(name)
        m.__standin__ = True
        if name == 'colorama':
            class _Blank(object):
                def __getattr__(self, attr):
                    return ''
            m.init = lambda *a, **k: None
            m.deinit = lambda *a, **k: None
            m.Fore, m.Style, m.Back = _Blank(), _Blank(), _Blank()
        elif name == 'terminaltables':
            class AsciiTable(object):
                def __init__(self, data, title=None):
                    self.table_data = data
                    self.title = title
                    self.outer_border = True
                    self.justify_columns = {}
                    self.inner_heading_row_border = True

                @property
                def table(self):
                    return '\n'.join(' | '.join(str(c) for c in row) for row in self.table_data)
            m.AsciiTable = AsciiTable
            m.SingleTable = AsciiTable
        else:
            m.File = File
            m.read_out = read_out
        sys.modules[name] = m
