"""Helpers around the vendored reference install (baseline/_ref, git-ignored; see install_ref.sh)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')


def have_reference():
    return os.path.isdir(os.path.join(REF_DIR, 'gprMax'))


def use_reference():
    """Make the pip-installed, unmodified reference importable (`import gprMax`) together with the stand-ins for its
    missing optional dependencies.  Returns the install directory."""
    if not have_reference():
        raise RuntimeError('baseline/_ref is missing: run baseline/install_ref.sh in the build container')
    from . import standins
    standins.install()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    return REF_DIR
