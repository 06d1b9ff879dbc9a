"""Vectorised `build_pmls` -- drop-in for the reference's gprMax/pml.py:367-429.

The reference averages the relative permittivity / permeability of the cells under every PML slab face with a Python double loop
that searches the material list for every cell (`next(x for x in G.materials if x.numID == numID)`): O(face cells x materials)
interpreted steps -- minutes for the 2048 x 1024 faces of a multi-billion-cell domain.  Here the face is gathered through a
per-material table and summed with np.cumsum, i.e. in the SAME left-to-right order as the reference's `sumer += material.er`
(np.sum would add pairwise and differ in the last bits), so the averages, sigma max and every R table are bit-identical
(tests/test_yee_build.py).  Everything else is the reference's own code: the `PML` objects and `calculate_update_coeffs`
(pml.py:149-274) are created and called exactly as the reference does.
"""
import numpy as np


def build_pmls(G, pbar):
    from gprMax.pml import PML

    nmax = max(m.numID for m in G.materials) + 1
    er = np.zeros(nmax, dtype=np.float64)
    mr = np.zeros(nmax, dtype=np.float64)
    for m in G.materials:
        er[m.numID], mr[m.numID] = m.er, m.mr

    def face_average(face):
        ids = np.ascontiguousarray(face).ravel()          # C order = the reference's loop order (outer index first)
        return float(np.cumsum(er[ids])[-1]) / ids.size, float(np.cumsum(mr[ids])[-1]) / ids.size

    for key, value in G.pmlthickness.items():
        if value <= 0:
            continue
        if key == 'x0':
            pml = PML(G, ID=key, direction='xminus', xf=value, yf=G.ny, zf=G.nz)
        elif key == 'xmax':
            pml = PML(G, ID=key, direction='xplus', xs=G.nx - value, xf=G.nx, yf=G.ny, zf=G.nz)
        elif key == 'y0':
            pml = PML(G, ID=key, direction='yminus', yf=value, xf=G.nx, zf=G.nz)
        elif key == 'ymax':
            pml = PML(G, ID=key, direction='yplus', ys=G.ny - value, xf=G.nx, yf=G.ny, zf=G.nz)
        elif key == 'z0':
            pml = PML(G, ID=key, direction='zminus', zf=value, xf=G.nx, yf=G.ny)
        else:
            pml = PML(G, ID=key, direction='zplus', zs=G.nz - value, xf=G.nx, yf=G.ny, zf=G.nz)
        G.pmls.append(pml)
        if key[0] == 'x':
            averageer, averagemr = face_average(G.solid[pml.xs, :, :])
        elif key[0] == 'y':
            averageer, averagemr = face_average(G.solid[:, pml.ys, :])
        else:
            averageer, averagemr = face_average(G.solid[:, :, pml.zs])
        pml.calculate_update_coeffs(averageer, averagemr, G)
        pbar.update()
