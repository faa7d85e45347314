"""Reference-element tables (Gauss weights, shape functions, parametric gradients) for the element
types on the hot path, in the reference's layout: w(nG), N(eNoN,nG), Nx(3,eNoN,nG), column-major.

TET4: Code/Source/solver/nn_elem_gip.h:214-226 (points), nn.cpp:174 ("origin-last" node order).
HEX8: Code/Source/solver/nn_elem_gip.h:13-50, VTK node order.
These are what nn::select_ele + fs::init_fs_msh leave in mshType.{w,N,Nx}; tests pin them against the
compiled reference.
"""
from __future__ import annotations

import numpy as np


def tet4_tables():
    s = (5.0 + 3.0 * np.sqrt(5.0)) / 20.0
    t = (5.0 - np.sqrt(5.0)) / 20.0
    xi = np.full((3, 4), t)
    for g in range(3):
        xi[g, g] = s
    w = np.full(4, 1.0 / 24.0)
    N = np.zeros((4, 4), order="F")
    N[0], N[1], N[2] = xi[0], xi[1], xi[2]
    N[3] = 1.0 - xi[0] - xi[1] - xi[2]
    Nx = np.zeros((3, 4, 4), order="F")
    for g in range(4):
        Nx[:, :, g] = np.array([[1.0, 0.0, 0.0, -1.0], [0.0, 1.0, 0.0, -1.0], [0.0, 0.0, 1.0, -1.0]])
    return w, N, Nx


def hex8_tables():
    s = 1.0 / np.sqrt(3.0)
    gp = np.array([[-s, -s, -s], [s, -s, -s], [s, s, -s], [-s, s, -s],
                   [-s, -s, s], [s, -s, s], [s, s, s], [-s, s, s]])
    sign = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1],
                     [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=float)
    w = np.ones(8)
    N = np.zeros((8, 8), order="F")
    Nx = np.zeros((3, 8, 8), order="F")
    for g in range(8):
        lx, ly, lz = gp[g]
        for a in range(8):
            sx, sy, sz = sign[a]
            fx, fy, fz = 1.0 + sx * lx, 1.0 + sy * ly, 1.0 + sz * lz
            N[a, g] = fx * fy * fz / 8.0
            Nx[0, a, g] = sx * fy * fz / 8.0
            Nx[1, a, g] = fx * sy * fz / 8.0
            Nx[2, a, g] = fx * fy * sz / 8.0
    return w, N, Nx


def tables(eNoN: int):
    if eNoN == 4:
        return tet4_tables()
    if eNoN == 8:
        return hex8_tables()
    raise ValueError(f"no reference-element table for eNoN={eNoN}")
