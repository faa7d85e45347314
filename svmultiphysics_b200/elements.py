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


def nxx_tables(eNoN: int):
    """Second parametric derivatives Nxx(6,eNoN,nG) of the shape functions (fs[0].Nxx of the reference, handed to
    nn::gn_nxx, Code/Source/solver/nn.cpp:1172-1283), Voigt order (00, 11, 22, 01, 12, 02).  TET4: identically zero."""
    if eNoN == 4:
        return np.zeros((6, 4, 4), order="F")
    if eNoN != 8:
        raise ValueError(f"no second-derivative table for eNoN={eNoN}")
    s = 1.0 / np.sqrt(3.0)
    gp = np.array([[-s, -s, -s], [s, -s, -s], [s, s, -s], [-s, s, -s],
                   [-s, -s, s], [s, -s, s], [s, s, s], [-s, s, s]])
    sign = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1],
                     [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=float)
    Nxx = np.zeros((6, 8, 8), order="F")
    for g in range(8):
        lx, ly, lz = gp[g]
        for a in range(8):
            sx, sy, sz = sign[a]
            Nxx[3, a, g] = sx * sy * (1.0 + sz * lz) / 8.0
            Nxx[4, a, g] = sy * sz * (1.0 + sx * lx) / 8.0
            Nxx[5, a, g] = sx * sz * (1.0 + sy * ly) / 8.0
    return Nxx


def tri3_face_tables(qm: float = 2.0 / 3.0):
    """TRI3 boundary face (of TET4): Code/Source/solver/nn_elem_gip.h:720-738 (points, qmTRI3 = 2/3), shape functions
    N = (xi0, xi1, 1 - xi0 - xi1) as evaluate_face_basis_values_and_gradients leaves them (pinned by tests against the
    compiled reference).  Returns w(3), N(3,3), Nx(2,3,3)."""
    s, t = qm, -0.5 * qm + 0.5
    xi = np.array([[t, s, t], [t, t, s]])
    w = np.full(3, 1.0 / 6.0)
    N = np.zeros((3, 3), order="F")
    N[0], N[1], N[2] = xi[0], xi[1], 1.0 - xi[0] - xi[1]
    Nx = np.zeros((2, 3, 3), order="F")
    for g in range(3):
        Nx[:, :, g] = np.array([[1.0, 0.0, -1.0], [0.0, 1.0, -1.0]])
    return w, N, Nx


def quad4_face_tables():
    """QUD4 boundary face (of HEX8): 2x2 Gauss, bilinear shape functions in VTK order."""
    s = 1.0 / np.sqrt(3.0)
    gp = np.array([[-s, -s], [s, -s], [s, s], [-s, s]])
    sign = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], dtype=float)
    w = np.ones(4)
    N = np.zeros((4, 4), order="F")
    Nx = np.zeros((2, 4, 4), order="F")
    for g in range(4):
        lx, ly = gp[g]
        for a in range(4):
            sx, sy = sign[a]
            N[a, g] = (1.0 + sx * lx) * (1.0 + sy * ly) / 4.0
            Nx[0, a, g] = sx * (1.0 + sy * ly) / 4.0
            Nx[1, a, g] = (1.0 + sx * lx) * sy / 4.0
    return w, N, Nx


def face_tables(eNoNb: int):
    if eNoNb == 3:
        return tri3_face_tables()
    if eNoNb == 4:
        return quad4_face_tables()
    raise ValueError(f"no face table for eNoNb={eNoNb}")


def tables(eNoN: int):
    if eNoN == 4:
        return tet4_tables()
    if eNoN == 8:
        return hex8_tables()
    raise ValueError(f"no reference-element table for eNoN={eNoN}")
