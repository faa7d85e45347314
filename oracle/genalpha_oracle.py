"""TEST INFRASTRUCTURE ONLY — numpy restatement of the generalised-alpha state updates of the reference.
Only tests/ and __graft_entry__.smoke() may import this; the product never does.

Parity pinning: since round 2 oracle/_ref/libsvref.so contains the reference's Integrator.cpp and set_bc.cpp (unmodified; harness
oracle/ref_harness_genalpha.cpp, binding oracle.refbind.GenAlphaRef).  tests/golden/genalpha.npz is generated from that compiled
reference and tests/test_genalpha_cpu.py checks that this restatement reproduces it BIT FOR BIT; the device kernels are compared
with the same fixtures in tests/test_gpu_genalpha.py.  Each line cites the statement it follows in
/root/reference/Code/Source/solver/Integrator.cpp.  Arrays are (tDof, nNo).
"""
import numpy as np


USTRUCT, FSI = 7, 2   # svb200_phys


def is_sst(q):
    """com_mod.sstEq for this equation's update: a ustruct equation, or an FSI equation flagged SVB200_EQTIME_SSTEQ (ustruct solids)."""
    return q.phys == USTRUCT or (q.phys == FSI and (q.reserved & 1))


def predictor(eqs, dt, dFlag, Ao, Yo, Do, An, Yn, Dn, Ad=None):
    """Integrator::predictor, Integrator.cpp:540-643 (state part); a ustruct equation takes the sstEq branch :626-630."""
    for q in eqs:
        r = slice(q.s, q.e + 1)
        coef = (q.gam - 1.0) / q.gam                      # :551
        An[r] = Ao[r] * coef                              # :555  eqn 87 of Bazilevs 2007
        Yn[r] = Yo[r]                                     # :616  eqn 86
        if dFlag and is_sst(q):
            Ad *= (q.gam - 1.0) / q.gam                   # :627-628
            Dn[r] = Do[r]                                 # :629
        elif dFlag:                                       # :618-623
            c = dt * dt * (0.5 * q.gam - q.beta) / (q.gam - 1.0)
            Dn[r] = (Do[r] + Yn[r] * dt) + An[r] * c
        else:
            Dn[r] = Do[r]                                 # :640


def initiator(eqs, Ao, Yo, Do, An, Yn, Dn, Ag, Yg, Dg):
    """Integrator::initiator, Integrator.cpp:704-741."""
    for q in eqs:
        r = slice(q.s, q.e + 1)
        c0, c1, c2, c3 = 1.0 - q.am, q.am, 1.0 - q.af, q.af      # :709-712
        Ag[r] = Ao[r] * c0 + An[r] * c1                   # :735  eqn 89
        Yg[r] = Yo[r] * c2 + Yn[r] * c3                   # :738  eqn 90
        Dg[r] = Do[r] * c2 + Dn[r] * c3                   # :740


def corrector(q, dt, R, An, Yn, Dn, mesh_s=-1, solid=None):
    """Integrator::corrector, Integrator.cpp:812-815 (coefficients), :861-872 (update), :887-912 (FSI copy)."""
    c0, c1 = q.gam * dt, q.beta * dt * dt
    n = q.e - q.s + 1
    r = slice(q.s, q.e + 1)
    An[r] = An[r] - R[:n]                                 # :864  eqn 94
    Yn[r] = Yn[r] - R[:n] * c0                            # :867  eqn 95
    Dn[r] = Dn[r] - R[:n] * c1                            # :869
    if mesh_s >= 0 and solid is not None:
        m = np.asarray(solid, bool)
        for X in (An, Yn, Dn):
            X[mesh_s:mesh_s + 3, m] = X[0:3, m]           # :905-909


def corrector_ustruct(q, dt, R, Rd, An, Yn, Dn, Ad, mesh_s=-1, solid=None):
    """Integrator::corrector for a ustruct equation or an FSI equation with ustruct solids (sstEq), Integrator.cpp:812-815
    (coefficients), :826-846, and the FSI copy :887-912."""
    c0, c2 = q.gam * dt, 1.0 / q.am
    c3 = q.af * c0 * c2
    r = slice(q.s, q.e + 1)
    An[r] = An[r] - R[:4]                                 # :832
    Yn[r] = Yn[r] - R[:4] * c0                            # :833
    dUl = Rd * c2 + R[:3] * c3                            # :837
    Ad -= dUl                                             # :838
    Dn[q.s:q.s + 3] = Dn[q.s:q.s + 3] - dUl * c0          # :839
    if mesh_s >= 0 and solid is not None:
        m = np.asarray(solid, bool)
        for X in (An, Yn, Dn):
            X[mesh_s:mesh_s + 3, m] = X[0:3, m]           # :905-909
