/* sv_oracle.c — CPU restatement of the svMultiPhysics hot path (TEST INFRASTRUCTURE ONLY).
 *
 * Plain C11 restatement, routine by routine, of the reference algorithms the CUDA library replaces.
 * It exists so that parity can be checked where the compiled reference (oracle/_ref/libsvref.so, the
 * unmodified sources) is not available, and it is itself pinned against that library by
 * tests/test_oracle_cpu.py and by the golden vectors under tests/golden/ (generated from libsvref.so by
 * tests/golden/make_golden.py).  Nothing in the product (svmultiphysics_b200/, libsvb200.so) links,
 * imports or calls this file.
 *
 * Unlike the CUDA kernels it follows the reference's own evaluation order Gauss point by Gauss point
 * (no hoisting), so it is an independent check of the device algebra.
 *
 * Restated routines (file:line into /root/reference):
 *   lhsa_ns::add_col / lhsa            Code/Source/solver/lhsa.cpp:13-54, 126-381
 *   nn::gnn (insd = 3)                 Code/Source/solver/nn.cpp:862-899
 *   fluid::construct_fluid             Code/Source/solver/fluid.cpp:480-762
 *   fluid::fluid_3d_m                  Code/Source/solver/fluid.cpp:1768-2237
 *   fluid::fluid_3d_c                  Code/Source/solver/fluid.cpp:1443-1760
 *   fluid::get_viscosity               Code/Source/solver/fluid.cpp:2240-2298
 *   utils::is_zero                     Code/Source/solver/utils.cpp:141-160
 *   all_fun::domain                    Code/Source/solver/all_fun.cpp:122-148
 *   lhsa_ns::do_assem                  Code/Source/solver/lhsa.cpp:70-114
 *   fsils_solve                        Code/Source/linear_solver/solve.cpp:23-166
 *   precond::precond_diag, pre/pos_mul Code/Source/linear_solver/precond.cpp:95-242, 534-611, 19-94
 *   gmres::gmres_v                     Code/Source/linear_solver/gmres.cpp:425-609
 *   cgrad::cgrad_v                     Code/Source/linear_solver/cgrad.cpp:139-219
 *   bicgs::bicgsv                      Code/Source/linear_solver/bicgs.cpp:22-120
 *   spar_mul::fsils_spar_mul_vv        Code/Source/linear_solver/spar_mul.cpp:164-231
 *   dot::fsils_nc_dot_v, norm::fsi_ls_normv, omp_la::omp_sum_v/omp_mul_v
 *                                      Code/Source/linear_solver/dot.cpp:107, norm.cpp:40, omp_la.cpp:21-122
 *   add_bc_mul (BCOP_TYPE_ADD)         Code/Source/linear_solver/add_bc_mul.cpp:26-124
 * Single partition only (the multi-rank exchange steps are identities on one rank).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "../include/svb200.h"

#define MAXE 8
#define MAXG 8

typedef struct {
  int eNoN, nEl, nG, lShpF;
  int *IEN, *eId;
  double w[MAXG], N[MAXE * MAXG], Nx[3 * MAXE * MAXG];
} OMesh;

typedef struct {
  int set, bGrp, dof, nNo;
  int* glob;
  double *val, *valM;
  int incFlag, coupledFlag;
  double res, nS;
} OFace;

typedef struct {
  int nNo, nnz, dof, tDof, nMsh, nFaces;
  double *x, *Bf, *Ag, *Yg, *Dg;
  OMesh msh[8];
  int *rowPtr, *colPtr, *diagPtr;
  double *R, *Val;
  OFace* face;
  double last_assemble_s, last_solve_s;
} OCase;

static _Thread_local char g_err[512];
static int fail(const char* msg) { snprintf(g_err, sizeof g_err, "%s", msg); return SVB200_ERR_NUMERIC; }
static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

const char* svorc_last_error(void) { return g_err; }
void* svorc_create(void) { return calloc(1, sizeof(OCase)); }

void svorc_destroy(void* h)
{
  OCase* c = h;
  if (!c) return;
  free(c->x); free(c->Bf); free(c->Ag); free(c->Yg); free(c->Dg);
  for (int i = 0; i < c->nMsh; i++) { free(c->msh[i].IEN); free(c->msh[i].eId); }
  free(c->rowPtr); free(c->colPtr); free(c->diagPtr); free(c->R); free(c->Val);
  for (int f = 0; f < c->nFaces; f++) { free(c->face[f].glob); free(c->face[f].val); free(c->face[f].valM); }
  free(c->face);
  free(c);
}

int svorc_set_coords(void* h, int nNo, const double* x)
{
  OCase* c = h;
  c->nNo = nNo;
  c->x = realloc(c->x, sizeof(double) * 3 * nNo);
  memcpy(c->x, x, sizeof(double) * 3 * nNo);
  c->Bf = realloc(c->Bf, sizeof(double) * 3 * nNo);
  memset(c->Bf, 0, sizeof(double) * 3 * nNo);
  return 0;
}

/* Reference-element tables: TET4 nn_elem_gip.h:214-226 + nn.cpp:174 ; HEX8 nn_elem_gip.h:13-50. */
static void fill_tables(OMesh* m)
{
  if (m->eNoN == 4) {
    const double s = (5.0 + 3.0 * sqrt(5.0)) / 20.0, t = (5.0 - sqrt(5.0)) / 20.0;
    m->nG = 4; m->lShpF = 1;
    for (int g = 0; g < 4; g++) {
      double xi[3] = {t, t, t};
      if (g < 3) xi[g] = s;
      m->w[g] = 1.0 / 24.0;
      m->N[0 + 4 * g] = xi[0]; m->N[1 + 4 * g] = xi[1]; m->N[2 + 4 * g] = xi[2];
      m->N[3 + 4 * g] = 1.0 - xi[0] - xi[1] - xi[2];
      const double d[4][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {-1, -1, -1}};
      for (int a = 0; a < 4; a++) for (int k = 0; k < 3; k++) m->Nx[k + 3 * (a + 4 * g)] = d[a][k];
    }
  } else {
    const double s = 1.0 / sqrt(3.0);
    const double sg[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};
    m->nG = 8; m->lShpF = 0;
    for (int g = 0; g < 8; g++) {
      const double lx = sg[g][0] * s, ly = sg[g][1] * s, lz = sg[g][2] * s;
      m->w[g] = 1.0;
      for (int a = 0; a < 8; a++) {
        const double fx = 1.0 + sg[a][0] * lx, fy = 1.0 + sg[a][1] * ly, fz = 1.0 + sg[a][2] * lz;
        m->N[a + 8 * g] = fx * fy * fz / 8.0;
        m->Nx[0 + 3 * (a + 8 * g)] = sg[a][0] * fy * fz / 8.0;
        m->Nx[1 + 3 * (a + 8 * g)] = fx * sg[a][1] * fz / 8.0;
        m->Nx[2 + 3 * (a + 8 * g)] = fx * fy * sg[a][2] / 8.0;
      }
    }
  }
}

int svorc_add_mesh(void* h, int eNoN, int nEl, const int* IEN, const int* eId, int nFn, const double* fN)
{
  OCase* c = h;
  (void)nFn; (void)fN;
  if (eNoN != 4 && eNoN != 8) return fail("[sv_oracle] only TET4 and HEX8 meshes are restated");
  if (c->nMsh >= 8) return fail("[sv_oracle] too many meshes");
  OMesh* m = &c->msh[c->nMsh++];
  m->eNoN = eNoN; m->nEl = nEl;
  m->IEN = malloc(sizeof(int) * eNoN * nEl);
  memcpy(m->IEN, IEN, sizeof(int) * eNoN * nEl);
  m->eId = NULL;
  if (eId) { m->eId = malloc(sizeof(int) * nEl); memcpy(m->eId, eId, sizeof(int) * nEl); }
  fill_tables(m);
  return 0;
}

int svorc_get_mesh_tables(void* h, int iM, int* nG, double* w, double* N, double* Nx)
{
  OCase* c = h;
  OMesh* m = &c->msh[iM];
  *nG = m->nG;
  if (w) memcpy(w, m->w, sizeof(double) * m->nG);
  if (N) memcpy(N, m->N, sizeof(double) * m->eNoN * m->nG);
  if (Nx) memcpy(Nx, m->Nx, sizeof(double) * 3 * m->eNoN * m->nG);
  return 0;
}

/* lhsa: sorted insertion per row (add_col), then compaction; fsils_lhs_create for one rank gives
 * diagPtr and the identity map. */
int svorc_build_graph(void* h, int nFaces, int* nnz_out)
{
  OCase* c = h;
  const int n = c->nNo;
  int cap = 40, grow;
  int* uInd = malloc(sizeof(int) * (size_t)cap * n);
  for (size_t i = 0; i < (size_t)cap * n; i++) uInd[i] = -1;
  do {
    grow = 0;
    for (int im = 0; im < c->nMsh && !grow; im++) {
      OMesh* m = &c->msh[im];
      for (int e = 0; e < m->nEl && !grow; e++)
        for (int a = 0; a < m->eNoN && !grow; a++) {
          const int row = m->IEN[a + m->eNoN * e];
          for (int b = 0; b < m->eNoN; b++) {
            const int col = m->IEN[b + m->eNoN * e];
            int* u = uInd + (size_t)cap * row;
            int i = 0;
            while (i < cap && u[i] != -1 && u[i] < col) i++;
            if (i < cap && u[i] == col) continue;
            if (u[cap - 1] != -1) { grow = 1; break; }
            for (int j = cap - 1; j > i; j--) u[j] = u[j - 1];
            u[i] = col;
          }
        }
    }
    if (grow) {   /* resiz: enlarge and start over (same final sorted lists) */
      cap += 40;
      uInd = realloc(uInd, sizeof(int) * (size_t)cap * n);
      for (size_t i = 0; i < (size_t)cap * n; i++) uInd[i] = -1;
    }
  } while (grow);
  int nnz = 0;
  for (int r = 0; r < n; r++) {
    if (uInd[(size_t)cap * r] == -1) { free(uInd); return fail("[sv_oracle] isolated node in lhsa"); }
    for (int i = 0; i < cap; i++) if (uInd[(size_t)cap * r + i] != -1) nnz++;
  }
  c->nnz = nnz;
  c->rowPtr = realloc(c->rowPtr, sizeof(int) * (n + 1));
  c->colPtr = realloc(c->colPtr, sizeof(int) * nnz);
  c->diagPtr = realloc(c->diagPtr, sizeof(int) * n);
  int j = 0;
  c->rowPtr[0] = 0;
  for (int r = 0; r < n; r++) {
    for (int i = 0; i < cap; i++) {
      const int col = uInd[(size_t)cap * r + i];
      if (col == -1) continue;
      if (col == r) c->diagPtr[r] = j;
      c->colPtr[j++] = col;
    }
    c->rowPtr[r + 1] = j;
  }
  free(uInd);
  c->nFaces = nFaces;
  c->face = calloc(nFaces > 0 ? nFaces : 1, sizeof(OFace));
  *nnz_out = nnz;
  return 0;
}

int svorc_get_graph(void* h, int* rowPtr, int* colPtr)
{
  OCase* c = h;
  memcpy(rowPtr, c->rowPtr, sizeof(int) * (c->nNo + 1));
  memcpy(colPtr, c->colPtr, sizeof(int) * c->nnz);
  return 0;
}

/* fsils_bc_create (linear_solver/bc.cpp:18-102), one rank. */
int svorc_set_face(void* h, int faIn, int bGrp, int face_dof, int nNo, const int* glob, const double* val)
{
  OCase* c = h;
  if (faIn >= c->nFaces) return fail("FSILS: faIn is exceeding lhs structure maximum number of faces");
  if (faIn <= -1) return fail("FSILS: faIn is smaller than zero");
  OFace* f = &c->face[faIn];
  free(f->glob); free(f->val); free(f->valM);
  f->set = 1; f->bGrp = bGrp; f->dof = face_dof; f->nNo = nNo;
  f->glob = malloc(sizeof(int) * (nNo > 0 ? nNo : 1));
  f->val = calloc((size_t)face_dof * (nNo > 0 ? nNo : 1), sizeof(double));
  f->valM = calloc((size_t)face_dof * (nNo > 0 ? nNo : 1), sizeof(double));
  memcpy(f->glob, glob, sizeof(int) * nNo);
  if (val) memcpy(f->val, val, sizeof(double) * face_dof * nNo);
  return 0;
}

int svorc_alloc(void* h, int dof)
{
  OCase* c = h;
  c->dof = dof;
  c->R = realloc(c->R, sizeof(double) * dof * c->nNo);
  c->Val = realloc(c->Val, sizeof(double) * dof * dof * (size_t)c->nnz);
  memset(c->R, 0, sizeof(double) * dof * c->nNo);
  memset(c->Val, 0, sizeof(double) * dof * dof * (size_t)c->nnz);
  return 0;
}

int svorc_set_state(void* h, int tDof, const double* Ag, const double* Yg, const double* Dg, const double* Bf)
{
  OCase* c = h;
  const size_t n = (size_t)tDof * c->nNo;
  if (tDof != c->tDof) {
    c->Ag = realloc(c->Ag, sizeof(double) * n); c->Yg = realloc(c->Yg, sizeof(double) * n); c->Dg = realloc(c->Dg, sizeof(double) * n);
    memset(c->Ag, 0, sizeof(double) * n); memset(c->Yg, 0, sizeof(double) * n); memset(c->Dg, 0, sizeof(double) * n);
    c->tDof = tDof;
  }
  if (Ag) memcpy(c->Ag, Ag, sizeof(double) * n);
  if (Yg) memcpy(c->Yg, Yg, sizeof(double) * n);
  if (Dg) memcpy(c->Dg, Dg, sizeof(double) * n);
  if (Bf) memcpy(c->Bf, Bf, sizeof(double) * 3 * c->nNo);
  return 0;
}

/* Old displacement (mesh-motion equation); kept for interface parity with the reference harness. */
int svorc_set_old_disp(void* h, int tDof, const double* Do)
{
  OCase* c = h;
  (void)tDof; (void)Do; (void)c;
  return 0;
}

#define REAL double
#include "fluid_gp.inc"
#undef REAL

/* all_fun::domain */
static int domain_of(const OMesh* m, const svb200_dmnparams* dmn, int nDmn, int e, int* err)
{
  int id = -1;
  for (int i = 0; i < nDmn; i++) { id = i; if (dmn[i].Id == -1) return id; }
  if (!m->eId) { *err = 1; return 0; }
  for (int i = 0; i < nDmn; i++) { id = i; if ((m->eId[e] >> dmn[i].Id) & 1) return id; }
  return id;
}

/* lhsa_ns::do_assem */
static void do_assem(OCase* c, int d, const int* eqN, const double* lK, const double* lR)
{
  const int dof = c->dof;
  for (int a = 0; a < d; a++) {
    const int rowN = eqN[a];
    if (rowN == -1) continue;
    for (int i = 0; i < dof; i++) c->R[i + dof * rowN] = c->R[i + dof * rowN] + lR[i + dof * a];
    for (int b = 0; b < d; b++) {
      const int colN = eqN[b];
      if (colN == -1) continue;
      int left = c->rowPtr[rowN], right = c->rowPtr[rowN + 1], ptr = (right + left) / 2;
      while (colN != c->colPtr[ptr]) {
        if (colN > c->colPtr[ptr]) left = ptr; else right = ptr;
        ptr = (right + left) / 2;
      }
      for (int i = 0; i < dof * dof; i++)
        c->Val[i + (size_t)dof * dof * ptr] = c->Val[i + (size_t)dof * dof * ptr] + lK[i + dof * dof * (a + d * b)];
    }
  }
}

/* fluid::construct_fluid for linear (lShpF) elements with equal-order VMS (nFs = 1). */
static int construct_fluid(OCase* c, const OMesh* m, const svb200_eqparams* eq, const svb200_dmnparams* dmn, int nDmn)
{
  const int eNoN = m->eNoN, tDof = c->tDof;
  if (eNoN != 4) return fail("[sv_oracle] construct_fluid is restated for TET4 only");
  if (c->dof != 4) return fail("[sv_oracle] fluid needs dof = 4");
  double xl[3 * MAXE], bfl[3 * MAXE], al[16 * MAXE], yl[16 * MAXE], lR[4 * MAXE], lK[16 * MAXE * MAXE], Nwx[3 * MAXE], ks[9];
  int ptr[MAXE];
  for (int e = 0; e < m->nEl; e++) {
    int err = 0;
    const int cDmn = domain_of(m, dmn, nDmn, e, &err);
    if (err) return fail("eId is not allocated");
    if (dmn[cDmn].phys != SVB200_PHYS_FLUID) continue;
    for (int a = 0; a < eNoN; a++) {
      const int Ac = m->IEN[a + eNoN * e];
      ptr[a] = Ac;
      for (int i = 0; i < 3; i++) { xl[i + 3 * a] = c->x[i + 3 * Ac]; bfl[i + 3 * a] = c->Bf[i + 3 * Ac]; }
      for (int i = 0; i < tDof; i++) { al[i + tDof * a] = c->Ag[i + (size_t)tDof * Ac]; yl[i + tDof * a] = c->Yg[i + (size_t)tDof * Ac]; }
    }
    memset(lR, 0, sizeof lR);
    memset(lK, 0, sizeof lK);
    double Jac = 0.0;
    for (int g = 0; g < m->nG; g++) {
      if (g == 0 || !m->lShpF) {
        gnn3(eNoN, m->Nx + 3 * eNoN * g, xl, Nwx, &Jac, ks);
        if (is_zero(Jac)) return fail("[construct_fluid] Jacobian for element is < 0.");
      }
      const double w = m->w[g] * Jac;
      fluid_3d_m(eq, &dmn[cDmn], eNoN, w, ks, m->N + eNoN * g, Nwx, al, yl, bfl, tDof, lR, lK);
    }
    for (int g = 0; g < m->nG; g++) {
      if (g == 0 || !m->lShpF) {
        gnn3(eNoN, m->Nx + 3 * eNoN * g, xl, Nwx, &Jac, ks);
        if (is_zero(Jac)) return fail("[construct_fluid] Jacobian for element is < 0.");
      }
      const double w = m->w[g] * Jac;
      fluid_3d_c(eq, &dmn[cDmn], eNoN, w, ks, m->N + eNoN * g, Nwx, al, yl, bfl, tDof, lR, lK);
    }
    do_assem(c, eNoN, ptr, lK, lR);
  }
  return 0;
}

int svorc_assemble(void* h, int iM, const svb200_eqparams* e, const svb200_dmnparams* dmn, int nDmn)
{
  OCase* c = h;
  if (iM < 0 || iM >= c->nMsh) return fail("[sv_oracle] bad mesh index");
  const double t0 = now_s();
  int rc;
  if (e->phys == SVB200_PHYS_FLUID) rc = construct_fluid(c, &c->msh[iM], e, dmn, nDmn);
  else rc = fail("[sv_oracle] physics not restated yet");
  c->last_assemble_s = now_s() - t0;
  return rc;
}

int svorc_get(void* h, int what, double* dst)
{
  OCase* c = h;
  if (what == SVB200_ARRAY_R) memcpy(dst, c->R, sizeof(double) * c->dof * c->nNo);
  else if (what == SVB200_ARRAY_VAL) memcpy(dst, c->Val, sizeof(double) * c->dof * c->dof * (size_t)c->nnz);
  else return fail("[sv_oracle] bad array id");
  return 0;
}

int svorc_put(void* h, int what, int dof, const double* src)
{
  OCase* c = h;
  if (dof != c->dof) svorc_alloc(h, dof);
  if (what == SVB200_ARRAY_R) memcpy(c->R, src, sizeof(double) * dof * c->nNo);
  else if (what == SVB200_ARRAY_VAL) memcpy(c->Val, src, sizeof(double) * dof * dof * (size_t)c->nnz);
  else return fail("[sv_oracle] bad array id");
  return 0;
}

/* ---- FSILS ---------------------------------------------------------------------------------------- */
static void spar_mul_vv(const OCase* c, int dof, const double* K, const double* U, double* KU)
{
  const int d2 = dof * dof;
  memset(KU, 0, sizeof(double) * dof * c->nNo);
  for (int i = 0; i < c->nNo; i++)
    for (int j = c->rowPtr[i]; j < c->rowPtr[i + 1]; j++) {
      const int col = c->colPtr[j];
      if (dof == 4) {
        const double* k = K + (size_t)16 * j; const double* u = U + 4 * (size_t)col; double* o = KU + 4 * (size_t)i;
        o[0] = o[0] + k[0] * u[0] + k[1] * u[1] + k[2] * u[2] + k[3] * u[3];
        o[1] = o[1] + k[4] * u[0] + k[5] * u[1] + k[6] * u[2] + k[7] * u[3];
        o[2] = o[2] + k[8] * u[0] + k[9] * u[1] + k[10] * u[2] + k[11] * u[3];
        o[3] = o[3] + k[12] * u[0] + k[13] * u[1] + k[14] * u[2] + k[15] * u[3];
      } else if (dof == 3) {
        const double* k = K + (size_t)9 * j; const double* u = U + 3 * (size_t)col; double* o = KU + 3 * (size_t)i;
        o[0] = o[0] + k[0] * u[0] + k[1] * u[1] + k[2] * u[2];
        o[1] = o[1] + k[3] * u[0] + k[4] * u[1] + k[5] * u[2];
        o[2] = o[2] + k[6] * u[0] + k[7] * u[1] + k[8] * u[2];
      } else {
        for (int l = 0; l < dof; l++) {
          double sum = 0.0;
          for (int k = 0; k < dof; k++) sum += K[(size_t)d2 * j + l * dof + k] * U[(size_t)dof * col + k];
          KU[(size_t)dof * i + l] = KU[(size_t)dof * i + l] + sum;
        }
      }
    }
}

static double nc_dot(int n, const double* U, const double* V)   /* fsils_nc_dot_v: node-by-node order */
{
  double r = 0.0;
  for (int i = 0; i < n; i++) r = r + U[i] * V[i];
  return r;
}
/* The reference adds the dof products of one node in a single expression; for dof<=4 that is the same
 * left-to-right order as the flat loop above. */
static double normv(int n, const double* U) { return sqrt(nc_dot(n, U, U)); }

static void add_bc_mul_add(OCase* c, int dof, const double* X, double* Y)
{
  for (int f = 0; f < c->nFaces; f++) {
    OFace* fa = &c->face[f];
    if (!fa->set || !fa->coupledFlag) continue;
    const int nsd = fa->dof < dof ? fa->dof : dof;
    double S = 0.0;
    for (int a = 0; a < fa->nNo; a++)
      for (int i = 0; i < nsd; i++) S = S + fa->valM[i + fa->dof * a] * X[i + dof * fa->glob[a]];
    S = fa->res * S;
    for (int a = 0; a < fa->nNo; a++)
      for (int i = 0; i < nsd; i++) Y[i + dof * fa->glob[a]] = Y[i + dof * fa->glob[a]] + fa->valM[i + fa->dof * a] * S;
  }
}

static void precond_diag(OCase* c, int dof, double* Val, double* R, double* W)
{
  const int n = c->nNo, d2 = dof * dof;
  for (int Ac = 0; Ac < n; Ac++)
    for (int i = 0; i < dof; i++) W[i + dof * Ac] = Val[(size_t)d2 * c->diagPtr[Ac] + i * dof + i];
  for (int i = 0; i < dof * n; i++) if (W[i] == 0.0) W[i] = 1.0;
  for (int i = 0; i < dof * n; i++) W[i] = 1.0 / sqrt(fabs(W[i]));
  for (int f = 0; f < c->nFaces; f++) {
    OFace* fa = &c->face[f];
    if (!fa->set || !fa->incFlag) continue;
    const int nd = fa->dof < dof ? fa->dof : dof;
    if (fa->bGrp == SVB200_BC_DIR)
      for (int a = 0; a < fa->nNo; a++)
        for (int i = 0; i < nd; i++) W[i + dof * fa->glob[a]] = W[i + dof * fa->glob[a]] * fa->val[i + fa->dof * a];
  }
  for (int Ac = 0; Ac < n; Ac++)      /* pre_mul */
    for (int j = c->rowPtr[Ac]; j < c->rowPtr[Ac + 1]; j++)
      for (int i = 0; i < dof; i++)
        for (int k = 0; k < dof; k++) Val[(size_t)d2 * j + i * dof + k] = Val[(size_t)d2 * j + i * dof + k] * W[i + dof * Ac];
  for (int i = 0; i < dof * n; i++) R[i] = W[i] * R[i];
  for (int Ac = 0; Ac < n; Ac++)      /* pos_mul */
    for (int j = c->rowPtr[Ac]; j < c->rowPtr[Ac + 1]; j++) {
      const int a = c->colPtr[j];
      for (int i = 0; i < dof; i++)
        for (int k = 0; k < dof; k++) Val[(size_t)d2 * j + i * dof + k] = Val[(size_t)d2 * j + i * dof + k] * W[k + dof * a];
    }
  for (int f = 0; f < c->nFaces; f++) {
    OFace* fa = &c->face[f];
    if (!fa->set || !fa->coupledFlag) continue;
    const int nd = fa->dof < dof ? fa->dof : dof;
    for (int a = 0; a < fa->nNo; a++)
      for (int i = 0; i < nd; i++) fa->valM[i + fa->dof * a] = fa->val[i + fa->dof * a] * W[i + dof * fa->glob[a]];
  }
}

static int gmres_v(OCase* c, int dof, const svb200_sublsparams* p, svb200_sublsresult* r, const double* Val, double* R,
                   svb200_lsresult* full)
{
  const int nNo = c->nNo, sD = p->sD;
  const size_t n = (size_t)dof * nNo;
  double* h = calloc((size_t)(sD + 1) * sD, sizeof(double));
  double* u = calloc(n * (sD + 1), sizeof(double));
  double* X = calloc(n, sizeof(double));
  double *y = calloc(sD, sizeof(double)), *cc = calloc(sD, sizeof(double)), *s = calloc(sD, sizeof(double)), *err = calloc(sD + 1, sizeof(double));
#define H(i, j) h[(i) + (size_t)(sD + 1) * (j)]
  int rc = 0;
  const double t0 = now_s();
  r->success = 0;
  double eps = normv((int)n, R);
  r->iNorm = eps; r->fNorm = eps;
  eps = fmax(p->absTol, p->relTol * eps);
  r->itr = 0;
  int last_i = 0;
  if (full) full->hist_n = 0;
  if (r->iNorm <= p->absTol) {
    r->callD = 2.220446049250313e-16; r->dB = 0.0; r->success = 1;
    goto done;
  }
  for (int l = 0; l < p->mItr; l++) {
    r->dB = r->fNorm;
    r->itr = r->itr + 1;
    spar_mul_vv(c, dof, Val, X, u);
    add_bc_mul_add(c, dof, X, u);
    for (size_t k = 0; k < n; k++) u[k] = R[k] - u[k];
    err[0] = normv((int)n, u);
    if (err[0] == 0.0) { rc = fail("FSILS: A zero matrix norm has been computed. This is probably caused by ill-posed boundary conditions."); goto done; }
    for (size_t k = 0; k < n; k++) u[k] = u[k] / err[0];
    for (int i = 0; i < sD; i++) {
      r->itr = r->itr + 1;
      last_i = i;
      double* ui = u + n * i; double* ui1 = u + n * (i + 1);
      spar_mul_vv(c, dof, Val, ui, ui1);
      add_bc_mul_add(c, dof, ui, ui1);
      for (int j = 0; j <= i + 1; j++) H(j, i) = nc_dot((int)n, u + n * j, ui1);
      for (int j = 0; j <= i; j++) {
        const double hj = H(j, i); const double* uj = u + n * j;
        for (size_t k = 0; k < n; k++) ui1[k] = ui1[k] + (-hj) * uj[k];
        H(i + 1, i) = H(i + 1, i) - hj * hj;
      }
      H(i + 1, i) = sqrt(fabs(H(i + 1, i)));
      { const double sc = 1.0 / H(i + 1, i); for (size_t k = 0; k < n; k++) ui1[k] = sc * ui1[k]; }
      for (int j = 0; j <= i - 1; j++) {
        const double tmp = cc[j] * H(j, i) + s[j] * H(j + 1, i);
        H(j + 1, i) = -s[j] * H(j, i) + cc[j] * H(j + 1, i);
        H(j, i) = tmp;
      }
      const double tmp = sqrt(H(i, i) * H(i, i) + H(i + 1, i) * H(i + 1, i));
      cc[i] = H(i, i) / tmp; s[i] = H(i + 1, i) / tmp;
      H(i, i) = tmp; H(i + 1, i) = 0.0;
      err[i + 1] = -s[i] * err[i];
      err[i] = cc[i] * err[i];
      if (full && full->hist && full->hist_n < full->hist_cap) full->hist[full->hist_n++] = fabs(err[i + 1]);
      if (fabs(err[i + 1]) < eps) { r->success = 1; break; }
    }
    if (last_i >= sD) last_i = sD - 1;
    for (int i = 0; i <= last_i; i++) y[i] = err[i];
    for (int j = last_i; j >= 0; j--) {
      for (int k = j + 1; k <= last_i; k++) y[j] = y[j] - H(j, k) * y[k];
      y[j] = y[j] / H(j, j);
    }
    for (int j = 0; j <= last_i; j++) { const double* uj = u + n * j; for (size_t k = 0; k < n; k++) X[k] = X[k] + y[j] * uj[k]; }
    r->fNorm = fabs(err[last_i + 1]);
    if (r->success) break;
  }
  memcpy(R, X, sizeof(double) * n);
  r->callD = now_s() - t0;
  r->dB = 10.0 * log(r->fNorm / r->dB);
done:
  free(h); free(u); free(X); free(y); free(cc); free(s); free(err);
  return rc;
#undef H
}

static int cgrad_v(OCase* c, int dof, const svb200_sublsparams* p, svb200_sublsresult* r, const double* K, double* R,
                   svb200_lsresult* full)
{
  if (full) full->hist_n = 0;
  const size_t n = (size_t)dof * c->nNo;
  double *P = malloc(sizeof(double) * n), *KP = malloc(sizeof(double) * n), *X = calloc(n, sizeof(double));
  const double t0 = now_s();
  r->success = 0;
  r->iNorm = normv((int)n, R);
  const double eps = pow(fmax(p->absTol, p->relTol * r->iNorm), 2.0);
  double errO = r->iNorm * r->iNorm, err = errO;
  memcpy(P, R, sizeof(double) * n);
  int last_i = 0;
  for (int i = 0; i < p->mItr; i++) {
    last_i = i;
    if (err < eps) { r->success = 1; break; }
    errO = err;
    spar_mul_vv(c, dof, K, P, KP);
    const double alpha = errO / nc_dot((int)n, P, KP);
    for (size_t k = 0; k < n; k++) X[k] = X[k] + alpha * P[k];
    for (size_t k = 0; k < n; k++) R[k] = R[k] + (-alpha) * KP[k];
    err = normv((int)n, R);
    err = err * err;
    if (full && full->hist && full->hist_n < full->hist_cap) full->hist[full->hist_n++] = sqrt(err);
    { const double q = errO / err; for (size_t k = 0; k < n; k++) P[k] = P[k] + q * R[k]; }
    { const double q = err / errO; for (size_t k = 0; k < n; k++) P[k] = q * P[k]; }
  }
  memcpy(R, X, sizeof(double) * n);
  r->itr = last_i;
  r->fNorm = sqrt(err);
  r->callD = now_s() - t0;
  r->dB = (errO < 2.220446049250313e-16) ? 0.0 : 5.0 * log(err / errO);
  free(P); free(KP); free(X);
  return 0;
}

static int bicgsv(OCase* c, int dof, const svb200_sublsparams* p, svb200_sublsresult* r, const double* K, double* R,
                  svb200_lsresult* full)
{
  if (full) full->hist_n = 0;
  const size_t n = (size_t)dof * c->nNo;
  double *P = malloc(sizeof(double) * n), *Rh = malloc(sizeof(double) * n), *X = calloc(n, sizeof(double)),
         *V = malloc(sizeof(double) * n), *S = malloc(sizeof(double) * n), *T = malloc(sizeof(double) * n);
  const double t0 = now_s();
  r->success = 0;
  double err = normv((int)n, R), errO = err;
  r->iNorm = err;
  const double eps = fmax(p->absTol, p->relTol * err);
  double rho = err * err;
  memcpy(P, R, sizeof(double) * n);
  memcpy(Rh, R, sizeof(double) * n);
  int i_itr = 1;
  for (int i = 0; i < p->mItr; i++) {
    if (err < eps) { r->success = 1; break; }
    spar_mul_vv(c, dof, K, P, V);
    const double alpha = rho / nc_dot((int)n, Rh, V);
    for (size_t k = 0; k < n; k++) S[k] = R[k] - alpha * V[k];
    spar_mul_vv(c, dof, K, S, T);
    double omega = normv((int)n, T);
    omega = nc_dot((int)n, T, S) / (omega * omega);
    for (size_t k = 0; k < n; k++) X[k] = X[k] + alpha * P[k] + omega * S[k];
    for (size_t k = 0; k < n; k++) R[k] = S[k] - omega * T[k];
    errO = err;
    err = normv((int)n, R);
    if (full && full->hist && full->hist_n < full->hist_cap) full->hist[full->hist_n++] = err;
    const double rhoO = rho;
    rho = nc_dot((int)n, R, Rh);
    const double beta = rho * alpha / (rhoO * omega);
    for (size_t k = 0; k < n; k++) P[k] = R[k] + beta * (P[k] - omega * V[k]);
    i_itr += 1;
  }
  memcpy(R, X, sizeof(double) * n);
  r->itr = i_itr - 1;
  r->fNorm = err;
  r->callD = now_s() - t0;
  r->dB = (errO < 2.220446049250313e-16) ? 0.0 : 10.0 * log(err / errO);
  free(P); free(Rh); free(X); free(V); free(S); free(T);
  return 0;
}

int svorc_solve(void* h, int dof, int ls_type, int prec, const svb200_lsparams* ls, int nFaces, const int* incL, const double* res,
                double* R_out, svb200_lsresult* out)
{
  OCase* c = h;
  (void)prec;
  if (dof != c->dof) return fail("[sv_oracle] dof mismatch");
  int anyNeu = 0;
  for (int f = 0; f < c->nFaces; f++) {
    c->face[f].incFlag = 1;
    if (incL && f < nFaces && incL[f] == 0) c->face[f].incFlag = 0;
    if (c->face[f].set && c->face[f].bGrp == SVB200_BC_NEU) anyNeu = 1;
  }
  if (anyNeu && !res) return fail("[fsils_solve] res is required for Neu surfaces");
  for (int f = 0; f < c->nFaces; f++) {
    OFace* fa = &c->face[f];
    fa->coupledFlag = 0;
    if (!fa->set || !fa->incFlag) continue;
    if (fa->bGrp == SVB200_BC_NEU && f < nFaces && res[f] != 0.0) { fa->res = res[f]; fa->coupledFlag = 1; }
  }
  const size_t n = (size_t)dof * c->nNo;
  double* W = malloc(sizeof(double) * n);
  svb200_lsresult local;
  memset(&local, 0, sizeof local);
  double* hist = out ? out->hist : NULL;
  int hist_cap = out ? out->hist_cap : 0;
  if (out) { memset(out, 0, sizeof *out); out->hist = hist; out->hist_cap = hist ? hist_cap : 0; }
  svb200_lsresult* o = out ? out : &local;
  const double t0 = now_s();
  precond_diag(c, dof, c->Val, c->R, W);
  int rc = 0;
  switch (ls_type) {
    case SVB200_LS_GMRES: rc = gmres_v(c, dof, &ls->RI, &o->RI, c->Val, c->R, o); break;
    case SVB200_LS_CG: rc = cgrad_v(c, dof, &ls->RI, &o->RI, c->Val, c->R, o); break;
    case SVB200_LS_BICGS: rc = bicgsv(c, dof, &ls->RI, &o->RI, c->Val, c->R, o); break;
    default: rc = fail("[sv_oracle] LS type not restated yet");
  }
  if (!rc) {
    for (size_t i = 0; i < n; i++) c->R[i] = W[i] * c->R[i];
    if (R_out) memcpy(R_out, c->R, sizeof(double) * n);
  }
  c->last_solve_s = now_s() - t0;
  free(W);
  return rc;
}

int svorc_spmv(void* h, int dof, const double* U, double* KU)
{
  OCase* c = h;
  spar_mul_vv(c, dof, c->Val, U, KU);
  return 0;
}

int svorc_last_timing(void* h, double* assemble_s, double* solve_s)
{
  OCase* c = h;
  if (assemble_s) *assemble_s = c->last_assemble_s;
  if (solve_s) *solve_s = c->last_solve_s;
  return 0;
}
