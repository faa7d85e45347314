// struct_elem.cuh — Gauss-point algebra of the displacement-based solid (struct_3d) in closed Voigt form.
//
// Reference: struct_ns::struct_3d (Code/Source/solver/sv_struct.cpp:541-826) and
// mat_models::compute_pk2cc<3> (Code/Source/solver/mat_models.cpp:291-817) with bar_to_iso (:232-255),
// compute_svol_p (:1441-1464) and cc_to_voigt_eigen (:104-155).
//
// The reference builds the 3x3x3x3 elasticity tensor with Eigen tensor contractions
// (P : CC_bar : P^T costs 2 x 81 x 9 FMAs even when CC_bar = 0).  Here every model is written directly
// on the 6x6 Voigt matrix (order 11,22,33,12,23,31):
//   * volumetric      S += p J Ci ;  Dm += -2 p J (Ci (.) Ci) + pl J (Ci x Ci)
//   * isochoric       S_iso = J2d S_bar - r1 Ci,  r1 = J2d (C : S_bar)/3
//                     Dm += sum_k c_k (A~_k x A~_k)            with CC_bar = sum_k c_k A_k x A_k,
//                                                              A~ = A - (C:A)/3 Ci   (= P : A)
//                           - 2/3 (Ci x S_iso + S_iso x Ci) + 2 r1 (Ci (.) Ci) - 2 r1/3 (Ci x Ci)
//   nHK: CC_bar = 0; Guccione: 7 dyads; Mooney-Rivlin: I x I and the symmetric identity handled in closed form;
//   HGO: 2 dyads (dispersed structure tensors); Holzapfel-Ogden: 4 dyads (I, fibre-sheet, fibre, sheet) with the smoothed
//   Heaviside switch; HO-ma: the isotropic dyad through bar_to_iso, the three anisotropic ones added un-projected.
// (x = dyadic product, (.) = symmetric dyadic product 1/2(A_ik B_jl + A_il B_jk), mat_fun.h:201-223).
#pragma once
#include "fluid_elem.cuh"   // SVB_HD, is_zero

namespace svb {

struct StructDmn {
  double rho, f[3], dmp;
  double Kpen, C10, C01, bff, bss, bfs;
  double st_a, st_b, aff, ass, afs, kap, khs;   // stModelType a, b, aff, ass, afs, kap, khs (HGO / Holzapfel-Ogden)
  double visc_mu;
  int isoType, volType, Id, isStruct;
  int viscType, active;   // svb200_solid_visc; active = dmn.active_stress != nullptr (nodal Ya_f / Ya_s / Ya_n are used)
  int cann_off, cann_rows;   // rows [cann_off, cann_off + cann_rows) of the argument block's CANN table (SVB200_ISO_CANN)
};

// One row of ArtificialNeuralNetMaterial's parameter table (ArtificialNeuralNetMaterial.h: invariant_indices,
// activation_functions(:,0..2), weights(:,0..2)).
struct CannRow {
  int inv, a0, a1, a2;
  double w0, w1, w2;
};
constexpr int MAX_CANN_ROWS = 32;   // all domains of one assembly call together

#define SVB_VI(a) ((a) < 3 ? (a) : ((a) == 3 ? 0 : ((a) == 4 ? 1 : 2)))
#define SVB_VJ(a) ((a) < 3 ? (a) : ((a) == 3 ? 1 : ((a) == 4 ? 2 : 0)))

// Dm += c * (A x B + B x A)/2-free helpers on full 3x3 symmetric inputs.
SVB_HD void dm_add_dyad(double Dm[6][6], double c, const double A[3][3], const double B[3][3])
{
#pragma unroll
  for (int a = 0; a < 6; a++)
#pragma unroll
    for (int b = 0; b < 6; b++) Dm[a][b] += c * A[SVB_VI(a)][SVB_VJ(a)] * B[SVB_VI(b)][SVB_VJ(b)];
}

SVB_HD void dm_add_symdyad(double Dm[6][6], double c, const double A[3][3])
{
#pragma unroll
  for (int a = 0; a < 6; a++)
#pragma unroll
    for (int b = 0; b < 6; b++) {
      const int i = SVB_VI(a), j = SVB_VJ(a), k = SVB_VI(b), l = SVB_VJ(b);
      Dm[a][b] += c * 0.5 * (A[i][k] * A[j][l] + A[i][l] * A[j][k]);
    }
}

// Dm += c * symmetric_dyadic_product(A, B): 1/2 (A_ik B_jl + A_il B_jk)  (mat_fun.h:203-223)
SVB_HD void dm_add_symdyad2(double Dm[6][6], double c, const double A[3][3], const double B[3][3])
{
#pragma unroll
  for (int a = 0; a < 6; a++)
#pragma unroll
    for (int b = 0; b < 6; b++) {
      const int i = SVB_VI(a), j = SVB_VJ(a), k = SVB_VI(b), l = SVB_VJ(b);
      Dm[a][b] += c * 0.5 * (A[i][k] * B[j][l] + A[i][l] * B[j][k]);
    }
}

SVB_HD double ddot(const double A[3][3], const double B[3][3])
{
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) s += A[i][j] * B[i][j];
  return s;
}

// The CANN model (mat_models.cpp:776-800 with ArtificialNeuralNetMaterial.cpp:16-190) on the Voigt matrix; below.
SVB_HD void cann_voigt(const CannRow* rows, int nrows, int nFn, const double fN[2][3], const double C[3][3], const double Ci[3][3],
                       double J, double J2d, double J4d, double S[3][3], double Dm[6][6]);

// compute_pk2cc<3> (without prestress / viscosity, which struct_3d adds): F -> S (3x3), Dm (6x6).
// fN[0] = fibre, fN[1] = sheet direction.  ya = {Tfa, Tsa, Tna}: active stresses along the fibre, sheet and sheet-normal
// directions (mat_models.cpp:321-326; nullptr = none).  Returns 0, or 1 for an unsupported model.
// CANN = false compiles the CANN branch AND the active-stress terms out ("lean" twin: the closed-form TET4 kernels lost 8 % to the extra
// live values when the terms were compiled in unconditionally; 2.13 -> 2.30 ms on 4.4 M tets): inlined, its ~60 live doubles and 3x3 temporaries land in the stack frame of every solid
// kernel (HEX8 struct kernel 8 -> 568 bytes, TET4 216 -> 896), so the hot instantiations are built without it and a domain with a CANN
// model selects the CANN = true twin of the kernel.
template <bool CANN = true>
SVB_HD int pk2cc_voigt(const StructDmn& dm, const double F[3][3], const double fN[2][3], const double* ya, const CannRow* cann,
                       int nFn, double S[3][3], double Dm[6][6])
{
#pragma unroll
  for (int a = 0; a < 6; a++)
#pragma unroll
    for (int b = 0; b < 6; b++) Dm[a][b] = 0.0;
  const double J = F[0][0] * (F[1][1] * F[2][2] - F[1][2] * F[2][1]) - F[0][1] * (F[1][0] * F[2][2] - F[1][2] * F[2][0]) +
                   F[0][2] * (F[1][0] * F[2][1] - F[1][1] * F[2][0]);
  const double J2d = pow(J, -2.0 / 3.0);
  const double J4d = J2d * J2d;
  double C[3][3], Ci[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      C[i][j] = F[0][i] * F[0][j] + F[1][i] * F[1][j] + F[2][i] * F[2][j];
      S[i][j] = 0.0;
    }
  {
    const double dC = J * J;   // det C
    Ci[0][0] = (C[1][1] * C[2][2] - C[1][2] * C[2][1]) / dC;
    Ci[0][1] = (C[0][2] * C[2][1] - C[0][1] * C[2][2]) / dC;
    Ci[0][2] = (C[0][1] * C[1][2] - C[0][2] * C[1][1]) / dC;
    Ci[1][1] = (C[0][0] * C[2][2] - C[0][2] * C[2][0]) / dC;
    Ci[1][2] = (C[0][2] * C[1][0] - C[0][0] * C[1][2]) / dC;
    Ci[2][2] = (C[0][0] * C[1][1] - C[0][1] * C[1][0]) / dC;
    Ci[1][0] = Ci[0][1]; Ci[2][0] = Ci[0][2]; Ci[2][1] = Ci[1][2];
  }
  const double trC = C[0][0] + C[1][1] + C[2][2];

  // volumetric part (mat_models.cpp:397-405, 1441-1464)
  if (!is_zero(dm.Kpen)) {
    double p = 0.0, pl = 0.0;
    if (dm.volType == SVB200_VOL_QUAD) { p = dm.Kpen * (J - 1.0); pl = dm.Kpen * (2.0 * J - 1.0); }
    else if (dm.volType == SVB200_VOL_ST91) { p = 0.5 * dm.Kpen * (J - 1.0 / J); pl = dm.Kpen * J; }
    else if (dm.volType == SVB200_VOL_M94) { p = dm.Kpen * (1.0 - 1.0 / J); pl = dm.Kpen; }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) S[i][j] += p * J * Ci[i][j];
    dm_add_symdyad(Dm, -2.0 * p * J, Ci);
    dm_add_dyad(Dm, pl * J, Ci, Ci);
  }

  // Active stress Tfa f(x)f [+ Tsa s(x)s + Tna n(x)n for the Guccione / HO / HO-ma models; the reference throws for the
  // others when Tsa or Tna > 0, which the host checks at svb200_set_active_tension] — added to S_bar before the deviatoric
  // projection (mat_models.cpp:443, 461, 503, 570-575, 636, 650, 661-664) or, for HO-ma and CANN, to S directly (:745-772, 800).
  double Sact[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  const bool act = CANN && (ya != nullptr) && dm.active;
  if (act) {
    const bool dirs = (dm.isoType == SVB200_ISO_GUCCIONE || dm.isoType == SVB200_ISO_HO || dm.isoType == SVB200_ISO_HO_MA);
    double nrm[3] = {0, 0, 0};
    const bool useN = dirs && ya[2] > 0.0;
    if (useN) {
      nrm[0] = fN[0][1] * fN[1][2] - fN[0][2] * fN[1][1];
      nrm[1] = fN[0][2] * fN[1][0] - fN[0][0] * fN[1][2];
      nrm[2] = fN[0][0] * fN[1][1] - fN[0][1] * fN[1][0];
      const double nn = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
      nrm[0] /= nn; nrm[1] /= nn; nrm[2] /= nn;
    }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        double t = ya[0] * fN[0][i] * fN[0][j];
        if (dirs) t += ya[1] * fN[1][i] * fN[1][j];
        if (useN) t += ya[2] * nrm[i] * nrm[j];
        Sact[i][j] = t;
      }
  }

  double Idm[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  if (CANN && dm.isoType == SVB200_ISO_CANN) {
    cann_voigt(cann + dm.cann_off, dm.cann_rows, nFn, fN, C, Ci, J, J2d, J4d, S, Dm);
    if (act) {
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) S[i][j] += ya[0] * fN[0][i] * fN[0][j];
    }
    return 0;
  }
  if (dm.isoType == SVB200_ISO_STVK) {        // mat_models.cpp:415-421
    const double g1 = dm.C10, g2 = dm.C01 * 2.0;
    const double trE = 0.5 * (trC - 3.0);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) S[i][j] += g1 * trE * Idm[i][j] + g2 * 0.5 * (C[i][j] - Idm[i][j]);
    dm_add_dyad(Dm, g1, Idm, Idm);
    dm_add_symdyad(Dm, g2, Idm);
    return 0;
  }

  double Sb[3][3];
  if (dm.isoType == SVB200_ISO_NHK) {         // mat_models.cpp:435-450
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) Sb[i][j] = 2.0 * dm.C10 * Idm[i][j];
  } else if (dm.isoType == SVB200_ISO_MR) {   // mat_models.cpp:453-468
    const double Inv1 = J2d * trC;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) Sb[i][j] = 2.0 * (dm.C10 + Inv1 * dm.C01) * Idm[i][j] - 2.0 * dm.C01 * J2d * C[i][j];
    // CC_bar = c (I x I - Isym), c = 4 J4d C01:  P:(I x I):P^T = I~ x I~ ;
    // P:Isym:P^T = Isym - 1/3 (Ci x C + C x Ci) + (C:C)/9 Ci x Ci
    const double c = 4.0 * J4d * dm.C01;
    double It[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) It[i][j] = Idm[i][j] - (trC / 3.0) * Ci[i][j];
    dm_add_dyad(Dm, c, It, It);
    dm_add_symdyad(Dm, -c, Idm);
    dm_add_dyad(Dm, c / 3.0, Ci, C);
    dm_add_dyad(Dm, c / 3.0, C, Ci);
    dm_add_dyad(Dm, -c * ddot(C, C) / 9.0, Ci, Ci);
  } else if (dm.isoType == SVB200_ISO_GUCCIONE) {   // mat_models.cpp:513-581
    double R[3][3];   // R[k] = k-th local axis (fibre, sheet, normal)
#pragma unroll
    for (int i = 0; i < 3; i++) { R[0][i] = fN[0][i]; R[1][i] = fN[1][i]; }
    R[2][0] = R[0][1] * R[1][2] - R[0][2] * R[1][1];
    R[2][1] = R[0][2] * R[1][0] - R[0][0] * R[1][2];
    R[2][2] = R[0][0] * R[1][1] - R[0][1] * R[1][0];
    const double nn = sqrt(R[2][0] * R[2][0] + R[2][1] * R[2][1] + R[2][2] * R[2][2]);
#pragma unroll
    for (int i = 0; i < 3; i++) R[2][i] /= nn;
    double E[3][3], Es[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) E[i][j] = 0.5 * (J2d * C[i][j] - Idm[i][j]);
#pragma unroll
    for (int p = 0; p < 3; p++)
#pragma unroll
      for (int q = 0; q < 3; q++) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 3; j++) s += R[p][i] * E[i][j] * R[q][j];
        Es[p][q] = s;
      }
    const double g1 = dm.bff, g2 = dm.bss, g3 = dm.bfs;
    const double QQ = g1 * Es[0][0] * Es[0][0] +
                      g2 * (Es[1][1] * Es[1][1] + Es[2][2] * Es[2][2] + Es[1][2] * Es[1][2] + Es[2][1] * Es[2][1]) +
                      g3 * (Es[0][1] * Es[0][1] + Es[1][0] * Es[1][0] + Es[0][2] * Es[0][2] + Es[2][0] * Es[2][0]);
    const double r2 = dm.C10 * exp(QQ);
    // symmetrised dyads H[k]: 00, 11, 22, 12, 01, 20 with their weights in S_hat and CC_bar
    const int pa[6] = {0, 1, 2, 1, 0, 2}, pb[6] = {0, 1, 2, 2, 1, 0};
    const double ws[6] = {g1 * Es[0][0], g2 * Es[1][1], g2 * Es[2][2], 2.0 * g2 * Es[1][2], 2.0 * g3 * Es[0][1], 2.0 * g3 * Es[0][2]};
    const double wc[6] = {g1, g2, g2, 2.0 * g2, 2.0 * g3, 2.0 * g3};
    double Sh[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    const double cbar = r2 * J4d;
#pragma unroll
    for (int k = 0; k < 6; k++) {
      double H[3][3], Ht[3][3];
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) H[i][j] = 0.5 * (R[pa[k]][i] * R[pb[k]][j] + R[pb[k]][i] * R[pa[k]][j]);
      const double cH = ddot(C, H) / 3.0;
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) { Sh[i][j] += ws[k] * H[i][j]; Ht[i][j] = H[i][j] - cH * Ci[i][j]; }
      dm_add_dyad(Dm, cbar * wc[k], Ht, Ht);
    }
    {
      double St[3][3];
      const double cS = ddot(C, Sh) / 3.0;
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) { St[i][j] = Sh[i][j] - cS * Ci[i][j]; Sb[i][j] = r2 * Sh[i][j]; }
      dm_add_dyad(Dm, 2.0 * cbar, St, St);
    }
  } else if (dm.isoType == SVB200_ISO_HGO || dm.isoType == SVB200_ISO_HO || dm.isoType == SVB200_ISO_HO_MA) {
    // mat_models.cpp:469-511 (HGO), 582-687 (HO), 689-773 (HO-ma); active stresses Tfa/Tsa/Tna = 0 (no CEP coupling)
    double Hff[3][3], Hss[3][3], Hfs[3][3];
    double Cf1[3], Cf2[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      Cf1[i] = C[i][0] * fN[0][0] + C[i][1] * fN[0][1] + C[i][2] * fN[0][2];
      Cf2[i] = C[i][0] * fN[1][0] + C[i][1] * fN[1][1] + C[i][2] * fN[1][2];
#pragma unroll
      for (int j = 0; j < 3; j++) {
        Hff[i][j] = fN[0][i] * fN[0][j];
        Hss[i][j] = fN[1][i] * fN[1][j];
        Hfs[i][j] = 0.5 * (fN[0][i] * fN[1][j] + fN[1][i] * fN[0][j]);
      }
    }
    const double I4 = fN[0][0] * Cf1[0] + fN[0][1] * Cf1[1] + fN[0][2] * Cf1[2];
    const double I6 = fN[1][0] * Cf2[0] + fN[1][1] * Cf2[1] + fN[1][2] * Cf2[2];
    const double I8 = fN[0][0] * Cf2[0] + fN[0][1] * Cf2[1] + fN[0][2] * Cf2[2];
    const double Inv1 = J2d * trC;
    // S_bar = sum_k sw[k] A_k, CC_bar = sum_k cw[k] A_k x A_k (projected); un-projected extras for HO-ma
    double A[4][3][3], sw[4] = {0, 0, 0, 0}, cw[4] = {0, 0, 0, 0};
    int nA = 0;
    if (dm.isoType == SVB200_ISO_HGO) {
      const double kap = dm.kap;
      const double Eff = kap * Inv1 + (1.0 - 3.0 * kap) * J2d * I4 - 1.0;
      const double Ess = kap * Inv1 + (1.0 - 3.0 * kap) * J2d * I6 - 1.0;
      const double ef = exp(dm.bff * Eff * Eff), es = exp(dm.bss * Ess * Ess);
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
          A[0][i][j] = Idm[i][j];
          A[1][i][j] = kap * Idm[i][j] + (1.0 - 3.0 * kap) * Hff[i][j];
          A[2][i][j] = kap * Idm[i][j] + (1.0 - 3.0 * kap) * Hss[i][j];
        }
      nA = 3;
      sw[0] = 2.0 * dm.C10; sw[1] = 2.0 * dm.aff * Eff * ef; sw[2] = 2.0 * dm.ass * Ess * es;
      cw[1] = 4.0 * J4d * dm.aff * (1.0 + 2.0 * dm.bff * Eff * Eff) * ef;
      cw[2] = 4.0 * J4d * dm.ass * (1.0 + 2.0 * dm.bss * Ess * Ess) * es;
    } else {
      const bool ma = (dm.isoType == SVB200_ISO_HO_MA);
      const double jf = ma ? 1.0 : J2d;             // HO: isochoric invariants; HO-ma: full invariants
      const double Eff = jf * I4 - 1.0, Ess = jf * I6 - 1.0, Efs = jf * I8;
      const double k = dm.khs;
      const double of = 1.0 / (exp(k * Eff) + 1.0), os = 1.0 / (exp(k * Ess) + 1.0);
      const double c4f = 1.0 - of, c4s = 1.0 - os;
      const double dc4f = k * (of - of * of), dc4s = k * (os - os * os);
      const double ddc4f = k * k * (-of + 3.0 * of * of - 2.0 * of * of * of);
      const double ddc4s = k * k * (-os + 3.0 * os * os - 2.0 * os * os * os);
      const double giso = dm.st_a * exp(dm.st_b * (Inv1 - 3.0));
      const double efs = exp(dm.bfs * Efs * Efs), rf = exp(dm.bff * Eff * Eff), rs = exp(dm.bss * Ess * Ess);
      const double s_fs = 2.0 * dm.afs * Efs * efs;
      const double c_fs = 4.0 * dm.afs * (1.0 + 2.0 * dm.bfs * Efs * Efs) * efs;
      const double s_ff = 2.0 * dm.aff * (c4f * Eff * rf + (0.5 * dc4f / dm.bff) * (rf - 1.0));
      const double c_ff = 4.0 * dm.aff * ((c4f * (1.0 + 2.0 * dm.bff * Eff * Eff) + 2.0 * dc4f * Eff) * rf + (0.5 * ddc4f / dm.bff) * (rf - 1.0));
      const double s_ss = 2.0 * dm.ass * (c4s * Ess * rs + (0.5 * dc4s / dm.bss) * (rs - 1.0));
      const double c_ss = 4.0 * dm.ass * ((c4s * (1.0 + 2.0 * dm.bss * Ess * Ess) + 2.0 * dc4s * Ess) * rs + (0.5 * ddc4s / dm.bss) * (rs - 1.0));
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) { A[0][i][j] = Idm[i][j]; A[1][i][j] = Hfs[i][j]; A[2][i][j] = Hff[i][j]; A[3][i][j] = Hss[i][j]; }
      sw[0] = giso; cw[0] = 2.0 * J4d * dm.st_b * giso;
      if (!ma) {
        nA = 4;
        sw[1] = s_fs; cw[1] = J4d * c_fs;
        sw[2] = s_ff; cw[2] = J4d * c_ff;
        sw[3] = s_ss; cw[3] = J4d * c_ss;
      } else {
        nA = 1;
        // anisotropic terms enter S and CC directly, without the deviatoric projection (mat_models.cpp:737-767)
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 3; j++) S[i][j] += s_fs * Hfs[i][j] + s_ff * Hff[i][j] + s_ss * Hss[i][j];
        dm_add_dyad(Dm, c_fs, Hfs, Hfs);
        dm_add_dyad(Dm, c_ff, Hff, Hff);
        dm_add_dyad(Dm, c_ss, Hss, Hss);
        if (act) {
#pragma unroll
          for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) { S[i][j] += Sact[i][j]; Sact[i][j] = 0.0; }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) Sb[i][j] = 0.0;
    for (int k = 0; k < nA; k++) {
      double At[3][3];
      const double cA = ddot(C, A[k]) / 3.0;
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) { Sb[i][j] += sw[k] * A[k][i][j]; At[i][j] = A[k][i][j] - cA * Ci[i][j]; }
      if (cw[k] != 0.0) dm_add_dyad(Dm, cw[k], At, At);
    }
  } else {
    return 1;
  }
  if (act) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) Sb[i][j] += Sact[i][j];
  }
  // bar_to_iso (mat_models.cpp:232-255)
  const double r1 = J2d * ddot(C, Sb) / 3.0;
  double Siso[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      Siso[i][j] = J2d * Sb[i][j] - r1 * Ci[i][j];
      S[i][j] += Siso[i][j];
    }
  dm_add_dyad(Dm, -2.0 / 3.0, Ci, Siso);
  dm_add_dyad(Dm, -2.0 / 3.0, Siso, Ci);
  dm_add_symdyad(Dm, 2.0 * r1, Ci);
  dm_add_dyad(Dm, -2.0 * r1 / 3.0, Ci, Ci);
  return 0;
}

// ---- CANN: constitutive artificial neural network (Peirlinck et al. 2025; ArtificialNeuralNetMaterial.cpp) ------------------
// psi = sum_rows W2 f2(f1(f0(I_k - ref_k))), S = 2 sum_k dpsi_k dI_k/dC, CC = 4 sum_k (dpsi_k d2I_k/dC2 + ddpsi_k dI_k (x) dI_k).
// The invariants and their first / second derivatives follow computeInvariantsAndDerivatives (:114-190) term by term; the
// fourth-order tensors are accumulated directly on the Voigt matrix (cc_to_voigt_eigen picks CC(i,j,k,l) with (i,j), (k,l) the
// Voigt pairs 00,11,22,01,12,20 of the upper triangle and mirrors it).
SVB_HD void cann_act(const CannRow& r, double x, double& d1, double& d2)
{
  // uCANN_h0 / h1 / h2 (:16-70) and the chain rule of uCANN (:73-88); returns d psi_row / dI and d2 psi_row / dI2 (without W2)
  double f0 = x, df0 = 1.0, ddf0 = 0.0;
  if (r.a0 == 2) { f0 = 0.5 * (fabs(x) + x); df0 = (x == 0.0) ? 0.0 : 0.5 * (fabs(x) / x + 1.0); }
  else if (r.a0 == 3) { f0 = fabs(x); df0 = fabs(x) / x; }
  double f1 = r.w0 * f0, df1 = r.w0, ddf1 = 0.0;
  if (r.a1 == 2) { f1 = r.w0 * r.w0 * f0 * f0; df1 = 2.0 * r.w0 * r.w0 * f0; ddf1 = 2.0 * r.w0 * r.w0; }
  double df2 = r.w1, ddf2 = 0.0;
  if (r.a2 == 2) { const double ex = exp(r.w1 * f1); df2 = r.w1 * ex; ddf2 = r.w1 * r.w1 * ex; }
  else if (r.a2 == 3) { const double q = 1.0 - r.w1 * f1; df2 = r.w1 / q; ddf2 = -r.w1 * r.w1 / (q * q); }
  d1 = df2 * df1 * df0;
  d2 = (ddf2 * df1 * df1 + df2 * ddf1) * df0 * df0 + df2 * df1 * ddf0;
}

// Invariant pair (J2d n.C m, J4d n.C^2 m) of a structure tensor N = sym(n (x) m): first derivatives dA = -IA/3 Ci + J2d N,
// dB = J4d (N C + C N) - IB/3 Ci and their second derivatives, weighted by (pA, ppA) and (pB, ppB) = (dpsi, ddpsi).
SVB_HD void cann_fibre_pair(const double N[3][3], double IA, double IB, double pA, double ppA, double pB, double ppB,
                            const double C[3][3], const double Ci[3][3], const double Idm[3][3], double J2d, double J4d,
                            double S[3][3], double Dm[6][6])
{
  double dA[3][3], dB[3][3], NC[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < 3; k++) t += N[i][k] * C[k][j] + C[i][k] * N[k][j];
      NC[i][j] = t;
    }
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      dA[i][j] = -IA / 3.0 * Ci[i][j] + J2d * N[i][j];
      dB[i][j] = J4d * NC[i][j] - IB / 3.0 * Ci[i][j];
      S[i][j] += 2.0 * (pA * dA[i][j] + pB * dB[i][j]);
    }
  if (pA != 0.0) {
    // ddA = -1/3 (dA (x) Ci + J2d Ci (x) N + IA dCidC), dCidC = -sym(Ci, Ci)
    const double c = 4.0 * pA * (-1.0 / 3.0);
    dm_add_dyad(Dm, c, dA, Ci);
    dm_add_dyad(Dm, c * J2d, Ci, N);
    dm_add_symdyad(Dm, -c * IA, Ci);
  }
  if (ppA != 0.0) dm_add_dyad(Dm, 4.0 * ppA, dA, dA);
  if (pB != 0.0) {
    // ddB = -1/3 (dB (x) Ci + IB dCidC + 2 J4d Ci (x) (N C + C N)) + J4d (2 sym(N, I) - N (x) I + 2 sym(I, N) - I (x) N)
    const double c = 4.0 * pB * (-1.0 / 3.0);
    dm_add_dyad(Dm, c, dB, Ci);
    dm_add_symdyad(Dm, -c * IB, Ci);
    dm_add_dyad(Dm, c * 2.0 * J4d, Ci, NC);
    const double d = 4.0 * pB * J4d;
    dm_add_symdyad2(Dm, 2.0 * d, N, Idm);
    dm_add_dyad(Dm, -d, N, Idm);
    dm_add_symdyad2(Dm, 2.0 * d, Idm, N);
    dm_add_dyad(Dm, -d, Idm, N);
  }
  if (ppB != 0.0) dm_add_dyad(Dm, 4.0 * ppB, dB, dB);
}

SVB_HD void cann_voigt(const CannRow* rows, int nrows, int nFn, const double fN[2][3], const double C[3][3], const double Ci[3][3],
                       double J, double J2d, double J4d, double S[3][3], double Dm[6][6])
{
  const double Idm[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  double C2[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C2[i][j] = C[i][0] * C[0][j] + C[i][1] * C[1][j] + C[i][2] * C[2][j];
  const double trC = C[0][0] + C[1][1] + C[2][2], trC2 = C2[0][0] + C2[1][1] + C2[2][2];
  double N1[3][3], N2[3][3], N12[3][3];
  double Inv[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  Inv[0] = J2d * trC;
  Inv[1] = 0.5 * (Inv[0] * Inv[0] - J4d * trC2);
  Inv[2] = J * J;   // det C
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      N1[i][j] = fN[0][i] * fN[0][j];
      N2[i][j] = fN[1][i] * fN[1][j];
      N12[i][j] = 0.5 * (fN[0][i] * fN[1][j] + fN[1][i] * fN[0][j]);
    }
  auto quad = [](const double a[3], const double M[3][3], const double b[3]) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) s += a[i] * M[i][j] * b[j];
    return s;
  };
  Inv[3] = J2d * quad(fN[0], C, fN[0]);
  Inv[4] = J4d * quad(fN[0], C2, fN[0]);
  if (nFn == 2) {
    Inv[5] = J2d * quad(fN[0], C, fN[1]);
    Inv[6] = J4d * quad(fN[0], C2, fN[1]);
    Inv[7] = J2d * quad(fN[1], C, fN[1]);
    Inv[8] = J4d * quad(fN[1], C2, fN[1]);
  }
  // evaluate (:90-112): dpsi / ddpsi per invariant
  const double ref[9] = {3, 3, 1, 1, 1, 0, 0, 1, 1};
  double dpsi[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, ddpsi[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int r = 0; r < nrows; r++) {
    const int k = rows[r].inv - 1;
    double d1, d2;
    cann_act(rows[r], Inv[k] - ref[k], d1, d2);
    dpsi[k] += rows[r].w2 * d1;
    ddpsi[k] += rows[r].w2 * d2;
  }
  // invariants 1-3
  double d1[3][3], d2[3][3], d3[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      d1[i][j] = -Inv[0] / 3.0 * Ci[i][j] + J2d * Idm[i][j];
      d3[i][j] = Inv[2] * Ci[i][j];
    }
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      d2[i][j] = (trC2 / 3.0) * Ci[i][j] + Inv[0] * d1[i][j] + J4d * C[i][j];
      S[i][j] += 2.0 * (dpsi[0] * d1[i][j] + dpsi[1] * d2[i][j] + dpsi[2] * d3[i][j]);
    }
  // ddInv1 = -1/3 (d1 (x) Ci - Inv0 sym(Ci,Ci) + J2d Ci (x) I); it also enters ddInv2 with the factor Inv0
  {
    const double c = (4.0 * dpsi[0] + 4.0 * dpsi[1] * Inv[0]) * (-1.0 / 3.0);
    if (c != 0.0) {
      dm_add_dyad(Dm, c, d1, Ci);
      dm_add_symdyad(Dm, -c * Inv[0], Ci);
      dm_add_dyad(Dm, c * J2d, Ci, Idm);
    }
  }
  if (ddpsi[0] != 0.0) dm_add_dyad(Dm, 4.0 * ddpsi[0], d1, d1);
  if (dpsi[1] != 0.0) {
    // ddInv2 - Inv0 ddInv1 = d1 (x) d1 - trC2/3 sym(Ci,Ci) + 1/3 (trC2 dJ4ddC + 2 J4d C) (x) Ci + dJ4ddC (x) C - J4d Isym,
    // dJ4ddC = -2/3 J4d Ci
    const double c = 4.0 * dpsi[1];
    double T[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) T[i][j] = trC2 * (-2.0 / 3.0 * J4d) * Ci[i][j] + 2.0 * J4d * C[i][j];
    dm_add_dyad(Dm, c, d1, d1);
    dm_add_symdyad(Dm, -c * trC2 / 3.0, Ci);
    dm_add_dyad(Dm, c / 3.0, T, Ci);
    dm_add_dyad(Dm, c * (-2.0 / 3.0 * J4d), Ci, C);
    dm_add_symdyad(Dm, -c * J4d, Idm);
  }
  if (ddpsi[1] != 0.0) dm_add_dyad(Dm, 4.0 * ddpsi[1], d2, d2);
  if (dpsi[2] != 0.0) {
    // ddInv3 = d3 (x) Ci - Inv2 sym(Ci,Ci)
    dm_add_dyad(Dm, 4.0 * dpsi[2], d3, Ci);
    dm_add_symdyad(Dm, -4.0 * dpsi[2] * Inv[2], Ci);
  }
  if (ddpsi[2] != 0.0) dm_add_dyad(Dm, 4.0 * ddpsi[2], d3, d3);
  // invariants 4-5 (fibre), 6-7 (fibre-sheet), 8-9 (sheet).  Without a second fibre family the reference leaves dInv6..9 = 0.
  if (dpsi[3] != 0.0 || ddpsi[3] != 0.0 || dpsi[4] != 0.0 || ddpsi[4] != 0.0)
    cann_fibre_pair(N1, Inv[3], Inv[4], dpsi[3], ddpsi[3], dpsi[4], ddpsi[4], C, Ci, Idm, J2d, J4d, S, Dm);
  if (nFn == 2) {
    if (dpsi[5] != 0.0 || ddpsi[5] != 0.0 || dpsi[6] != 0.0 || ddpsi[6] != 0.0)
      cann_fibre_pair(N12, Inv[5], Inv[6], dpsi[5], ddpsi[5], dpsi[6], ddpsi[6], C, Ci, Idm, J2d, J4d, S, Dm);
    if (dpsi[7] != 0.0 || ddpsi[7] != 0.0 || dpsi[8] != 0.0 || ddpsi[8] != 0.0)
      cann_fibre_pair(N2, Inv[7], Inv[8], dpsi[7], ddpsi[7], dpsi[8], ddpsi[8], C, Ci, Idm, J2d, J4d, S, Dm);
  }
  // cc_to_voigt_eigen keeps the upper triangle and mirrors it
#pragma unroll
  for (int a = 1; a < 6; a++)
#pragma unroll
    for (int b = 0; b < a; b++) Dm[a][b] = Dm[b][a];
}

// nn::gnn for insd = 3 and any eNoN (Code/Source/solver/nn.cpp:862-899): Nxi[a][k] -> Nx[a][i], Jac.
template <int ENON>
SVB_HD double gnn3(const double Nxi[][3], const double xl[][3], double Nx[][3])
{
  double xXi[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, xiX[3][3];
#pragma unroll
  for (int a = 0; a < ENON; a++)
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int k = 0; k < 3; k++) xXi[i][k] += xl[a][i] * Nxi[a][k];
  const double Jac = xXi[0][0] * xXi[1][1] * xXi[2][2] + xXi[0][1] * xXi[1][2] * xXi[2][0] + xXi[0][2] * xXi[1][0] * xXi[2][1] -
                     xXi[0][0] * xXi[1][2] * xXi[2][1] - xXi[0][1] * xXi[1][0] * xXi[2][2] - xXi[0][2] * xXi[1][1] * xXi[2][0];
  const double iJ = 1.0 / Jac;
  xiX[0][0] = (xXi[1][1] * xXi[2][2] - xXi[1][2] * xXi[2][1]) * iJ;
  xiX[0][1] = (xXi[2][1] * xXi[0][2] - xXi[2][2] * xXi[0][1]) * iJ;
  xiX[0][2] = (xXi[0][1] * xXi[1][2] - xXi[0][2] * xXi[1][1]) * iJ;
  xiX[1][0] = (xXi[1][2] * xXi[2][0] - xXi[1][0] * xXi[2][2]) * iJ;
  xiX[1][1] = (xXi[2][2] * xXi[0][0] - xXi[2][0] * xXi[0][2]) * iJ;
  xiX[1][2] = (xXi[0][2] * xXi[1][0] - xXi[0][0] * xXi[1][2]) * iJ;
  xiX[2][0] = (xXi[1][0] * xXi[2][1] - xXi[1][1] * xXi[2][0]) * iJ;
  xiX[2][1] = (xXi[2][0] * xXi[0][1] - xXi[2][1] * xXi[0][0]) * iJ;
  xiX[2][2] = (xXi[0][0] * xXi[1][1] - xXi[0][1] * xXi[1][0]) * iJ;
#pragma unroll
  for (int a = 0; a < ENON; a++)
#pragma unroll
    for (int i = 0; i < 3; i++) Nx[a][i] = Nxi[a][0] * xiX[0][i] + Nxi[a][1] * xiX[1][i] + Nxi[a][2] * xiX[2][i];
  return Jac;
}

// Bm(:,i) for node a (sv_struct.cpp:705-729): Bm[r][i], r = Voigt row, i = displacement component.
SVB_HD void make_Bm(const double Nx[3], const double F[3][3], double Bm[6][3])
{
#pragma unroll
  for (int i = 0; i < 3; i++) {
    Bm[0][i] = Nx[0] * F[i][0];
    Bm[1][i] = Nx[1] * F[i][1];
    Bm[2][i] = Nx[2] * F[i][2];
    Bm[3][i] = Nx[0] * F[i][1] + F[i][0] * Nx[1];
    Bm[4][i] = Nx[1] * F[i][2] + F[i][1] * Nx[2];
    Bm[5][i] = Nx[2] * F[i][0] + F[i][2] * Nx[0];
  }
}

SVB_HD void make_DBm(const double Dm[6][6], const double Bm[6][3], double DBm[6][3])
{
#pragma unroll
  for (int r = 0; r < 6; r++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < 6; q++) s += Dm[r][q] * Bm[q][j];
      DBm[r][j] = s;
    }
}

// K(i,j) += w [ delta_ij (amd Na Nb + afu gradNa.S.gradNb) + afu Bm_a(:,i).DBm_b(:,j) ]  (sv_struct.cpp:736-825)
SVB_HD void struct_block(double K[3][3], double w, double amdNaNb, double afu, const double SNxa[3], const double Nxb[3],
                         const double Bma[6][3], const double DBmb[6][3])
{
  const double NxSNx = SNxa[0] * Nxb[0] + SNxa[1] * Nxb[1] + SNxa[2] * Nxb[2];
  const double T1 = amdNaNb + afu * NxSNx;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
#pragma unroll
      for (int r = 0; r < 6; r++) s += Bma[r][i] * DBmb[r][j];
      K[i][j] += w * ((i == j ? T1 : 0.0) + afu * s);
    }
}


// ---- linear tetrahedra in closed form (assemble_struct_tet4_kernel, assemble_mesh_tet4_kernel) -------------------------
// Moments of the quadrature rule: W = sum_g w_g, m1_a = sum_g w_g N_a(g), M2_ab = sum_g w_g N_a(g) N_b(g); `scale` is the
// element Jacobian (struct, lElas) or 1 (mesh equation: Jacobian-free weight, mesh.cpp:122).  N is indexed [g][a] with row
// stride ldN.
struct Tet4Mom {
  double W, m1[4], M2[4][4];
};

SVB_HD void tet4_moments(const double* w, const double* N, int ldN, double scale, Tet4Mom& q)
{
  q.W = 0.0;
#pragma unroll
  for (int a = 0; a < 4; a++) {
    q.m1[a] = 0.0;
#pragma unroll
    for (int b = 0; b < 4; b++) q.M2[a][b] = 0.0;
  }
#pragma unroll
  for (int g = 0; g < 4; g++) {
    const double wg = w[g] * scale;
    q.W += wg;
#pragma unroll
    for (int a = 0; a < 4; a++) {
      q.m1[a] += wg * N[g * ldN + a];
#pragma unroll
      for (int b = 0; b < 4; b++) q.M2[a][b] += wg * N[g * ldN + a] * N[g * ldN + b];
    }
  }
}

// struct_3d residual of node a: lR(i) = -rho f_i m1_a + sum_b M2_ab q_b(i) + W (F S grad N_a)_i, q_b = rho (a_b - bf_b) + dmp v_b.
SVB_HD void struct_tet4_residual(const StructDmn& dm, const Tet4Mom& q, int a, const double Nxa[3], const double Pk[3][3],
                                 const double ql[4][3], double lR[3])
{
#pragma unroll
  for (int i = 0; i < 3; i++) {
    double r = -dm.rho * dm.f[i] * q.m1[a] + q.W * (Pk[i][0] * Nxa[0] + Pk[i][1] * Nxa[1] + Pk[i][2] * Nxa[2]);
#pragma unroll
    for (int b = 0; b < 4; b++) r += q.M2[a][b] * ql[b][i];
    lR[i] = r;
  }
}

// struct_3d block (a,b): K(i,j) = delta_ij (amd M2_ab + afu W grad N_a . S grad N_b) + afu W Bm_a(:,i) . DBm_b(:,j).
SVB_HD void struct_tet4_block(double wafu, double amdMab, const double Nxa[3], const double SNb[3], const double Bma[6][3],
                              const double DBmb[6][3], double K[3][3])
{
  const double T1 = amdMab + wafu * (Nxa[0] * SNb[0] + Nxa[1] * SNb[1] + Nxa[2] * SNb[2]);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double t = 0.0;
#pragma unroll
      for (int r = 0; r < 6; r++) t += Bma[r][i] * DBmb[r][j];
      K[i][j] = wafu * t + (i == j ? T1 : 0.0);
    }
}

// l_elas_3d (mesh / lElas equation) on a linear tet: Voigt stress of the constant strain, residual of node a, block (a,b).
SVB_HD void lelas_tet4_stress(double lambda, double mu, const double Nx[4][3], const double dl[4][3], double S[6])
{
  double ed[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int b = 0; b < 4; b++) {
    ed[0] += Nx[b][0] * dl[b][0];
    ed[1] += Nx[b][1] * dl[b][1];
    ed[2] += Nx[b][2] * dl[b][2];
    ed[3] += Nx[b][1] * dl[b][0] + Nx[b][0] * dl[b][1];
    ed[4] += Nx[b][2] * dl[b][1] + Nx[b][1] * dl[b][2];
    ed[5] += Nx[b][0] * dl[b][2] + Nx[b][2] * dl[b][0];
  }
  const double divD = lambda * (ed[0] + ed[1] + ed[2]);
  S[0] = divD + 2.0 * mu * ed[0]; S[1] = divD + 2.0 * mu * ed[1]; S[2] = divD + 2.0 * mu * ed[2];
  S[3] = mu * ed[3]; S[4] = mu * ed[4]; S[5] = mu * ed[5];
}

SVB_HD void lelas_tet4_residual(double rho, const double f[3], const Tet4Mom& q, int a, const double Nxa[3], const double S[6],
                                const double ql[4][3], double lR[3])
{
  double r[3];
  r[0] = q.W * (Nxa[0] * S[0] + Nxa[1] * S[3] + Nxa[2] * S[5]);
  r[1] = q.W * (Nxa[0] * S[3] + Nxa[1] * S[1] + Nxa[2] * S[4]);
  r[2] = q.W * (Nxa[0] * S[5] + Nxa[1] * S[4] + Nxa[2] * S[2]);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    double t = -f[i] * q.m1[a];
#pragma unroll
    for (int b = 0; b < 4; b++) t += q.M2[a][b] * ql[b][i];
    lR[i] = r[i] + rho * t;
  }
}

// K(i,j) = T1c [ delta_ij (amd M2_ab + mu W grad N_a . grad N_b) + mu W (lDm Nx_a(i) Nx_b(j) + Nx_a(j) Nx_b(i)) ]; c0 = T1c amd,
// c1 = T1c mu W.
SVB_HD void lelas_tet4_block(double c0Mab, double c1, double lDm, const double Nxa[3], const double Nxb[3], double K[3][3])
{
  const double T1 = c0Mab + c1 * (Nxa[0] * Nxb[0] + Nxa[1] * Nxb[1] + Nxa[2] * Nxb[2]);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) K[i][j] = c1 * (lDm * Nxa[i] * Nxb[j] + Nxa[j] * Nxb[i]) + (i == j ? T1 : 0.0);
}

// ---- solid viscosity (mat_models.cpp:1583-1762) ----------------------------------------------------------
// Both models of compute_visc_stress_and_tangent have the same structure: three vectors per element node,
//     V1_a = T grad N_a,  V2_a = A V1_a,  V3_a = B V1_a,
// one 3x3 matrix M and one scalar c per Gauss point, and a tangent block that is bilinear in (V_a, V_b):
//   Potential  (:1583-1637)  T = I, A = F, B = afu vx + afv F, M = afu F vx^T + afv F F^T, c = mu/2
//       afu Kvis_u + afv Kvis_v = c [ V2_b(i) V3_a(j) + (V1_a.V1_b) M(i,j) ]
//   Newtonian  (:1660-1733)  T = F^-T, A = dev sym(vx F^-1), B = M = vx F^-1, c = mu J
//       Kvis_u = c [ 2 (V2_a(i) V1_b(j) - V2_b(i) V1_a(j)) - ((V1_a.V1_b) M(i,j) + V1_b(i) V3_a(j) - 2/3 V1_a(i) V3_b(j)) ]
//       Kvis_v = c [ (V1_a.V1_b) delta_ij + V1_b(i) V1_a(j) - 2/3 V1_a(i) V1_b(j) ]
// (the reference's own comment notes that the Newtonian tangent is probably not the exact derivative; it is
// reproduced as written).
struct ViscGP {
  double T[3][3], A[3][3], B[3][3], M[3][3];
  double c;
};

// Svis (added to S before the prestress / P = F S, sv_struct.cpp:666-669) and the Gauss-point terms above.
SVB_HD void visc_gauss_point(int viscType, double mu, double afu, double afv, const double F[3][3], const double vx[3][3],
                             double Svis[3][3], ViscGP& gp)
{
  if (viscType == SVB200_SOLID_VISC_POTENTIAL) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const double Ftvx_ij = F[0][i] * vx[0][j] + F[1][i] * vx[1][j] + F[2][i] * vx[2][j];
        const double Ftvx_ji = F[0][j] * vx[0][i] + F[1][j] * vx[1][i] + F[2][j] * vx[2][i];
        Svis[i][j] = mu * 0.5 * (Ftvx_ij + Ftvx_ji);
        gp.T[i][j] = (i == j) ? 1.0 : 0.0;
        gp.A[i][j] = F[i][j];
        gp.B[i][j] = afu * vx[i][j] + afv * F[i][j];
        const double FFt = F[i][0] * F[j][0] + F[i][1] * F[j][1] + F[i][2] * F[j][2];
        const double Fvxt = F[i][0] * vx[j][0] + F[i][1] * vx[j][1] + F[i][2] * vx[j][2];
        gp.M[i][j] = afu * Fvxt + afv * FFt;
      }
    gp.c = 0.5 * mu;
    return;
  }
  // Newtonian
  const double J = F[0][0] * (F[1][1] * F[2][2] - F[1][2] * F[2][1]) - F[0][1] * (F[1][0] * F[2][2] - F[1][2] * F[2][0]) +
                   F[0][2] * (F[1][0] * F[2][1] - F[1][1] * F[2][0]);
  double Fi[3][3];
  Fi[0][0] = (F[1][1] * F[2][2] - F[1][2] * F[2][1]) / J;
  Fi[0][1] = (F[0][2] * F[2][1] - F[0][1] * F[2][2]) / J;
  Fi[0][2] = (F[0][1] * F[1][2] - F[0][2] * F[1][1]) / J;
  Fi[1][0] = (F[1][2] * F[2][0] - F[1][0] * F[2][2]) / J;
  Fi[1][1] = (F[0][0] * F[2][2] - F[0][2] * F[2][0]) / J;
  Fi[1][2] = (F[0][2] * F[1][0] - F[0][0] * F[1][2]) / J;
  Fi[2][0] = (F[1][0] * F[2][1] - F[1][1] * F[2][0]) / J;
  Fi[2][1] = (F[0][1] * F[2][0] - F[0][0] * F[2][1]) / J;
  Fi[2][2] = (F[0][0] * F[1][1] - F[0][1] * F[1][0]) / J;
  double vF[3][3], dd[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) vF[i][j] = vx[i][0] * Fi[0][j] + vx[i][1] * Fi[1][j] + vx[i][2] * Fi[2][j];
  const double tr3 = (vF[0][0] + vF[1][1] + vF[2][2]) / 3.0;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) dd[i][j] = 0.5 * (vF[i][j] + vF[j][i]) - (i == j ? tr3 : 0.0);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 3; k++)
#pragma unroll
        for (int l = 0; l < 3; l++) s += Fi[i][k] * dd[k][l] * Fi[j][l];
      Svis[i][j] = 2.0 * mu * J * s;
      gp.T[i][j] = Fi[j][i];
      gp.A[i][j] = dd[i][j];
      gp.B[i][j] = vF[i][j];
      gp.M[i][j] = vF[i][j];
    }
  gp.c = mu * J;
}

// V[0..2] = V1, V[3..5] = V2, V[6..8] = V3 of one node.
SVB_HD void visc_node(const ViscGP& gp, const double Nx[3], double V[9])
{
#pragma unroll
  for (int i = 0; i < 3; i++) V[i] = gp.T[i][0] * Nx[0] + gp.T[i][1] * Nx[1] + gp.T[i][2] * Nx[2];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    V[3 + i] = gp.A[i][0] * V[0] + gp.A[i][1] * V[1] + gp.A[i][2] * V[2];
    V[6 + i] = gp.B[i][0] * V[0] + gp.B[i][1] * V[1] + gp.B[i][2] * V[2];
  }
}

// K(i,j) += wc (afu Kvis_u(i,j;a,b) + afv Kvis_v(i,j;a,b)) / c, wc = Gauss weight * gp.c  (sv_struct.cpp:759-823)
SVB_HD void visc_block(int viscType, double wc, double afu, double afv, const double M[9], const double Va[9], const double Vb[9],
                       double K[3][3])
{
  const double dot = Va[0] * Vb[0] + Va[1] * Vb[1] + Va[2] * Vb[2];
  if (viscType == SVB200_SOLID_VISC_POTENTIAL) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) K[i][j] += wc * (Vb[3 + i] * Va[6 + j] + dot * M[3 * i + j]);
    return;
  }
  const double r2d = 2.0 / 3.0;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const double ku = 2.0 * (Va[3 + i] * Vb[j] - Vb[3 + i] * Va[j]) - (dot * M[3 * i + j] + Vb[i] * Va[6 + j] - r2d * Va[i] * Vb[6 + j]);
      const double kv = (i == j ? dot : 0.0) + Vb[i] * Va[j] - r2d * Va[i] * Vb[j];
      K[i][j] += wc * (afu * ku + afv * kv);
    }
}

}  // namespace svb
