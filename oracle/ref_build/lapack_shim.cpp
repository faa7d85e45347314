// Minimal LAPACK stand-ins (dgesv_, dgetrf_, dgetri_) for the oracle build of the
// reference: column-major LU with partial pivoting, same pivot rule as LAPACK
// (largest |a| in the column, first occurrence).  Test infrastructure only.
// Call sites in the reference: solver/nn.cpp:1277 (6x6 second-derivative solve,
// RHS == 0 for tet4), solver/mat_fun.cpp (dense inverse, cold).
#include <cmath>
#include <vector>
#include <utility>

static int lu_factor(int n, double* A, int lda, int* ipiv)
{
  int info = 0;
  for (int k = 0; k < n; k++) {
    int p = k; double amax = std::fabs(A[k + k*lda]);
    for (int i = k+1; i < n; i++) {
      double v = std::fabs(A[i + k*lda]);
      if (v > amax) { amax = v; p = i; }
    }
    ipiv[k] = p + 1;
    if (A[p + k*lda] == 0.0) { if (!info) info = k+1; continue; }
    if (p != k) for (int j = 0; j < n; j++) std::swap(A[k + j*lda], A[p + j*lda]);
    double inv = 1.0 / A[k + k*lda];
    for (int i = k+1; i < n; i++) A[i + k*lda] *= inv;
    for (int j = k+1; j < n; j++) {
      double akj = A[k + j*lda];
      for (int i = k+1; i < n; i++) A[i + j*lda] -= A[i + k*lda]*akj;
    }
  }
  return info;
}

static void lu_solve(int n, const double* A, int lda, const int* ipiv, double* b)
{
  for (int k = 0; k < n; k++) { int p = ipiv[k]-1; if (p != k) std::swap(b[k], b[p]); }
  for (int k = 0; k < n; k++) for (int i = k+1; i < n; i++) b[i] -= A[i + k*lda]*b[k];
  for (int k = n-1; k >= 0; k--) { b[k] /= A[k + k*lda]; for (int i = 0; i < k; i++) b[i] -= A[i + k*lda]*b[k]; }
}

extern "C" {
int dgetrf_(int* m, int* n, double* A, int* lda, int* ipiv, int* info)
{
  (void)m; *info = lu_factor(*n, A, *lda, ipiv); return 0;
}
int dgetri_(int* n, double* A, int* lda, int* ipiv, double* work, int* lwork, int* info)
{
  (void)work; (void)lwork;
  int N = *n, L = *lda;
  std::vector<double> inv(N*N, 0.0), col(N);
  for (int j = 0; j < N; j++) {
    for (int i = 0; i < N; i++) col[i] = (i == j) ? 1.0 : 0.0;
    lu_solve(N, A, L, ipiv, col.data());
    for (int i = 0; i < N; i++) inv[i + j*N] = col[i];
  }
  for (int j = 0; j < N; j++) for (int i = 0; i < N; i++) A[i + j*L] = inv[i + j*N];
  *info = 0; return 0;
}
int dgesv_(const int* n, const int* nrhs, double* A, const int* lda, int* ipiv, double* B, const int* ldb, int* info)
{
  *info = lu_factor(*n, A, *lda, ipiv);
  if (*info) return 0;
  for (int r = 0; r < *nrhs; r++) lu_solve(*n, A, *lda, ipiv, B + r*(*ldb));
  return 0;
}
}
