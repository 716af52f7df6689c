// kabsch.cuh -- rotation of the weighted Procrustes problem (modules/registration/procrustes.py:6-73), shared by the
// local-to-global registration (matching.cu) and the RANSAC estimator (ransac.cu).
#pragma once
#include <cuda_runtime.h>

// 3x3 SVD by one-sided Jacobi in double precision; R = V diag(1,1,det(V U^T)) U^T.
__device__ inline void kabsch_rotation(const double H[3][3], double R[3][3]) {
  double A[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) A[i][j] = H[i][j];
  for (int sweep = 0; sweep < 30; sweep++) {
    double off = 0.0;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int i = 0; i < 3; i++) {
          alpha += A[i][p] * A[i][p];
          beta += A[i][q] * A[i][q];
          gamma += A[i][p] * A[i][q];
        }
        off = fmax(off, fabs(gamma) / sqrt(fmax(alpha * beta, 1e-300)));
        if (fabs(gamma) < 1e-300) continue;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int i = 0; i < 3; i++) {
          const double ap = A[i][p], aq = A[i][q];
          A[i][p] = c * ap - s * aq;
          A[i][q] = s * ap + c * aq;
          const double vp = V[i][p], vq = V[i][q];
          V[i][p] = c * vp - s * vq;
          V[i][q] = s * vp + c * vq;
        }
      }
    if (off < 1e-15) break;
  }
  // singular values = column norms; order descending so the det correction hits the smallest
  double sv[3];
  int ord[3] = {0, 1, 2};
  for (int j = 0; j < 3; j++) sv[j] = sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
  for (int a = 0; a < 2; a++)
    for (int b = a + 1; b < 3; b++)
      if (sv[ord[b]] > sv[ord[a]]) { const int t = ord[a]; ord[a] = ord[b]; ord[b] = t; }
  double U[3][3], Vs[3][3];
  for (int j = 0; j < 3; j++) {
    const int o = ord[j];
    for (int i = 0; i < 3; i++) {
      Vs[i][j] = V[i][o];
      U[i][j] = sv[o] > 1e-30 ? A[i][o] / sv[o] : 0.0;
    }
  }
  // Complete degenerate columns of U (singular value ~ 0) from the matching V column, orthogonalised
  // against the columns already fixed, so that V U^T tends to the identity on the null space: H = 0 (no
  // correspondences / all-zero weights) gives R = I like torch.svd of a zero matrix (U = V = I, procrustes.py:53).
  auto norm3 = [](double* x) { return sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]); };
  const double tiny = 1e-30;
  for (int j = 0; j < 3; j++) {
    if (sv[ord[j]] > tiny) continue;
    double best[3] = {0, 0, 0};
    double best_n = 0.0;
    for (int cand = 0; cand < 4 && best_n < 1e-6; cand++) {
      double w[3];
      for (int i = 0; i < 3; i++) w[i] = cand == 0 ? Vs[i][j] : (i == cand - 1 ? 1.0 : 0.0);
      for (int c = 0; c < j; c++) {
        const double dp = w[0] * U[0][c] + w[1] * U[1][c] + w[2] * U[2][c];
        for (int i = 0; i < 3; i++) w[i] -= dp * U[i][c];
      }
      const double n = norm3(w);
      if (n > best_n) { best_n = n; for (int i = 0; i < 3; i++) best[i] = w[i]; }
    }
    for (int i = 0; i < 3; i++) U[i][j] = best[i] / best_n;
  }
  // M = V U^T, d = sign(det M)
  double M[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) M[i][j] = Vs[i][0] * U[j][0] + Vs[i][1] * U[j][1] + Vs[i][2] * U[j][2];
  const double det = M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
                     M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
  const double d = det > 0 ? 1.0 : (det < 0 ? -1.0 : 0.0);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[i][j] = Vs[i][0] * U[j][0] + Vs[i][1] * U[j][1] + d * Vs[i][2] * U[j][2];
}

