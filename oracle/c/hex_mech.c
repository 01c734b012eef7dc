/* Oracle (TEST INFRASTRUCTURE / CPU baseline only): plain-C restatement of the reference's residual +
 * Jacobian assembly for Hex8 small-strain elasticity, 2x2x2 Gauss rule, float64 -- the arithmetic of
 *   fol/geometries/hexahedra_3d_8.py:23-33, 79-114, fol/geometries/geometry.py:88-97,
 *   fol/loss_functions/mechanical.py:46-58, 63-82, 98-117 (dense B^T (D B) per Gauss point, as written there),
 *   fol/loss_functions/fe_loss.py:191-230 (row mask), :299-306 (data + scatter-add).
 * Elements are processed in parallel with OpenMP; the residual scatter uses `omp atomic`, like the
 * reference's own host FFI does (ffi_functions/kr_small_displacement_element.cc:81-93).
 * Built by oracle/c/Makefile into oracle/_build/liboracle_hex.so; never linked by the product. */
#include <math.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <string.h>

static const double SX[8] = {-1, 1, 1, -1, -1, 1, 1, -1};
static const double SY[8] = {-1, -1, 1, 1, -1, -1, 1, 1};
static const double SZ[8] = {-1, -1, -1, -1, 1, 1, 1, 1};

static void element(const double X[8][3], const double de[8], const double u[24], double E, double nu,
                    const double body[3], double Ke[24][24], double re[24]) {
  const double p = 1.0 / sqrt(3.0);
  const double c1 = E / ((1.0 + nu) * (1.0 - 2.0 * nu));
  const double c2 = c1 * (1.0 - nu), c3 = c1 * nu, c4 = c1 * 0.5 * (1.0 - 2.0 * nu);
  double D[6][6];
  memset(D, 0, sizeof D);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) D[i][j] = (i == j) ? c2 : c3;
  D[3][3] = D[4][4] = D[5][5] = c4;
  double Fe[24];
  memset(Ke, 0, sizeof(double) * 576);
  memset(Fe, 0, sizeof Fe);
  for (int g = 0; g < 8; ++g) {
    const double xi = SX[g] * p, eta = SY[g] * p, zeta = SZ[g] * p;
    double N[8], dN[8][3], J[3][3] = {{0}};
    for (int a = 0; a < 8; ++a) {
      const double fx = 1 + SX[a] * xi, fy = 1 + SY[a] * eta, fz = 1 + SZ[a] * zeta;
      N[a] = 0.125 * fx * fy * fz;
      dN[a][0] = 0.125 * SX[a] * fy * fz;
      dN[a][1] = 0.125 * SY[a] * fx * fz;
      dN[a][2] = 0.125 * SZ[a] * fx * fy;
    }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        for (int a = 0; a < 8; ++a) J[i][j] += X[a][i] * dN[a][j];
    const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02, r = 1.0 / det;
    double inv[3][3];
    inv[0][0] = c00 * r; inv[1][0] = c01 * r; inv[2][0] = c02 * r;
    inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * r;
    inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * r;
    inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * r;
    inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * r;
    inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * r;
    inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * r;
    double B[6][24], DB[6][24], eg = 0.0;
    memset(B, 0, sizeof B);
    for (int a = 0; a < 8; ++a) {
      double gx[3];
      for (int k = 0; k < 3; ++k) gx[k] = dN[a][0] * inv[0][k] + dN[a][1] * inv[1][k] + dN[a][2] * inv[2][k];
      B[0][3 * a] = gx[0]; B[1][3 * a + 1] = gx[1]; B[2][3 * a + 2] = gx[2];
      B[3][3 * a] = gx[1]; B[3][3 * a + 1] = gx[0];
      B[4][3 * a + 1] = gx[2]; B[4][3 * a + 2] = gx[1];
      B[5][3 * a] = gx[2]; B[5][3 * a + 2] = gx[0];
      eg += N[a] * de[a];
    }
    for (int s = 0; s < 6; ++s)
      for (int n = 0; n < 24; ++n) {
        double acc = 0.0;
        for (int t = 0; t < 6; ++t) acc += D[s][t] * B[t][n];
        DB[s][n] = acc;
      }
    const double w = det * eg; /* Gauss weight 1 */
    for (int m = 0; m < 24; ++m)
      for (int n = 0; n < 24; ++n) {
        double acc = 0.0;
        for (int s = 0; s < 6; ++s) acc += B[s][m] * DB[s][n];
        Ke[m][n] += w * acc;
      }
    for (int a = 0; a < 8; ++a)
      for (int k = 0; k < 3; ++k) Fe[3 * a + k] += det * N[a] * body[k];
  }
  for (int m = 0; m < 24; ++m) {
    double acc = 0.0;
    for (int n = 0; n < 24; ++n) acc += Ke[m][n] * u[n];
    re[m] = acc - Fe[m];
  }
}

/* data (ne*576), residual (3*nn, zeroed here); dir_flag[ndof] = 1 on Dirichlet dofs */
/* thread count of the parallel loops (the launcher may have exported OMP_NUM_THREADS=1, e.g. torchrun) */
void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_hex_mech_assemble(int64_t ne, int64_t nn, const double* xyz, const int32_t* conn, const double* ctrl,
                              const double* uvec, const uint8_t* dir_flag, double E, double nu, const double* body,
                              int transpose, double* data, double* residual) {
  memset(residual, 0, sizeof(double) * 3 * (size_t)nn);
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < ne; ++e) {
    double X[8][3], de[8], u[24], bc[24], Ke[24][24], re[24];
    int32_t gd[24];
    for (int a = 0; a < 8; ++a) {
      const int32_t n = conn[e * 8 + a];
      de[a] = ctrl[n];
      for (int k = 0; k < 3; ++k) {
        X[a][k] = xyz[3 * (int64_t)n + k];
        gd[3 * a + k] = 3 * n + k;
        u[3 * a + k] = uvec[3 * (int64_t)n + k];
        bc[3 * a + k] = dir_flag[3 * (int64_t)n + k] ? 0.0 : 1.0;
      }
    }
    element(X, de, u, E, nu, body, Ke, re);
    double* out = data + (size_t)e * 576;
    for (int m = 0; m < 24; ++m)
      for (int n = 0; n < 24; ++n) {
        const double v = transpose ? Ke[n][m] : Ke[m][n];
        out[m * 24 + n] = (m == n) ? v : bc[m] * v;
      }
    for (int m = 0; m < 24; ++m) {
      const double v = bc[m] * re[m];
#pragma omp atomic
      residual[gd[m]] += v;
    }
  }
}
