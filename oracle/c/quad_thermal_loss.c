/* Oracle (TEST INFRASTRUCTURE / CPU baseline only): plain-C restatement of the batched FOL physics loss and its
 * gradient for the thermal Quad4 loss, 2x2 Gauss rule, float64 -- the arithmetic of
 *   fol/geometries/quadrilateral_2d_4.py:23-30, 54-70, fol/geometries/geometry.py:88-97,
 *   fol/loss_functions/thermal.py:28-49 (dense Se per element as written there, energy T^T (Se sg(T) - Fe)),
 *   fol/loss_functions/fe_loss.py:166-176, 250-262 (sum over elements, mean over the batch)
 * and of the JAX-AD gradient under the reference's stop_gradient placement:
 *   dE/dT = assembled residual Se T,   dE/dK_a = sum_g N_a (1 + beta T_g^c) |grad T_g|^2 detJ w.
 * Samples are independent: the OpenMP loop runs over samples, so the scatter needs no atomics.
 * U must already carry the Dirichlet values (fe_loss.py:255).  Built into oracle/_build/liboracle_hex.so. */
#include <math.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <string.h>

static const double QX[4] = {-1, 1, 1, -1};
static const double QY[4] = {-1, -1, 1, 1};

void oracle_quad_thermal_loss_grads(int64_t nb, int64_t ne, int64_t nn, const double* coords /* (nn,3) */,
                                    const int32_t* conn /* (ne,4) */, const double* K /* (nb,nn) */,
                                    const double* U /* (nb,nn) */, double beta, double cexp,
                                    double* energy /* (nb) */, double* gU /* (nb,nn) */, double* gK /* (nb,nn) */) {
  const double p = 1.0 / sqrt(3.0);
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t b = 0; b < nb; ++b) {
    const double* Kb = K + b * nn;
    const double* Ub = U + b * nn;
    double* gu = gU + b * nn;
    double* gk = gK + b * nn;
    memset(gu, 0, sizeof(double) * nn);
    memset(gk, 0, sizeof(double) * nn);
    double Eb = 0.0;
    for (int64_t e = 0; e < ne; ++e) {
      const int32_t* c = conn + e * 4;
      double X[4][2], de[4], T[4];
      for (int a = 0; a < 4; ++a) {
        X[a][0] = coords[3 * (int64_t)c[a]];
        X[a][1] = coords[3 * (int64_t)c[a] + 1];
        de[a] = Kb[c[a]];
        T[a] = Ub[c[a]];
      }
      double Se[4][4], dK[4];
      memset(Se, 0, sizeof Se);
      memset(dK, 0, sizeof dK);
      for (int g = 0; g < 4; ++g) {
        const double xi = QX[g] * p, eta = QY[g] * p;
        double N[4], dN[4][2], J[2][2] = {{0, 0}, {0, 0}};
        for (int a = 0; a < 4; ++a) {
          const double fx = 1 + QX[a] * xi, fy = 1 + QY[a] * eta;
          N[a] = 0.25 * fx * fy;
          dN[a][0] = 0.25 * QX[a] * fy;
          dN[a][1] = 0.25 * QY[a] * fx;
        }
        for (int a = 0; a < 4; ++a)
          for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) J[i][j] += X[a][i] * dN[a][j];
        const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
        const double inv[2][2] = {{J[1][1] / det, -J[0][1] / det}, {-J[1][0] / det, J[0][0] / det}};
        double gN[4][2], kg = 0.0, tg = 0.0, gT[2] = {0, 0};
        for (int a = 0; a < 4; ++a) {
          for (int k = 0; k < 2; ++k) gN[a][k] = dN[a][0] * inv[0][k] + dN[a][1] * inv[1][k];
          kg += N[a] * de[a];
          tg += N[a] * T[a];
        }
        for (int a = 0; a < 4; ++a)
          for (int k = 0; k < 2; ++k) gT[k] += gN[a][k] * T[a];
        const double nl = 1.0 + beta * pow(tg, cexp);
        const double wd = det; /* weights are 1 */
        const double kappa = kg * nl;
        for (int a = 0; a < 4; ++a) {
          for (int bb = 0; bb < 4; ++bb) Se[a][bb] += kappa * (gN[a][0] * gN[bb][0] + gN[a][1] * gN[bb][1]) * wd;
          dK[a] += N[a] * nl * (gT[0] * gT[0] + gT[1] * gT[1]) * wd;
        }
      }
      for (int a = 0; a < 4; ++a) {
        double re = 0.0;
        for (int bb = 0; bb < 4; ++bb) re += Se[a][bb] * T[bb];
        Eb += T[a] * re;
        gu[c[a]] += re;
        gk[c[a]] += dK[a];
      }
    }
    energy[b] = Eb;
  }
}
