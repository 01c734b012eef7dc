// Element library for the sm_100a kernels: Gauss rules, shape functions, local gradients.
// Same tables as the reference element classes (cited per block); evaluated in registers.
// (__host__ __device__ so that tests/host_shim can run the sequential per-element routines of adjoint.cuh on
//  the CPU; the library itself never calls them on the host.)
//   fol/geometries/hexahedra_3d_8.py:17-114, quadrilateral_2d_4.py:17-70,
//   tetrahedra_3d_4.py:17-60, triangle_2d_3.py:17-58, geometry.py:88-97
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fol {

enum : int { HEX = 0, QUAD = 1, TET = 2, TRI = 3 };

// 1/sqrt(3) and sqrt(3/5) rounded to double exactly as `1.0/np.sqrt(3.0)`, `np.sqrt(0.6)`
#define FOL_S3 0.57735026918962584
#define FOL_S35 0.7745966692414834

__host__ __device__ constexpr int elem_nnode(int e) { return e == HEX ? 8 : (e == TRI ? 3 : 4); }
__host__ __device__ constexpr int elem_dim(int e) { return (e == HEX || e == TET) ? 3 : 2; }
__host__ __device__ constexpr int elem_ngauss(int e, int order) {
  return e == HEX ? (order == 1 ? 1 : order == 2 ? 8 : 27)
       : e == QUAD ? (order == 1 ? 1 : order == 2 ? 4 : 9)
       : e == TET ? (order == 1 ? 1 : order == 2 ? 4 : 8)
                  : (order == 1 ? 1 : order == 2 ? 3 : 4);
}

// sign patterns of the tensor-product elements (node order of hexahedra_3d_8.py:79-88 and
// quadrilateral_2d_4.py:54-58)
__host__ __device__ __forceinline__ double sgn_x(int a) { return ((a & 3) == 1 || (a & 3) == 2) ? 1.0 : -1.0; }
__host__ __device__ __forceinline__ double sgn_y(int a) { return (a & 2) ? 1.0 : -1.0; }
__host__ __device__ __forceinline__ double sgn_z(int a) { return (a & 4) ? 1.0 : -1.0; }

// Gauss point g of integration order ORDER: xi[3] and weight.
template <int ELEM, int ORDER>
__host__ __device__ __forceinline__ void gauss_point(int g, double xi[3], double& w) {
  if constexpr (ELEM == HEX) {
    if constexpr (ORDER == 1) { xi[0] = xi[1] = xi[2] = 0.0; w = 8.0; }
    else if constexpr (ORDER == 2) {  // ordered like the nodes, hexahedra_3d_8.py:23-33
      xi[0] = sgn_x(g) * FOL_S3; xi[1] = sgn_y(g) * FOL_S3; xi[2] = sgn_z(g) * FOL_S3; w = 1.0;
    } else {  // x fastest, hexahedra_3d_8.py:37-76
      const int i = g % 3, j = (g / 3) % 3, k = g / 9;
      const double p[3] = {-FOL_S35, 0.0, FOL_S35};
      const double q[3] = {5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0};
      xi[0] = p[i]; xi[1] = p[j]; xi[2] = p[k]; w = q[i] * q[j] * q[k];
    }
  } else if constexpr (ELEM == QUAD) {
    xi[2] = 0.0;
    if constexpr (ORDER == 1) { xi[0] = xi[1] = 0.0; w = 4.0; }
    else if constexpr (ORDER == 2) { xi[0] = sgn_x(g) * FOL_S3; xi[1] = sgn_y(g) * FOL_S3; w = 1.0; }
    else {
      const int i = g % 3, j = g / 3;
      const double p[3] = {-FOL_S35, 0.0, FOL_S35};
      const double q[3] = {5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0};
      xi[0] = p[i]; xi[1] = p[j]; w = q[i] * q[j];
    }
  } else if constexpr (ELEM == TET) {
    if constexpr (ORDER == 1) { xi[0] = xi[1] = xi[2] = 0.25; w = 1.0 / 6.0; }
    else if constexpr (ORDER == 2) {  // 8-digit literals of tetrahedra_3d_4.py:24-27
      const double a = 0.58541020, b = 0.13819660;
      xi[0] = (g == 0) ? a : b; xi[1] = (g == 1) ? a : b; xi[2] = (g == 2) ? a : b; w = 1.0 / 24.0;
    } else {  // tetrahedra_3d_4.py:31-44
      const bool hi = g >= 4; const int h = g & 3;
      const double a = hi ? 0.67914317820120795168 : 0.015835909865720057993;
      const double b = hi ? 0.10695227393293068277 : 0.32805469671142664734;
      xi[0] = (h == 0) ? a : b; xi[1] = (h == 1) ? a : b; xi[2] = (h == 2) ? a : b;
      w = hi ? 0.01857867224802297628 : 0.02308799441864369039;
    }
  } else {  // TRI, triangle_2d_3.py:17-41
    xi[2] = 0.0;
    if constexpr (ORDER == 1) { xi[0] = xi[1] = 1.0 / 3.0; w = 0.5; }
    else if constexpr (ORDER == 2) {
      xi[0] = (g == 1) ? 2.0 / 3.0 : 1.0 / 6.0; xi[1] = (g == 2) ? 2.0 / 3.0 : 1.0 / 6.0; w = 1.0 / 6.0;
    } else {
      xi[0] = (g == 1) ? 0.6 : (g == 3 ? 1.0 / 3.0 : 0.2);
      xi[1] = (g == 2) ? 0.6 : (g == 3 ? 1.0 / 3.0 : 0.2);
      w = (g == 3) ? -27.0 / 96.0 : 25.0 / 96.0;
    }
  }
}

// Shape-function values N[a] and local gradients dN[a][dim] at xi.
template <int ELEM, class T>
__host__ __device__ __forceinline__ void shape_functions(const double xi[3], T* N, T (*dN)[elem_dim(ELEM)]) {
  if constexpr (ELEM == HEX) {
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const double fx = 1.0 + sgn_x(a) * xi[0], fy = 1.0 + sgn_y(a) * xi[1], fz = 1.0 + sgn_z(a) * xi[2];
      N[a] = (T)(0.125 * fx * fy * fz);
      dN[a][0] = (T)(0.125 * sgn_x(a) * fy * fz);
      dN[a][1] = (T)(0.125 * sgn_y(a) * fx * fz);
      dN[a][2] = (T)(0.125 * sgn_z(a) * fx * fy);
    }
  } else if constexpr (ELEM == QUAD) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const double fx = 1.0 + sgn_x(a) * xi[0], fy = 1.0 + sgn_y(a) * xi[1];
      N[a] = (T)(0.25 * fx * fy);
      dN[a][0] = (T)(0.25 * sgn_x(a) * fy);
      dN[a][1] = (T)(0.25 * sgn_y(a) * fx);
    }
  } else if constexpr (ELEM == TET) {
    N[0] = (T)(1.0 - (xi[0] + xi[1] + xi[2])); N[1] = (T)xi[0]; N[2] = (T)xi[1]; N[3] = (T)xi[2];
    dN[0][0] = dN[0][1] = dN[0][2] = (T)-1.0;
    dN[1][0] = (T)1.0; dN[1][1] = (T)0.0; dN[1][2] = (T)0.0;
    dN[2][0] = (T)0.0; dN[2][1] = (T)1.0; dN[2][2] = (T)0.0;
    dN[3][0] = (T)0.0; dN[3][1] = (T)0.0; dN[3][2] = (T)1.0;
  } else {
    N[0] = (T)(1.0 - xi[0] - xi[1]); N[1] = (T)xi[0]; N[2] = (T)xi[1];
    dN[0][0] = dN[0][1] = (T)-1.0;
    dN[1][0] = (T)1.0; dN[1][1] = (T)0.0;
    dN[2][0] = (T)0.0; dN[2][1] = (T)1.0;
  }
}

// J = (dN^T X)^T, its determinant and grad N = dN . J^-1 (geometry.py:88-97).
// X is a*3 (2-D elements use the first two columns, quadrilateral_2d_4.py:68-70).
// TRANSPOSED = true reproduces `B_mat = invJ @ dN^T` of transient_thermal.py:57-58 / phase_field.py:47-48,
// where the inverse Jacobian enters un-transposed: gN[a][k] = sum_j dN[a][j] inv[k][j].
template <int ELEM, class T, bool TRANSPOSED = false>
__host__ __device__ __forceinline__ T global_gradients(const T* X /* [A][3] */, const T (*dN)[elem_dim(ELEM)],
                                              T (*gN)[elem_dim(ELEM)]) {
  constexpr int A = elem_nnode(ELEM), D = elem_dim(ELEM);
  T J[D][D];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      T acc = (T)0;
#pragma unroll
      for (int a = 0; a < A; ++a) acc += X[a * 3 + i] * dN[a][j];
      J[i][j] = acc;
    }
  T inv[D][D];
  T det;
  if constexpr (D == 2) {
    det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const T r = (T)1 / det;
    inv[0][0] = J[1][1] * r; inv[0][1] = -J[0][1] * r;
    inv[1][0] = -J[1][0] * r; inv[1][1] = J[0][0] * r;
  } else {
    const T c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    const T c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    const T c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    const T r = (T)1 / det;
    inv[0][0] = c00 * r; inv[1][0] = c01 * r; inv[2][0] = c02 * r;
    inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * r;
    inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * r;
    inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * r;
    inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * r;
    inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * r;
    inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * r;
  }
#pragma unroll
  for (int a = 0; a < A; ++a)
#pragma unroll
    for (int k = 0; k < D; ++k) {
      T acc = (T)0;
#pragma unroll
      for (int j = 0; j < D; ++j) acc += dN[a][j] * (TRANSPOSED ? inv[k][j] : inv[j][k]);
      gN[a][k] = acc;
    }
  return det;
}

}  // namespace fol
