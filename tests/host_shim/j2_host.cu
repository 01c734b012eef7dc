// TEST INFRASTRUCTURE: the __host__ __device__ J2 point update of folax_b200/csrc/j2_point.cuh called on the CPU
// (float64, host pointers) -- see adjoint_host.cu.  Never linked into libfolax_b200, never used by the product.
#include <stdint.h>

#include "../../folax_b200/csrc/j2_point.cuh"

using namespace fol;

// n points: eps (n, V), state (n, V+1) -> sigma (n, V), tangent (n, V, V), state_new (n, V+1); mat = E, nu, y0, h1, h2
extern "C" int host_j2_points(int dim, long long n, const double* eps, const double* state, const double* mat,
                              double* sigma, double* tangent, double* state_new) {
  for (long long i = 0; i < n; ++i) {
    if (dim == 3)
      j2_point<double, 3>(eps + i * 6, state + i * 7, mat[0], mat[1], mat[2], mat[3], mat[4], sigma + i * 6,
                          tangent + i * 36, state_new + i * 7);
    else if (dim == 2)
      j2_point<double, 2>(eps + i * 3, state + i * 4, mat[0], mat[1], mat[2], mat[3], mat[4], sigma + i * 3,
                          tangent + i * 9, state_new + i * 4);
    else
      return -1;
  }
  return 0;
}
