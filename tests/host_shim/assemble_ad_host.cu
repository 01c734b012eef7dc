// TEST INFRASTRUCTURE: the __host__ __device__ thread body of folax_b200/csrc/assemble_ad_threads.cuh looped on the
// CPU (float64, host pointers) -- see adjoint_host.cu.  Never linked into libfolax_b200, never used by the product.
#include "../../folax_b200/csrc/assemble_ad_threads.cuh"

using namespace fol;

#define CASES(X)                                                                                   \
  X(HEX, 1) X(HEX, 2) X(HEX, 3) X(QUAD, 1) X(QUAD, 2) X(QUAD, 3) X(TET, 1) X(TET, 2) X(TET, 3)     \
  X(TRI, 1) X(TRI, 2) X(TRI, 3)

extern "C" int host_assemble_ad(int physics, int element, int num_gp, int transpose, long long ne, const double* xyz,
                                const int32_t* conn, const double* ctrl, const double* u, const uint8_t* dir,
                                const double* params, double* ke, double* re, double* energy) {
  AdAsmArgs<double> a{xyz, conn, ctrl, u, dir, ke, re, energy, ne, transpose, make_params<double>(params)};
#define X(E, O)                                                                                  \
  if (element == E && num_gp == O) {                                                             \
    for (long long e = 0; e < ne; ++e) {                                                         \
      if (physics == LAW_NEOHOOKE_AD) assemble_ad_thread<double, E, O, LAW_NEOHOOKE_AD>(e, a);   \
      else assemble_ad_thread<double, E, O, LAW_STVK_AD>(e, a);                                  \
    }                                                                                            \
    return 0;                                                                                    \
  }
  CASES(X)
#undef X
  return -3;
}
