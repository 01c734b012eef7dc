// TEST INFRASTRUCTURE: runs the __host__ __device__ per-thread bodies of folax_b200/csrc/adjoint_threads.cuh
// on the CPU (a plain loop over the elements instead of a kernel launch), so the arithmetic AND the indexing of
// the adjoint-sensitivity kernels can be checked against the oracle in a container without a GPU.
// Built by tests/test_adjoint_host_shim.py with `nvcc -shared` (host code only is called); float64, host pointers.
// Never linked into libfolax_b200 and never used by the product.
#include "../../folax_b200/csrc/adjoint_threads.cuh"

namespace fol {
thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};
}

using namespace fol;

#define CASES(X)                                                                                   \
  X(HEX, 1) X(HEX, 2) X(HEX, 3) X(QUAD, 1) X(QUAD, 2) X(QUAD, 3) X(TET, 1) X(TET, 2) X(TET, 3)     \
  X(TRI, 1) X(TRI, 2) X(TRI, 3)

extern "C" {

int host_gauss_interpolate(int element, int num_gp, int dpn, long long ne, const int32_t* conn, const double* ctrl,
                           const double* u, double* kg, double* ug) {
  InterpArgs<double> a{conn, ctrl, u, kg, ug, ne, dpn};
#define X(E, O)                                                                         \
  if (element == E && num_gp == O) {                                                    \
    for (long long e = 0; e < ne; ++e) gauss_interpolate_thread<double, E, O>(e, a);    \
    return 0;                                                                           \
  }
  CASES(X)
#undef X
  return -3;
}

int host_response_elements(int element, int num_gp, int dpn, long long ne, const double* xyz, const int32_t* conn,
                           const double* f, const double* fk, const double* fu, double* val, double* du, double* dk,
                           double* dx) {
  ResponseArgs<double> a{xyz, conn, f, fk, fu, val, du, dk, dx, ne, dpn};
#define X(E, O)                                                                \
  if (element == E && num_gp == O) {                                           \
    for (long long e = 0; e < ne; ++e) response_thread<double, E, O>(e, a);    \
    return 0;                                                                  \
  }
  CASES(X)
#undef X
  return -3;
}

int host_residual_adjoint_elements(int physics, int element, int num_gp, int accumulate, long long ne,
                                   const double* xyz, const int32_t* conn, const double* ctrl, const double* u,
                                   const double* lam, const double* aux, const double* params, double* dk,
                                   double* dx) {
  AdjointArgs<double> a{xyz, conn, ctrl, u, lam, aux, dk, dx, ne, accumulate, make_params<double>(params)};
#define X(E, O)                                                                                   \
  if (element == E && num_gp == O) {                                                              \
    for (long long e = 0; e < ne; ++e) {                                                          \
      switch (physics) {                                                                          \
        case 0: residual_adjoint_thread<double, E, O, ADJ_MECH>(e, a); break;                     \
        case 1: residual_adjoint_thread<double, E, O, ADJ_THERMAL>(e, a); break;                  \
        case 2: residual_adjoint_thread<double, E, O, ADJ_NEOHOOKE>(e, a); break;                 \
        case 4: residual_adjoint_thread<double, E, O, ADJ_STVK>(e, a); break;                     \
        case 5: residual_adjoint_thread<double, E, O, ADJ_TTHERMAL>(e, a); break;                 \
        case 6: residual_adjoint_thread<double, E, O, ADJ_ALLENCAHN>(e, a); break;                \
        default: return -3;                                                                       \
      }                                                                                           \
    }                                                                                             \
    return 0;                                                                                     \
  }
  CASES(X)
#undef X
  return -3;
}

int host_element_energies(int physics, int element, int num_gp, long long ne, const double* xyz, const int32_t* conn,
                          const double* ctrl, const double* u, const double* aux, const double* params,
                          double* energy) {
  AdjointArgs<double> a{xyz, conn, ctrl, u, nullptr, aux, energy, nullptr, ne, 0, make_params<double>(params)};
#define X(E, O)                                                                                   \
  if (element == E && num_gp == O) {                                                              \
    for (long long e = 0; e < ne; ++e) {                                                          \
      switch (physics) {                                                                          \
        case 0: element_energy_thread<double, E, O, ADJ_MECH>(e, a); break;                       \
        case 1: element_energy_thread<double, E, O, ADJ_THERMAL>(e, a); break;                    \
        case 2: element_energy_thread<double, E, O, ADJ_NEOHOOKE>(e, a); break;                   \
        case 4: element_energy_thread<double, E, O, ADJ_STVK>(e, a); break;                       \
        case 5: element_energy_thread<double, E, O, ADJ_TTHERMAL>(e, a); break;                   \
        case 6: element_energy_thread<double, E, O, ADJ_ALLENCAHN>(e, a); break;                  \
        default: return -3;                                                                       \
      }                                                                                           \
    }                                                                                             \
    return 0;                                                                                     \
  }
  CASES(X)
#undef X
  return -3;
}

// whole-element forward-mode sweeps for every physics (cross-check of the product routes)
int host_residual_adjoint_dual_reference(int physics, int element, int num_gp, long long ne, const double* xyz,
                                         const int32_t* conn, const double* ctrl, const double* u, const double* lam,
                                         const double* aux, const double* params, double* dk, double* dx) {
  AdjointArgs<double> a{xyz, conn, ctrl, u, lam, aux, dk, dx, ne, 0, make_params<double>(params)};
#define X(E, O)                                                                                        \
  if (element == E && num_gp == O) {                                                                   \
    for (long long e = 0; e < ne; ++e) {                                                               \
      switch (physics) {                                                                               \
        case 0: residual_adjoint_dual_reference_thread<double, E, O, ADJ_MECH>(e, a); break;           \
        case 1: residual_adjoint_dual_reference_thread<double, E, O, ADJ_THERMAL>(e, a); break;        \
        case 2: residual_adjoint_dual_reference_thread<double, E, O, ADJ_NEOHOOKE>(e, a); break;       \
        case 4: residual_adjoint_dual_reference_thread<double, E, O, ADJ_STVK>(e, a); break;           \
        case 5: residual_adjoint_dual_reference_thread<double, E, O, ADJ_TTHERMAL>(e, a); break;       \
        case 6: residual_adjoint_dual_reference_thread<double, E, O, ADJ_ALLENCAHN>(e, a); break;      \
        default: return -3;                                                                            \
      }                                                                                                \
    }                                                                                                  \
    return 0;                                                                                          \
  }
  CASES(X)
#undef X
  return -3;
}

}  // extern "C"
