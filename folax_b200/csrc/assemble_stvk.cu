// Instantiations of the generic element-stage kernel for Saint-Venant-Kirchhoff elasticity
// (mechanical_saint_venant.py:271-305), all elements, orders 1-3.
#include "assemble.cuh"

namespace fol {
int assemble_stvk_f64(cudaStream_t s, int element, int num_gp, const AsmArgs<double>& a) {
  return dispatch_assemble<double, STVK>(s, element, num_gp, a);
}
int assemble_stvk_f32(cudaStream_t s, int element, int num_gp, const AsmArgs<float>& a) {
  return dispatch_assemble<float, STVK>(s, element, num_gp, a);
}
}  // namespace fol
