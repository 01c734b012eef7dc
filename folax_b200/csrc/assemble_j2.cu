// Instantiations of the generic element-stage kernel for J2 elastoplasticity with Gauss-point history
// (mechanical_elastoplasticity.py:153-235), all elements, orders 1-3.
#include "assemble.cuh"

namespace fol {
int assemble_j2_f64(cudaStream_t s, int element, int num_gp, const AsmArgs<double>& a) {
  return dispatch_assemble<double, J2>(s, element, num_gp, a);
}
int assemble_j2_f32(cudaStream_t s, int element, int num_gp, const AsmArgs<float>& a) {
  return dispatch_assemble<float, J2>(s, element, num_gp, a);
}
}  // namespace fol
