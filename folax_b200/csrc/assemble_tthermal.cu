// Instantiations of the generic element-stage kernel for physics TTHERMAL (all elements, orders 1-3).
#include "assemble.cuh"

namespace fol {
int assemble_tthermal_f64(cudaStream_t s, int element, int num_gp, const AsmArgs<double>& a) {
  return dispatch_assemble<double, TTHERMAL>(s, element, num_gp, a);
}
int assemble_tthermal_f32(cudaStream_t s, int element, int num_gp, const AsmArgs<float>& a) {
  return dispatch_assemble<float, TTHERMAL>(s, element, num_gp, a);
}
}  // namespace fol
