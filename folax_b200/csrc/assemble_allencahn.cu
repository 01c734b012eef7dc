// Instantiations of the generic element-stage kernel for physics ALLENCAHN (all elements, orders 1-3).
#include "assemble.cuh"

namespace fol {
int assemble_allencahn_f64(cudaStream_t s, int element, int num_gp, const AsmArgs<double>& a) {
  return dispatch_assemble<double, ALLENCAHN>(s, element, num_gp, a);
}
int assemble_allencahn_f32(cudaStream_t s, int element, int num_gp, const AsmArgs<float>& a) {
  return dispatch_assemble<float, ALLENCAHN>(s, element, num_gp, a);
}
}  // namespace fol
