// Instantiations of the generic element-stage kernel for physics NEOHOOKE (all elements, orders 1-3).
#include "assemble.cuh"

namespace fol {
int assemble_neohooke_f64(cudaStream_t s, int element, int num_gp, const AsmArgs<double>& a) {
  return dispatch_assemble<double, NEOHOOKE>(s, element, num_gp, a);
}
int assemble_neohooke_f32(cudaStream_t s, int element, int num_gp, const AsmArgs<float>& a) {
  return dispatch_assemble<float, NEOHOOKE>(s, element, num_gp, a);
}
}  // namespace fol
