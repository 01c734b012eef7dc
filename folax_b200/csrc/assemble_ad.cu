// Kernel + launcher of the *_AD.py loss variants (assemble_ad_threads.cuh): one element per thread, stiffness by
// forward-mode sweeps.  Reached through fol_assemble_elements with physics FOL_NEOHOOKE_AD / FOL_STVENANT_AD.
#include "assemble_ad_threads.cuh"

namespace fol {

template <class T, int ELEM, int ORDER, int LAW>
__global__ void __launch_bounds__(64) assemble_ad_kernel(const AdAsmArgs<T> a) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < a.ne) assemble_ad_thread<T, ELEM, ORDER, LAW>(e, a);
}

template <class T, int LAW>
static int launch_ad(cudaStream_t s, int element, int num_gp, const AdAsmArgs<T>& a) {
  if (a.ne == 0) return FOL_OK;
  const unsigned grid = (unsigned)cdiv(a.ne, 64);
#define X(E, O)                                                   \
  if (element == E && num_gp == O) {                              \
    assemble_ad_kernel<T, E, O, LAW><<<grid, 64, 0, s>>>(a);      \
    return check_launch("assemble_ad_kernel");                    \
  }
  X(HEX, 1) X(HEX, 2) X(HEX, 3) X(QUAD, 1) X(QUAD, 2) X(QUAD, 3) X(TET, 1) X(TET, 2) X(TET, 3)
  X(TRI, 1) X(TRI, 2) X(TRI, 3)
#undef X
  return fail(FOL_ERR_UNSUPPORTED, "fol_assemble_elements (AD variant): unsupported element / num_gp");
}

template <class T>
int assemble_ad(cudaStream_t s, int physics, int element, int num_gp, const AdAsmArgs<T>& a) {
  if (physics == LAW_NEOHOOKE_AD) return launch_ad<T, LAW_NEOHOOKE_AD>(s, element, num_gp, a);
  if (physics == LAW_STVK_AD) return launch_ad<T, LAW_STVK_AD>(s, element, num_gp, a);
  return fail(FOL_ERR_UNSUPPORTED, "assemble_ad: unknown law");
}

template int assemble_ad<double>(cudaStream_t, int, int, int, const AdAsmArgs<double>&);
template int assemble_ad<float>(cudaStream_t, int, int, int, const AdAsmArgs<float>&);

}  // namespace fol
