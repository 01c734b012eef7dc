// J2 plasticity point update (return mapping + forward-mode tangent); defined in a later step.
#pragma once
namespace fol {
// eps: total strain in the Voigt order of the linear B matrix (engineering shears), state: history
// [eps_p (V), xi]; writes sigma (V), tangent d sigma / d eps (V*V), new state.
template <class T, int D>
__device__ void j2_point(const T* eps, const T* state, T E, T nu, T y0, T h1, T h2, T* sigma, T* tangent,
                         T* state_new);
}  // namespace fol
