// Argument block of the AD-variant element stage (assemble_ad.cu), shared with api.cu.
#pragma once
#include "common.cuh"

namespace fol {

template <class T>
struct AdAsmArgs {
  const T* xyz;
  const int32_t* conn;
  const T* ctrl;
  const T* u;
  const uint8_t* dir;
  T* ke;        // (ne, nd, nd) masked (and transposed first when `transpose`), fe_loss.py:191-230, 299
  T* re;        // (ne, nd) masked element residuals
  T* energy;    // (ne) element energies (first return value of ComputeElement) or null
  long long ne;
  int transpose;
  Params<T> p;
};

}  // namespace fol
