// Per-thread bodies of the linear-algebra kernels behind folax_b200/solvers (krylov.cu): sparse matrix-vector
// product on the sliced-ELLPACK copy of the duplicate-free CSR, value gather, vector updates.
// __host__ __device__ so that tests/host_shim can loop them on the CPU (same approach as adjoint_threads.cuh).
//
// Layout (built once per mesh by folax_b200/sell_plan.py from the CSR structure of csr_plan.py): rows are cut into
// slices of 32 consecutive rows; a slice of width w (its longest row) owns w*32 entries starting at
// slice_ptr[s], stored COLUMN-major: entry k of row r sits at slice_ptr[r/32] + k*32 + r%32.  One thread per row
// then reads 32 consecutive values / column indices per step -- coalesced without any cross-lane reduction, and
// the per-row sum runs in the fixed CSR order, so products are deterministic.  FE rows have near-uniform length
// (81 entries for interior Hex8 elasticity dofs), so the padding (value 0, column 0) is a few per cent.
#pragma once
#include <math.h>

#include "common.cuh"

namespace fol {

// Sums of products are written with EXPLICIT fused multiply-adds (a x + b y as fma(a, x, b*y); row sums as
// acc = fma(v, x, acc)) in every kernel that forms them: left to the compiler, the contraction could pick a different
// product in two kernels (or none on the host build of the tests), and the scalar- and block-column products, or the
// host-scalar and device-scalar BiCGSTAB loops, would drift apart by an ulp per update instead of being bit-identical.
__host__ __device__ inline double fol_fma(double a, double b, double c) { return fma(a, b, c); }
__host__ __device__ inline float fol_fma(float a, float b, float c) { return fmaf(a, b, c); }

template <class T>
struct SellArgs {
  const long long* slice_ptr;   // (nslices + 1), in entries
  const int32_t* cols;          // padded entries: column 0
  const T* vals;                // padded entries: 0
  const T* x;
  T* y;
  long long nrows;
};

template <class T>
__host__ __device__ inline void sell_spmv_thread(long long row, const SellArgs<T>& a) {
  const long long s = row >> 5;
  const int lane = (int)(row & 31);
  const long long base = a.slice_ptr[s];
  const int width = (int)((a.slice_ptr[s + 1] - base) >> 5);
  T acc = (T)0;
  for (int k = 0; k < width; ++k) {
    const long long idx = base + (long long)k * 32 + lane;
    acc = fol_fma(a.vals[idx], a.x[a.cols[idx]], acc);
  }
  a.y[row] = acc;
}

// Block variant for Jacobians with D dofs per node: the columns of a row come in runs of D consecutive dofs of one
// neighbour node (csr_plan.py builds them that way), so ONE int32 per run -- the neighbour node -- replaces D column
// indices: 8 + 4/D bytes per stored entry instead of 12 (9.33 B at D = 3, -22 %).  Values keep the layout above
// (entry q*D + j of row r at slice_ptr[r/32] + (q*D + j)*32 + r%32); node_cols holds, for run q of row r, the node at
// slice_ptr[r/32]/D + q*32 + r%32.  Same fixed summation order as the scalar kernel => bit-identical results.
template <class T>
struct BlockSellArgs {
  const long long* slice_ptr;
  const int32_t* node_cols;
  const T* vals;
  const T* x;
  T* y;
  long long nrows;
};

template <class T, int D>
__host__ __device__ inline void sell_spmv_block_thread(long long row, const BlockSellArgs<T>& a) {
  const long long s = row >> 5;
  const int lane = (int)(row & 31);
  const long long base = a.slice_ptr[s];
  const int runs = (int)((a.slice_ptr[s + 1] - base) >> 5) / D;
  const long long nbase = base / D;
  T acc = (T)0;
  for (int q = 0; q < runs; ++q) {
    const long long m = a.node_cols[nbase + (long long)q * 32 + lane];
    const T* xm = a.x + m * D;
    const T* v = a.vals + base + (long long)q * (D * 32) + lane;
#pragma unroll
    for (int j = 0; j < D; ++j) acc = fol_fma(v[j * 32], xm[j], acc);
  }
  a.y[row] = acc;
}

// dst[i] = src_index[i] >= 0 ? src[src_index[i]] : 0   (CSR values -> SELL values; CSR values -> diagonal)
template <class T>
__host__ __device__ inline void gather_values_thread(long long i, const int32_t* src_index, const T* src, T* dst) {
  const int32_t k = src_index[i];
  dst[i] = k >= 0 ? src[k] : (T)0;
}

enum : int { VEC_AXPBY = 0, VEC_AXY = 1, VEC_AX_OVER_Y = 2 };


// op 0: out = a x + b y (y may be null when b == 0);  op 1: out = a x * y;  op 2: out = a x / y
template <class T>
__host__ __device__ inline void vec_op_thread(long long i, int op, T a, const T* x, T b, const T* y, T* out) {
  if (op == VEC_AXPBY) out[i] = y ? fol_fma(a, x[i], b * y[i]) : a * x[i];
  else if (op == VEC_AXY) out[i] = a * x[i] * y[i];
  else out[i] = a * x[i] / y[i];
}


// ---- BiCGSTAB with its scalars on the device --------------------------------------------------------------------
// The recurrence coefficients live in a small device array; a one-thread kernel advances them between the vector
// kernels, which read their coefficients through that array and do nothing once the iteration has stopped.  The host
// enqueues whole iterations without reading anything back and looks at (state, k) once per batch -- same iterates,
// same stopping iteration and the same break-down codes as the host-scalar loop of folax_b200/linalg.py.
enum : int { BS_BB = 0, BS_RS, BS_RHO, BS_ALPHA, BS_OMEGA, BS_RHO_NEW, BS_BETA, BS_RQ, BS_SS, BS_TS, BS_TT, BS_ATOL2,
             BS_STATE, BS_K, BS_RS_NEXT, BS_RHO_NEXT, BS_MAXITER, BS_COUNT };
enum : int { BS_RUN = 0, BS_HALF_EXIT = 1, BS_DONE = 2, BS_BROKEN = 3 };            // values of sc[BS_STATE]
enum : int { BSTAGE_TOP = 0, BSTAGE_ALPHA = 1, BSTAGE_HALF = 2, BSTAGE_OMEGA = 3, BSTAGE_END = 4 };

template <class T>
__host__ __device__ inline void bicg_scalar_stage(int stage, T* sc) {
  const int state = (int)sc[BS_STATE];
  if (stage == BSTAGE_TOP) {
    if (state != BS_RUN) return;
    if (!(sc[BS_RS] > sc[BS_ATOL2] && sc[BS_K] >= (T)0 && sc[BS_K] < sc[BS_MAXITER])) {
      sc[BS_STATE] = (T)BS_DONE;
    } else if (sc[BS_RHO_NEW] == (T)0) {
      sc[BS_STATE] = (T)BS_BROKEN;
      sc[BS_K] = (T)-10;
    } else {
      sc[BS_BETA] = sc[BS_RHO_NEW] / sc[BS_RHO] * sc[BS_ALPHA] / sc[BS_OMEGA];
    }
  } else if (stage == BSTAGE_ALPHA) {
    if (state != BS_RUN) return;
    if (sc[BS_RQ] == (T)0) {
      sc[BS_STATE] = (T)BS_BROKEN;
      sc[BS_K] = (T)-11;
    } else {
      sc[BS_ALPHA] = sc[BS_RHO_NEW] / sc[BS_RQ];
    }
  } else if (stage == BSTAGE_HALF) {
    if (state == BS_RUN && sc[BS_SS] < sc[BS_ATOL2]) sc[BS_STATE] = (T)BS_HALF_EXIT;
  } else if (stage == BSTAGE_OMEGA) {
    if (state == BS_RUN) sc[BS_OMEGA] = (sc[BS_TT] != (T)0) ? sc[BS_TS] / sc[BS_TT] : (T)0;
  } else {  // BSTAGE_END
    if (state == BS_HALF_EXIT) {
      sc[BS_RS] = sc[BS_SS];
      sc[BS_RHO] = sc[BS_RHO_NEW];
      sc[BS_K] += (T)1;
      sc[BS_STATE] = (T)BS_DONE;
    } else if (state == BS_RUN) {
      sc[BS_RHO] = sc[BS_RHO_NEW];
      sc[BS_RS] = sc[BS_RS_NEXT];
      sc[BS_RHO_NEW] = sc[BS_RHO_NEXT];
      if (sc[BS_OMEGA] == (T)0 || sc[BS_ALPHA] == (T)0) {
        sc[BS_STATE] = (T)BS_BROKEN;
        sc[BS_K] = (T)-11;
      } else {
        sc[BS_K] += (T)1;
      }
    }
  }
}

// out[i] = a x[i] + b y[i] with a = sa * (ia >= 0 ? sc[ia] : 1), b likewise; executed only while sc[BS_STATE] is one of
// the states in `mask` (bit s set = state s allowed).  y may be null when sb == 0.
template <class T>
__host__ __device__ inline void vec_op_dev_thread(long long i, const T* sc, int mask, int ia, T sa, const T* x, int ib,
                                                  T sb, const T* y, T* out) {
  const int state = (int)sc[BS_STATE];
  if (!((mask >> state) & 1)) return;
  const T a = sa * (ia >= 0 ? sc[ia] : (T)1);
  if (y == nullptr) {
    out[i] = a * x[i];
    return;
  }
  const T b = sb * (ib >= 0 ? sc[ib] : (T)1);
  out[i] = fol_fma(a, x[i], b * y[i]);
}

}  // namespace fol
