// Host-side integer plans of the solver hand-off, built once per mesh (no device work in this file):
//   * fol_csr_plan_count_host / fol_csr_plan_fill_host: structure of the duplicate-free CSR of the BCOO Jacobian and the
//     fixed-order value plan fol_csr_values consumes (what scipy.sparse.csr_array + sum_duplicates gives the
//     reference's solvers, fe_solver.py:71-72); folax_b200/csr_plan.py holds the NumPy restatement the tests compare
//     this against, entry for entry;
//   * fol_sell_plan_fill_host: sliced-ELLPACK copy of that CSR for the Krylov solvers (folax_b200/sell_plan.py).
// Node rows are independent, so both run on all host threads.  128^3 Hex8 (134 M contributors, 514 M non-zeros):
// seconds instead of the ~25 s of the NumPy versions.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <thread>
#include <utility>
#include <vector>

#include "common.cuh"

namespace {

template <class F>
void parallel_ranges(int64_t n, F f) {   // f(begin, end) on disjoint ranges covering [0, n)
  unsigned nt = std::thread::hardware_concurrency();
  nt = nt < 1 ? 1 : (nt > 32 ? 32 : nt);
  if (n < (1 << 14)) nt = 1;
  // many more ranges than threads: rows differ in cost, a shared counter balances them
  const int64_t grain = std::max<int64_t>(1, n / (int64_t)(nt * 16));
  std::atomic<int64_t> next{0};
  auto worker = [&]() {
    for (;;) {
      const int64_t b = next.fetch_add(grain);
      if (b >= n) break;
      f(b, std::min(n, b + grain));
    }
  };
  if (nt == 1) { worker(); return; }
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t) th.emplace_back(worker);
  for (auto& x : th) x.join();
}

}  // namespace

extern "C" {

int fol_csr_plan_count_host(const int32_t* conn, int64_t ne, int nnode, int64_t nn, int32_t* adj_ptr, int32_t* adj,
                            int32_t* deg) {
  FOL_REQUIRE(conn && adj_ptr && adj && deg, "fol_csr_plan_count_host: null pointer");
  FOL_REQUIRE(ne >= 0 && nn >= 0 && nnode >= 1 && nnode <= 64, "fol_csr_plan_count_host: bad sizes");
  FOL_REQUIRE(ne * nnode < (1LL << 31), "fol_csr_plan_count_host: adjacency exceeds int32 indexing");
  const int64_t m = ne * nnode;
  for (int64_t i = 0; i <= nn; ++i) adj_ptr[i] = 0;
  for (int64_t i = 0; i < m; ++i) {
    if (conn[i] < 0 || conn[i] >= nn) return ::fol::fail(FOL_ERR_INVALID, "fol_csr_plan_count_host: node index out of range");
    ++adj_ptr[conn[i] + 1];
  }
  for (int64_t i = 0; i < nn; ++i) adj_ptr[i + 1] += adj_ptr[i];
  {
    std::vector<int32_t> fill(adj_ptr, adj_ptr + nn);
    for (int64_t i = 0; i < m; ++i) adj[fill[conn[i]]++] = (int32_t)i;   // ascending e * nnode + a within a node
  }
  parallel_ranges(nn, [&](int64_t n0, int64_t n1) {
    std::vector<int32_t> cols;
    for (int64_t n = n0; n < n1; ++n) {
      cols.clear();
      for (int32_t k = adj_ptr[n]; k < adj_ptr[n + 1]; ++k) {
        const int32_t* ce = conn + (int64_t)(adj[k] / nnode) * nnode;
        cols.insert(cols.end(), ce, ce + nnode);
      }
      std::sort(cols.begin(), cols.end());
      deg[n] = (int32_t)(std::unique(cols.begin(), cols.end()) - cols.begin());
    }
  });
  return FOL_OK;
}

int fol_csr_plan_fill_host(const int32_t* conn, int64_t ne, int nnode, int64_t nn, int d, const int32_t* adj_ptr,
                           const int32_t* adj, const int64_t* node_ptr, int32_t* pair_ptr, int32_t* contrib,
                           int32_t* out_base, int32_t* row_stride, int32_t* indptr, int32_t* indices) {
  FOL_REQUIRE(conn && adj_ptr && adj && node_ptr && pair_ptr && contrib && out_base && row_stride && indptr && indices,
              "fol_csr_plan_fill_host: null pointer");
  FOL_REQUIRE(d >= 1 && d <= 3 && nnode >= 1 && nnode <= 64, "fol_csr_plan_fill_host: bad sizes");
  const int64_t npairs = node_ptr[nn];
  FOL_REQUIRE(ne * nnode * nnode < (1LL << 31) && npairs * d * d < (1LL << 31),
              "fol_csr_plan_fill_host: plan exceeds int32 indexing");
  parallel_ranges(nn, [&](int64_t n0, int64_t n1) {
    std::vector<std::pair<int32_t, int32_t>> ent;   // (column node, contributor (e * nnode + a) * nnode + b)
    for (int64_t n = n0; n < n1; ++n) {
      ent.clear();
      for (int32_t k = adj_ptr[n]; k < adj_ptr[n + 1]; ++k) {
        const int32_t ea = adj[k];
        const int32_t* ce = conn + (int64_t)(ea / nnode) * nnode;
        for (int b = 0; b < nnode; ++b) ent.emplace_back(ce[b], ea * nnode + b);
      }
      // generated in ascending contributor order: a stable sort by column keeps that order inside every pair
      std::stable_sort(ent.begin(), ent.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
      const int64_t p0 = node_ptr[n], deg = node_ptr[n + 1] - p0;
      const int64_t c0 = (int64_t)adj_ptr[n] * nnode;          // every (e, a) of the node contributes nnode entries
      const int64_t row0 = (int64_t)d * d * p0, stride = (int64_t)d * deg;
      for (int i = 0; i < d; ++i) indptr[d * n + i] = (int32_t)(row0 + i * stride);
      int64_t q = -1;
      for (size_t t = 0; t < ent.size(); ++t) {
        if (t == 0 || ent[t].first != ent[t - 1].first) {
          ++q;
          pair_ptr[p0 + q] = (int32_t)(c0 + (int64_t)t);
          out_base[p0 + q] = (int32_t)(row0 + q * d);
          row_stride[p0 + q] = (int32_t)stride;
          for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) indices[row0 + i * stride + q * d + j] = d * ent[t].first + j;
        }
        contrib[c0 + (int64_t)t] = ent[t].second;
      }
    }
  });
  pair_ptr[npairs] = (int32_t)(ne * nnode * nnode);
  indptr[(int64_t)d * nn] = (int32_t)((int64_t)d * d * npairs);
  return FOL_OK;
}

int fol_sell_plan_fill_host(const int64_t* indptr, const int32_t* indices, int64_t nrows, int d, int slice_height,
                            const int64_t* slice_ptr, int32_t* cols, int32_t* src, int32_t* diag_src,
                            int32_t* node_cols, int* blocked) {
  FOL_REQUIRE(indptr && slice_ptr && cols && src && diag_src && blocked, "fol_sell_plan_fill_host: null pointer");
  FOL_REQUIRE(nrows >= 0 && slice_height >= 1 && d >= 1, "fol_sell_plan_fill_host: bad sizes");
  FOL_REQUIRE(indices || indptr[nrows] == 0, "fol_sell_plan_fill_host: null indices");
  const int64_t S = slice_height, nsl = (nrows + S - 1) / S;
  std::atomic<int> fits{node_cols ? 1 : 0};
  parallel_ranges(nsl, [&](int64_t s0, int64_t s1) {
    for (int64_t s = s0; s < s1; ++s) {
      const int64_t base = slice_ptr[s];
      for (int64_t r = s * S; r < std::min(nrows, (s + 1) * S); ++r) {
        const int64_t e0 = indptr[r], e1 = indptr[r + 1], lane = r - s * S;
        for (int64_t e = e0; e < e1; ++e) {
          const int64_t pos = base + (e - e0) * S + lane;
          cols[pos] = indices[e];
          src[pos] = (int32_t)e;
          if (indices[e] == r) diag_src[r] = (int32_t)e;
        }
        if (node_cols && fits.load(std::memory_order_relaxed)) {
          // rows made of runs of d consecutive dofs of one node: one node column per run (fol_sell_spmv_block)
          bool ok = (e1 - e0) % d == 0;
          for (int64_t e = e0; ok && e < e1; e += d) {
            ok = indices[e] % d == 0;
            for (int j = 1; ok && j < d; ++j) ok = indices[e + j] == indices[e] + j;
          }
          if (!ok) {
            fits.store(0, std::memory_order_relaxed);
          } else {
            for (int64_t e = e0; e < e1; e += d)
              node_cols[base / d + ((e - e0) / d) * S + lane] = indices[e] / d;
          }
        }
      }
    }
  });
  *blocked = fits.load();
  return FOL_OK;
}

}  // extern "C"
