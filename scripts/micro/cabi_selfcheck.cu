// Torch-free check of the tuned Hex8 float64 element-stage kernels through the C ABI: the same assembly is run with the
// tuned kernels (the default dispatch: assemble_hex_mech_f64_kernel / assemble_hex_j2_f64_kernel in their default
// layouts) and with the generic kernel (fol_set_tuned_kernels(0)) and the outputs are compared.  A two-second hardware
// check of the shipped libfolax_b200.so that needs neither Python nor the oracle.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o cabi_selfcheck cabi_selfcheck.cu -ldl
//   ./cabi_selfcheck folax_b200/lib/libfolax_b200.so
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <dlfcn.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)

typedef int (*assemble_fn)(void*, int, int, int, int, int, int64_t, int64_t, const void*, const int32_t*, const void*,
                           const void*, const uint8_t*, const double*, void*, void*, const void*, void*);
typedef int (*set_tuned_fn)(int);
typedef const char* (*last_error_fn)(void);

static double lcg(uint64_t& s) {
  s = s * 6364136223846793005ULL + 1442695040888963407ULL;
  return (double)((s >> 11) & ((1ULL << 53) - 1)) / (double)(1ULL << 53);
}

static double max_rel(const std::vector<double>& a, const std::vector<double>& b) {
  double num = 0.0, den = 0.0;
  for (size_t i = 0; i < a.size(); ++i) {
    num = fmax(num, fabs(a[i] - b[i]));
    den = fmax(den, fabs(b[i]));
  }
  return den > 0.0 ? num / den : num;
}

int main(int argc, char** argv) {
  const char* path = argc > 1 ? argv[1] : "folax_b200/lib/libfolax_b200.so";
  void* h = dlopen(path, RTLD_NOW);
  if (!h) { printf("cannot load %s: %s\n", path, dlerror()); return 2; }
  assemble_fn assemble = (assemble_fn)dlsym(h, "fol_assemble_elements");
  set_tuned_fn set_tuned = (set_tuned_fn)dlsym(h, "fol_set_tuned_kernels");
  last_error_fn last_error = (last_error_fn)dlsym(h, "fol_last_error");
  if (!assemble || !set_tuned || !last_error) { printf("missing symbols\n"); return 2; }

  int bad = 0;
  const int dims[2][3] = {{7, 5, 3}, {41, 37, 43}};   // 105 elements (ragged last tile) and 65 231 (several persistent rounds)
  for (int physics : {0 /* FOL_MECHANICAL */, 3 /* FOL_J2PLASTICITY */}) {
    for (int m = 0; m < 2; ++m) {
      const int nx = dims[m][0], ny = dims[m][1], nz = dims[m][2];
      const int64_t nn = (int64_t)(nx + 1) * (ny + 1) * (nz + 1), ne = (int64_t)nx * ny * nz, ndof = 3 * nn;
      auto node = [&](int i, int j, int k) { return (int32_t)((i * (ny + 1) + j) * (nz + 1) + k); };
      uint64_t seed = 1234567 + 17 * m + physics;
      std::vector<double> xyz(3 * nn), ctrl(nn), u(ndof);
      std::vector<uint8_t> flag(ndof, 0);
      for (int i = 0; i <= nx; ++i)
        for (int j = 0; j <= ny; ++j)
          for (int k = 0; k <= nz; ++k) {
            const int32_t n = node(i, j, k);
            const bool inner = i > 0 && i < nx && j > 0 && j < ny && k > 0 && k < nz;
            const double hx = 1.0 / nx, hy = 1.0 / ny, hz = 1.0 / nz;
            xyz[3 * n + 0] = i * hx + (inner ? 0.2 * hx * (lcg(seed) - 0.5) : 0.0);
            xyz[3 * n + 1] = j * hy + (inner ? 0.2 * hy * (lcg(seed) - 0.5) : 0.0);
            xyz[3 * n + 2] = k * hz + (inner ? 0.2 * hz * (lcg(seed) - 0.5) : 0.0);
            ctrl[n] = physics == 3 ? 1.0 : 0.1 + 0.9 * lcg(seed);
            const double amp = physics == 3 ? 0.3 * hx : 0.01;   // J2: strains well past the yield limit in many points
            for (int d = 0; d < 3; ++d) u[3 * n + d] = amp * (lcg(seed) - 0.5);
            if (i == 0 || i == nx)
              for (int d = 0; d < 3; ++d) flag[3 * n + d] = 1;    // Dirichlet faces (row mask path)
          }
      std::vector<int32_t> conn(8 * ne);
      int64_t e = 0;
      for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j)
          for (int k = 0; k < nz; ++k, ++e) {
            const int32_t c[8] = {node(i, j, k),         node(i + 1, j, k),         node(i + 1, j + 1, k),     node(i, j + 1, k),
                                  node(i, j, k + 1),     node(i + 1, j, k + 1),     node(i + 1, j + 1, k + 1), node(i, j + 1, k + 1)};
            for (int a = 0; a < 8; ++a) conn[8 * e + a] = c[a];
          }
      double params[12] = {3.0, 0.3, 0.1, -0.2, 0.3, 0.2, 0.4, 10.0, 0, 0, 0, 0};   // E, nu, body force, J2: y0, h1, h2
      const int64_t nstate = physics == 3 ? ne * 8 * 7 : 0;
      double *d_xyz, *d_ctrl, *d_u, *d_ke, *d_re, *d_s0 = nullptr, *d_s1 = nullptr;
      int32_t* d_conn;
      uint8_t* d_flag;
      CK(cudaMalloc(&d_xyz, xyz.size() * 8)); CK(cudaMalloc(&d_ctrl, nn * 8)); CK(cudaMalloc(&d_u, ndof * 8));
      CK(cudaMalloc(&d_ke, ne * 576 * 8)); CK(cudaMalloc(&d_re, ne * 24 * 8));
      CK(cudaMalloc(&d_conn, conn.size() * 4)); CK(cudaMalloc(&d_flag, ndof));
      if (nstate) { CK(cudaMalloc(&d_s0, nstate * 8)); CK(cudaMalloc(&d_s1, nstate * 8)); CK(cudaMemset(d_s0, 0, nstate * 8)); }
      CK(cudaMemcpy(d_xyz, xyz.data(), xyz.size() * 8, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d_ctrl, ctrl.data(), nn * 8, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d_u, u.data(), ndof * 8, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d_conn, conn.data(), conn.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d_flag, flag.data(), ndof, cudaMemcpyHostToDevice));
      std::vector<double> ke[2], re[2], st[2];
      for (int tuned = 1; tuned >= 0; --tuned) {
        set_tuned(tuned);
        CK(cudaMemset(d_ke, 0xff, ne * 576 * 8)); CK(cudaMemset(d_re, 0xff, ne * 24 * 8));
        if (nstate) CK(cudaMemset(d_s1, 0xff, nstate * 8));
        const int rc = assemble(nullptr, 1 /* FOL_F64 */, physics, 0 /* FOL_HEXAHEDRON */, 2, 0, ne, nn, d_xyz, d_conn, d_ctrl,
                                d_u, d_flag, params, d_ke, d_re, d_s0, d_s1);
        if (rc != 0) { printf("fol_assemble_elements failed (%d): %s\n", rc, last_error()); return 2; }
        CK(cudaDeviceSynchronize());
        ke[tuned].resize(ne * 576); re[tuned].resize(ne * 24); st[tuned].resize(nstate);
        CK(cudaMemcpy(ke[tuned].data(), d_ke, ne * 576 * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(re[tuned].data(), d_re, ne * 24 * 8, cudaMemcpyDeviceToHost));
        if (nstate) CK(cudaMemcpy(st[tuned].data(), d_s1, nstate * 8, cudaMemcpyDeviceToHost));
      }
      set_tuned(1);
      const double eke = max_rel(ke[1], ke[0]), ere = max_rel(re[1], re[0]), est = nstate ? max_rel(st[1], st[0]) : 0.0;
      double plastic = 0.0;
      for (int64_t p = 0; p < nstate / 7; ++p) plastic += st[0][7 * p + 6] > 0.0 ? 1.0 : 0.0;
      const bool ok = eke <= 1e-12 && ere <= 1e-11 && est <= 1e-11 && std::isfinite(eke) && std::isfinite(ere);
      printf("{\"physics\": %d, \"elements\": %lld, \"ke_rel\": %.3e, \"re_rel\": %.3e, \"state_rel\": %.3e, \"plastic_points\": %.3f, \"ok\": %s}\n",
             physics, (long long)ne, eke, ere, est, nstate ? plastic / (nstate / 7) : 0.0, ok ? "true" : "false");
      bad += ok ? 0 : 1;
      cudaFree(d_xyz); cudaFree(d_ctrl); cudaFree(d_u); cudaFree(d_ke); cudaFree(d_re); cudaFree(d_conn); cudaFree(d_flag);
      if (nstate) { cudaFree(d_s0); cudaFree(d_s1); }
    }
  }
  printf(bad ? "SELFCHECK FAILED\n" : "selfcheck ok\n");
  return bad ? 1 : 0;
}
