// Dev tool (not part of the library): per-role clock64 timeline of CTA (0,0) of the tcgen05 GEMM kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DPERT_TC_TRACE -o gemm_trace gemm_trace.cu
#include "../gemm_tc.cu"
#include <cstdio>
#include <vector>

static void dump(const char* name, int roles, int iters, int evs, float ms) {
  long long h[6][32][4];
  cudaMemcpyFromSymbol(h, g_trace, sizeof(h));
  long long t0 = h[0][0][0];
  printf("== %s  %.2f us/launch (cycles relative to first producer stamp)\n", name, ms * 1000.f);
  for (int r = 0; r < roles; ++r)
    for (int i = 0; i < iters; ++i) {
      printf("  role %d it %2d:", r, i);
      for (int e = 0; e < evs; ++e) printf(" %8lld", h[r][i][e] ? h[r][i][e] - t0 : -1);
      printf("\n");
    }
  static unsigned long long ct[512][2];
  cudaMemcpyFromSymbol(ct, g_cta_t, sizeof(ct));
  unsigned long long lo = ~0ull, hi = 0, lo_exit = ~0ull, hi_entry = 0;
  int n = 0;
  for (int i = 0; i < 512; ++i)
    if (ct[i][0]) {
      ++n;
      if (ct[i][0] < lo) lo = ct[i][0];
      if (ct[i][0] > hi_entry) hi_entry = ct[i][0];
      if (ct[i][1] > hi) hi = ct[i][1];
      if (ct[i][1] < lo_exit) lo_exit = ct[i][1];
    }
  printf("  %d CTAs: first entry 0, last entry +%llu ns, first exit +%llu ns, last exit +%llu ns; CTA0 entry +%llu exit +%llu\n", n,
         hi_entry - lo, lo_exit - lo, hi - lo, ct[0][0] - lo, ct[0][1] - lo);
  static unsigned long long zc[512][2];
  cudaMemcpyToSymbol(g_cta_t, zc, sizeof(zc));
  static long long z[6][32][4];
  cudaMemcpyToSymbol(g_trace, z, sizeof(z));
}

int main() {
  const long long N = 51200;
  const int H = 64;
  float *X, *W, *P, *dX, *dW, *bias, *cs;
  cudaMalloc(&X, N * H * 4); cudaMalloc(&W, 4 * H * H * 4); cudaMalloc(&P, 4 * N * H * 4);
  cudaMalloc(&dX, N * H * 4); cudaMalloc(&dW, 4 * H * H * 4); cudaMalloc(&bias, 4 * H * 4); cudaMalloc(&cs, 4 * H * 4);
  std::vector<float> hx(N * H);
  for (size_t i = 0; i < hx.size(); ++i) hx[i] = (float)((i * 2654435761u) >> 8 & 0xffff) / 65536.f - 0.5f;
  cudaMemcpy(X, hx.data(), N * H * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(W, hx.data(), 4 * H * H * 4, cudaMemcpyHostToDevice);
  cudaMemset(bias, 0, 4 * H * 4); cudaMemset(dW, 0, 4 * H * H * 4); cudaMemset(cs, 0, 4 * H * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    int rc = 0;
    for (int k = 0; k < 10; ++k) rc |= pert_gemm_nt_tc(X, H, 0, 0, W, H, bias, P, H, H, N * H, N, 4 * H, H, 0, 0);   // train of 10: launch gaps amortised
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    ms /= 10.f;
    if (rc) printf("nt fwd rc %d\n", rc);
  }
  dump("NT forward [N,64]x[256,64]^T -> planes", 6, 8, 4, ms);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    // data gradient: dX[N,64] = dP[N,256 blocked] . Wt[64,256]^T   (B rows = output features, K = 256)
    int rc = 0;
    for (int k = 0; k < 10; ++k) rc |= pert_gemm_nt_tc(P, H, H, N * H, W, 4 * H, nullptr, dX, H, 0, 0, N, H, 4 * H, 0, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    ms /= 10.f;
    if (rc) printf("nt dgrad rc %d\n", rc);
  }
  dump("NT dgrad planes x [64,256]^T -> [N,64]", 6, 14, 4, ms);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    int rc = 0;
    for (int k = 0; k < 10; ++k) rc |= pert_gemm_tn_tc(P, H, H, N * H, X, H, 0, 0, dW, H, cs, N, 4 * H, H, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    ms /= 10.f;
    if (rc) printf("tn rc %d\n", rc);
  }
  dump("TN wgrad planes^T x X -> [256,64]", 4, 12, 3, ms);
  long long h[4][32][4];
  (void)h;
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
