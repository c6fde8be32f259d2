// Probe 2: both operands MN-major from smem (no swizzle): D[128,N] = sum_r A[r][m] * B[r][n]  (the weight-gradient form).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  uint32_t d = 0;
  d |= 1u << 4; d |= 2u << 7; d |= 2u << 10;
  d |= (uint32_t)a_mn << 15; d |= (uint32_t)b_mn << 16;
  d |= (uint32_t)(N >> 3) << 17; d |= (uint32_t)(M >> 4) << 24;
  return d;
}
// MN-major canonical (no swizzle): (mn, k) -> (mn%4)*4 + (k%8)*16 + (mn/4)*SBO + (k/8)*LBO   [bytes]
template <int N, int R, int VARIANT, int A_MN, int B_MN>
__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  unsigned char* sA = smem;                       // 128 x R
  unsigned char* sB = smem + 128 * R * 4;         // N x R
  const int tid = threadIdx.x, warp = tid >> 5;
  // VARIANT 0: SBO = 128 (mn blocks adjacent), LBO = (MN/4)*128 ; VARIANT 1: LBO = 128 (k blocks adjacent), SBO = (R/8)*128
  const uint32_t a_sbo = A_MN ? (VARIANT == 0 ? 128 : (R / 8) * 128) : (R / 4) * 128;
  const uint32_t a_lbo = A_MN ? (VARIANT == 0 ? (128 / 4) * 128 : 128) : 128;
  const uint32_t b_sbo = B_MN ? (VARIANT == 0 ? 128 : (R / 8) * 128) : (R / 4) * 128;
  const uint32_t b_lbo = B_MN ? (VARIANT == 0 ? (N / 4) * 128 : 128) : 128;
  const uint32_t a_step = A_MN ? a_lbo : 2 * a_lbo, b_step = B_MN ? b_lbo : 2 * b_lbo;   // bytes per K=8 step
  if (A_MN) {
    for (int idx = tid; idx < R * 32; idx += 128) {       // A[r][m..m+3]
      int r = idx / 32, mq = idx % 32;
      float4 v = *reinterpret_cast<const float4*>(A + (size_t)r * 128 + mq * 4);
      *reinterpret_cast<float4*>(sA + (r % 8) * 16 + mq * a_sbo + (r / 8) * a_lbo) = v;
    }
  } else {                                                 // K-major: element (m, r) at (m/8)*SBO + (m%8)*16 + (r/4)*LBO + (r%4)*4
    for (int idx = tid; idx < R * 128; idx += 128) {
      int r = idx / 128, m = idx % 128;
      *reinterpret_cast<float*>(sA + (m / 8) * a_sbo + (m % 8) * 16 + (r / 4) * a_lbo + (r % 4) * 4) = A[(size_t)r * 128 + m];
    }
  }
  if (B_MN) {
    for (int idx = tid; idx < R * (N / 4); idx += 128) {
      int r = idx / (N / 4), nq = idx % (N / 4);
      float4 v = *reinterpret_cast<const float4*>(B + (size_t)r * N + nq * 4);
      *reinterpret_cast<float4*>(sB + (r % 8) * 16 + nq * b_sbo + (r / 8) * b_lbo) = v;
    }
  } else {
    for (int idx = tid; idx < R * N; idx += 128) {
      int r = idx / N, n = idx % N;
      *reinterpret_cast<float*>(sB + (n / 8) * b_sbo + (n % 8) * 16 + (r / 4) * b_lbo + (r % 4) * 4) = B[(size_t)r * N + n];
    }
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, N, A_MN, B_MN);
    for (int s = 0; s < R / 8; ++s) {
      const uint64_t adesc = make_desc(smem_u32(sA) + s * a_step, a_lbo, a_sbo);
      const uint64_t bdesc = make_desc(smem_u32(sB) + s * b_step, b_lbo, b_sbo);
      const uint32_t acc = s > 0;
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                   "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc),
                   "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tmem + lane_off + c0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}
static float tf32r(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; float y; memcpy(&y, &u, 4); return y; }
template <int N, int R, int V, int A_MN, int B_MN>
int run() {
  std::vector<float> A(R * 128), B(R * N), D(128 * N), Rf(128 * N);
  srand(2);
  for (auto& x : A) x = tf32r((rand() % 2001 - 1000) / 500.f);
  for (auto& x : B) x = tf32r((rand() % 2001 - 1000) / 500.f);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) { double s = 0; for (int r = 0; r < R; ++r) s += (double)A[r * 128 + m] * B[r * N + n]; Rf[m * N + n] = (float)s; }
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, D.size() * 4));
  size_t smem = (size_t)(128 + N) * R * 4;
  CK(cudaFuncSetAttribute(probe<N, R, V, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe<N, R, V, A_MN, B_MN><<<1, 128, smem>>>(dA, dB, dD);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (size_t i = 0; i < D.size(); ++i) { maxerr = fmax(maxerr, fabs(D[i] - Rf[i])); maxref = fmax(maxref, fabs(Rf[i])); }
  printf("A_MN=%d B_MN=%d N=%d R=%d variant=%d : max abs err %.3e (max |ref| %.3e) D[0..2]=%.4f %.4f %.4f ref %.4f %.4f %.4f\n", A_MN, B_MN, N, R, V, maxerr,
         maxref, D[0], D[1], D[2], Rf[0], Rf[1], Rf[2]);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return maxerr < 1e-3 * maxref ? 0 : 1;
}
int main() {
  run<64, 32, 0, 0, 0>();
  run<64, 32, 0, 0, 1>();
  run<64, 32, 1, 0, 1>();
  run<64, 32, 0, 1, 0>();
  run<64, 32, 1, 1, 0>();
  run<64, 32, 0, 1, 1>();
  run<64, 32, 1, 1, 1>();
  return 0;
}
