// Standalone probe for the tcgen05 building blocks used by gemm_tc.cu: TMEM alloc, tcgen05.st (A operand in TMEM),
// no-swizzle K-major smem descriptor (B operand), kind::tf32 MMA, commit -> mbarrier, tcgen05.ld epilogue.
// D[128, N] = A[128, K] . B[N, K]^T with inputs pre-rounded to tf32 so the result must match fp32 CPU closely.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu && timeout 60 ./umma_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// no-swizzle K-major canonical layout: 8x16B core matrices; LBO = byte stride between core matrices adjacent in K,
// SBO = byte stride between core matrices adjacent in M/N.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;    // D format F32
  d |= 2u << 7;    // A format TF32
  d |= 2u << 10;   // B format TF32
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;        // a_major = b_major = 0 (K-major)
}

template <int N, int K, bool A_IN_TMEM>
__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  float* sB = reinterpret_cast<float*>(smem);                 // N x K tf32, canonical
  float* sA = reinterpret_cast<float*>(smem + N * K * 4);     // 128 x K (only when !A_IN_TMEM)
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr uint32_t LBO = 128, SBO = (K / 4) * 128;
  // B -> smem canonical
  for (int idx = tid; idx < N * (K / 4); idx += 128) {
    int n = idx / (K / 4), kc = idx % (K / 4);
    float4 v = *reinterpret_cast<const float4*>(B + (size_t)n * K + kc * 4);
    *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(sB) + (n / 8) * SBO + (n % 8) * 16 + kc * LBO) = v;
  }
  if (!A_IN_TMEM) {
    for (int kc = 0; kc < K / 4; ++kc) {
      float4 v = *reinterpret_cast<const float4*>(A + (size_t)tid * K + kc * 4);
      *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(sA) + (tid / 8) * SBO + (tid % 8) * 16 + kc * LBO) = v;
    }
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the MMA (async proxy)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t t_d = tmem;            // columns [0, N)
  const uint32_t t_a = tmem + 256;      // columns [256, 256+K)
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  if (A_IN_TMEM) {
    // thread = row; registers = K consecutive elements -> TMEM columns
    for (int k0 = 0; k0 < K; k0 += 8) {
      float4 v0 = *reinterpret_cast<const float4*>(A + (size_t)tid * K + k0);
      float4 v1 = *reinterpret_cast<const float4*>(A + (size_t)tid * K + k0 + 4);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(t_a + lane_off + k0),
                   "r"(__float_as_uint(v0.x)), "r"(__float_as_uint(v0.y)), "r"(__float_as_uint(v0.z)), "r"(__float_as_uint(v0.w)),
                   "r"(__float_as_uint(v1.x)), "r"(__float_as_uint(v1.y)), "r"(__float_as_uint(v1.z)), "r"(__float_as_uint(v1.w))
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = make_idesc(128, N);
    for (int s = 0; s < K / 8; ++s) {
      const uint64_t bdesc = make_desc(smem_u32(sB) + s * 2 * LBO, LBO, SBO);
      const uint32_t acc = s > 0;
      if (A_IN_TMEM) {
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}" ::"r"(t_d), "r"(t_a + s * 8), "l"(bdesc),
                     "r"(idesc), "r"(acc) : "memory");
      } else {
        const uint64_t adesc = make_desc(smem_u32(sA) + s * 2 * LBO, LBO, SBO);
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(t_d), "l"(adesc), "l"(bdesc),
                     "r"(idesc), "r"(acc) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait for the MMAs
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // epilogue: thread = row, 8 columns per ld
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(t_d + lane_off + c0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

static float tf32r(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; float y; memcpy(&y, &u, 4); return y; }

template <int N, int K, bool AT>
int run() {
  std::vector<float> A(128 * K), B(N * K), D(128 * N), R(128 * N);
  srand(1);
  for (auto& x : A) x = tf32r((rand() % 2001 - 1000) / 500.f);
  for (auto& x : B) x = tf32r((rand() % 2001 - 1000) / 500.f);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k]; R[m * N + n] = (float)s; }
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, D.size() * 4));
  size_t smem = (size_t)N * K * 4 + 128 * K * 4;
  CK(cudaFuncSetAttribute(probe<N, K, AT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe<N, K, AT><<<1, 128, smem>>>(dA, dB, dD);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (size_t i = 0; i < D.size(); ++i) { maxerr = fmax(maxerr, fabs(D[i] - R[i])); maxref = fmax(maxref, fabs(R[i])); }
  printf("N=%d K=%d A_in_tmem=%d : max abs err %.3e (max |ref| %.3e)  D[0..3]=%.4f %.4f %.4f %.4f ref %.4f %.4f %.4f %.4f\n", N, K,
         (int)AT, maxerr, maxref, D[0], D[1], D[2], D[3], R[0], R[1], R[2], R[3]);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return maxerr < 1e-3 * maxref ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run<64, 64, false>();
  bad += run<64, 64, true>();
  bad += run<128, 64, true>();
  bad += run<256, 32, true>();
  bad += run<80, 16, true>();
  printf(bad ? "PROBE FAILED (%d)\n" : "PROBE OK\n", bad);
  return bad;
}
