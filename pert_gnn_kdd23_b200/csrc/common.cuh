// Shared device/host helpers for libpertgnn (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pertgnn.h"  // prototypes + PERT_ERR_* (keeps definitions and ABI header in sync)

#define PERT_NUM_SMS 148          // B200: 2 dies x 74 SMs

#define PERT_LAUNCH_CHECK()                          \
  do {                                               \
    cudaError_t e__ = cudaPeekAtLastError();         \
    if (e__ != cudaSuccess) return (int)e__;         \
  } while (0)

static inline int pert_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ float4 ld4(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
__device__ __forceinline__ void st4(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
__device__ __forceinline__ float4 f4add(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 f4fma(float s, float4 a, float4 acc) {
  return make_float4(fmaf(s, a.x, acc.x), fmaf(s, a.y, acc.y), fmaf(s, a.z, acc.z), fmaf(s, a.w, acc.w));
}
__device__ __forceinline__ float f4dot(float4 a, float4 b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ float4 f4max(float4 a, float4 b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4scale(float s, float4 a) {
  return make_float4(s * a.x, s * a.y, s * a.z, s * a.w);
}
// 16-byte vector reduction to global memory (REDG.E.ADD.F32x4, sm_90+).
__device__ __forceinline__ void red4(float* p, float4 v) {
  atomicAdd(reinterpret_cast<float4*>(p), v);
}
// butterfly sum over a sub-warp group of LPR lanes; every lane of the group gets the sum.
template <int LPR>
__device__ __forceinline__ float group_sum(float v, unsigned gmask) {
#pragma unroll
  for (int off = LPR >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(gmask, v, off);
  return v;
}

// engine-internal entry points (not part of the C-ABI)
struct PertTiles {            // graph-aligned tile list of one batch for one row width (csrc/tconv_tile.cu)
  const int* tile_ptr;        // [*ntiles + 1] node boundaries (device)
  const int* ntiles;          // device scalar
  int max_tiles, T, ecap;     // capacity of tile_ptr; nodes / staged edges a tile may hold
};
unsigned int* pert_ticket_slot();   // next slot of the self-resetting ticket ring (csrc/gemm_tc.cu)
long long pert_tile_list_ints(long long N, long long B);
bool pert_tile_fixed_ok(long long N, long long E, long long B, int H, int n_rpc);
int pert_tile_list_view(long long N, long long E, long long B, int H, int n_rpc, int* tiles_mem, PertTiles* out);
int pert_tile_list_bounds(const int64_t* batch, long long N, long long B, int* tiles_mem, cudaStream_t st);
int pert_tile_list_build(int has_batch, long long N, long long E, long long B, const int* rowptr, int H, int n_rpc,
                         int* tiles_mem, PertTiles* out, cudaStream_t st);
int pert_tconv_fwd_stats(const float* q, const float* k, const float* v, const float* s, int ld, const int* rowptr,
                         const int* csr_src, const int* csr_if, const int* csr_rpc, const float* t_if,
                         const float* t_rpc, float* out, int ld_out, float* alpha, int n_rpc, long long N, long long E,
                         long long B_hint, int H, double* bn_acc, int* fused, const PertTiles* tiles, void* stream);
int pert_tconv_bwd_tiles(const float* g, int ld_g, const float* q, const float* k, const float* v, int ld,
                         const int* rowptr, const int* csr_src, const int* csr_if, const int* csr_rpc,
                         const int* colptr, const int* csc_pos, const int* csc_dst, const float* t_if,
                         const float* t_rpc, const float* alpha, float* dq, float* dk, float* dv, int ld_d, float* dsp,
                         float* rpc_ws, float* dt_if, float* dt_rpc, int n_rpc, long long N, long long E,
                         long long B_hint, int H, const PertTiles* tiles, void* stream);
int pert_bn_fwd_ex(const float* x, int ld_x, const float* gamma, const float* beta, float* running_mean,
                   float* running_var, long long* num_batches_tracked, float eps, float momentum, int training,
                   int relu, float* mean, float* rstd, float* y, int ld_y, long long N, int H, void* workspace,
                   long long workspace_bytes, int stats_ready, void* stream);
