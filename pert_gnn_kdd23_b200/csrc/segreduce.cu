// Segmented reduce over CSR segments: out[i, :] = reduce_{p in [rowptr[i], rowptr[i+1])} msg[row(p), :]
// with reduce = max | sum, any width.  This is the "level-wise scatter-max / scatter-add" kernel of
// BASELINE.json: the generalisation (width H) of the reference's segment-max / segment-sum inside PyG
// utils.softmax and of aggr='add' / global_add_pool.
//
// Semantics = torch_geometric.utils.scatter (2.4.0): segments that receive nothing yield 0
// (zeros + scatter_reduce_(include_self=False) / scatter_add_).
//
// row(p) = p when the messages are already in CSR (target-sorted, level-major) order -- the layout the
// collation produces -- or perm[p] when they are in original COO order (gather through the stable
// permutation).  HBM-bound: algorithmic bytes = 4*E*H (msg, read once) + 4*(N+1) (rowptr; 4*E more with
// perm) + 4*N*H (out).  A group of LPR lanes owns a segment; rows of consecutive segments are contiguous,
// so a warp streams one contiguous chunk with 16-byte loads, 4 rows in flight per group; loads bypass L1
// allocation (read-once stream), stores are streaming.
#include "common.cuh"
#include <type_traits>

namespace {

__device__ __forceinline__ float4 ld_stream4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st_stream4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

template <bool IS_MAX>
__device__ __forceinline__ float4 comb(float4 a, float4 b) {
  return IS_MAX ? f4max(a, b) : f4add(a, b);
}

// argmax tracking variant is separate (training of a max-aggregation); the metric kernel is the plain one.
template <int LPR, int VPL, bool IS_MAX, bool HAS_PERM>
__global__ void __launch_bounds__(256) k_segreduce(const float* __restrict__ msg, const int* __restrict__ rowptr,
                                                   const int* __restrict__ perm, float* __restrict__ out, int N) {
  constexpr int H = 4 * LPR * VPL;
  constexpr int GPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int lig = lane % LPR;
  const int grp = lane / LPR;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int i = warp * GPW + grp;
  if (i >= N) return;
  const int p0 = __ldg(rowptr + i), p1 = __ldg(rowptr + i + 1);
  const float init = IS_MAX ? -INFINITY : 0.f;
  float4 acc[VPL];
#pragma unroll
  for (int u = 0; u < VPL; ++u) acc[u] = make_float4(init, init, init, init);
  int p = p0;
  for (; p + 4 <= p1; p += 4) {
    float4 r[4][VPL];
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const size_t row = HAS_PERM ? (size_t)__ldg(perm + p + x) : (size_t)(p + x);
#pragma unroll
      for (int u = 0; u < VPL; ++u) r[x][u] = ld_stream4(msg + row * H + lig * 4 + u * LPR * 4);
    }
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int u = 0; u < VPL; ++u) acc[u] = comb<IS_MAX>(acc[u], r[x][u]);
  }
  {
    // tail of up to 3 rows, all loads issued before the first use
    float4 r[3][VPL];
    const int rem = p1 - p;
#pragma unroll
    for (int x = 0; x < 3; ++x) {
      if (x < rem) {
        const size_t row = HAS_PERM ? (size_t)__ldg(perm + p + x) : (size_t)(p + x);
#pragma unroll
        for (int u = 0; u < VPL; ++u) r[x][u] = ld_stream4(msg + row * H + lig * 4 + u * LPR * 4);
      }
    }
#pragma unroll
    for (int x = 0; x < 3; ++x)
      if (x < rem) {
#pragma unroll
        for (int u = 0; u < VPL; ++u) acc[u] = comb<IS_MAX>(acc[u], r[x][u]);
      }
  }
  if (IS_MAX && p0 == p1) {
#pragma unroll
    for (int u = 0; u < VPL; ++u) acc[u] = f4zero();
  }
#pragma unroll
  for (int u = 0; u < VPL; ++u) st_stream4(out + (size_t)i * H + lig * 4 + u * LPR * 4, acc[u]);
}

// any width (incl. the reference's literal [E,1] case): one thread per (segment, column)
template <bool IS_MAX>
__global__ void k_segreduce_generic(const float* __restrict__ msg, const int* __restrict__ rowptr,
                                    const int* __restrict__ perm, float* __restrict__ out, long long total, int H) {
  long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= total) return;
  int i = (int)(id / H), c = (int)(id % H);
  int p0 = rowptr[i], p1 = rowptr[i + 1];
  float acc = IS_MAX ? -INFINITY : 0.f;
  for (int p = p0; p < p1; ++p) {
    size_t row = perm ? (size_t)perm[p] : (size_t)p;
    float x = msg[row * H + c];
    acc = IS_MAX ? fmaxf(acc, x) : acc + x;
  }
  if (IS_MAX && p0 == p1) acc = 0.f;
  out[id] = acc;
}

// backward: sum -> dmsg[row(p)] = dout[seg(p)];  max -> dout routed to the arg-max row(s).
// PyG/ATen 'amax' backward splits the gradient evenly among ties; we do the same.
template <bool IS_MAX>
__global__ void k_segreduce_bwd(const float* __restrict__ dout, const float* __restrict__ msg,
                                const float* __restrict__ out, const int* __restrict__ rowptr,
                                const int* __restrict__ perm, float* __restrict__ dmsg, long long total, int H) {
  long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= total) return;
  int i = (int)(id / H), c = (int)(id % H);
  int p0 = rowptr[i], p1 = rowptr[i + 1];
  float g = dout[id];
  if (!IS_MAX) {
    for (int p = p0; p < p1; ++p) {
      size_t row = perm ? (size_t)perm[p] : (size_t)p;
      dmsg[row * H + c] = g;
    }
  } else {
    float mx = out[id];
    int ties = 0;
    for (int p = p0; p < p1; ++p) {
      size_t row = perm ? (size_t)perm[p] : (size_t)p;
      ties += (msg[row * H + c] == mx);
    }
    float share = ties > 0 ? g / (float)ties : 0.f;
    for (int p = p0; p < p1; ++p) {
      size_t row = perm ? (size_t)perm[p] : (size_t)p;
      dmsg[row * H + c] = (msg[row * H + c] == mx) ? share : 0.f;
    }
  }
}

template <typename F>
int dispatch_w(int H, F&& f) {
  switch (H) {
    case 4: return f(std::integral_constant<int, 1>{}, std::integral_constant<int, 1>{});
    case 8: return f(std::integral_constant<int, 2>{}, std::integral_constant<int, 1>{});
    case 16: return f(std::integral_constant<int, 4>{}, std::integral_constant<int, 1>{});
    case 32: return f(std::integral_constant<int, 8>{}, std::integral_constant<int, 1>{});
    case 64: return f(std::integral_constant<int, 16>{}, std::integral_constant<int, 1>{});
    case 128: return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 1>{});
    case 256: return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 2>{});
    default: return PERT_ERR_UNSUPPORTED;
  }
}

}  // namespace

int pert_segreduce_stream(const float* msg, const int* rowptr, float* out, long long N, int H, int op,
                          cudaStream_t st);

extern "C" {

// op: 0 = sum, 1 = max.  perm may be null (msg already in CSR order).
int pert_segment_reduce_fwd(const float* msg, const int* rowptr, const int* perm, float* out, long long N, int H,
                            int op, void* stream) {
  if (N < 0 || H <= 0 || !rowptr || !out || (op != 0 && op != 1)) return PERT_ERR_BADARG;
  if (N == 0) return PERT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = PERT_ERR_UNSUPPORTED;
  if (!perm) {   // CSR-ordered messages: TMA-pipelined streaming kernel (segreduce_tma.cu)
    rc = pert_segreduce_stream(msg, rowptr, out, N, H, op, st);
    if (rc == PERT_OK) {
      PERT_LAUNCH_CHECK();
      return PERT_OK;
    }
    if (rc != PERT_ERR_UNSUPPORTED) return rc;
  }
  if ((((uintptr_t)msg | (uintptr_t)out) & 15) == 0) {
    rc = dispatch_w(H, [&](auto lpr, auto vpl) {
      constexpr int LPR = decltype(lpr)::value, VPL = decltype(vpl)::value;
      const int gpw = 32 / LPR;
      const long long warps = (N + gpw - 1) / gpw;
      const int grid = pert_cdiv(warps * 32, 256);
      if (op == 1) {
        if (perm) k_segreduce<LPR, VPL, true, true><<<grid, 256, 0, st>>>(msg, rowptr, perm, out, (int)N);
        else k_segreduce<LPR, VPL, true, false><<<grid, 256, 0, st>>>(msg, rowptr, perm, out, (int)N);
      } else {
        if (perm) k_segreduce<LPR, VPL, false, true><<<grid, 256, 0, st>>>(msg, rowptr, perm, out, (int)N);
        else k_segreduce<LPR, VPL, false, false><<<grid, 256, 0, st>>>(msg, rowptr, perm, out, (int)N);
      }
      return PERT_OK;
    });
  }
  if (rc == PERT_ERR_UNSUPPORTED) {
    long long total = N * (long long)H;
    int grid = pert_cdiv(total, 256);
    if (op == 1) k_segreduce_generic<true><<<grid, 256, 0, st>>>(msg, rowptr, perm, out, total, H);
    else k_segreduce_generic<false><<<grid, 256, 0, st>>>(msg, rowptr, perm, out, total, H);
    rc = PERT_OK;
  }
  if (rc) return rc;
  PERT_LAUNCH_CHECK();
  return PERT_OK;
}

// dmsg must hold one row per edge; every row belongs to exactly one segment so all of dmsg is written.
int pert_segment_reduce_bwd(const float* dout, const float* msg, const float* out, const int* rowptr,
                            const int* perm, float* dmsg, long long N, int H, int op, void* stream) {
  if (N < 0 || H <= 0 || !rowptr || !dout || !dmsg || (op != 0 && op != 1)) return PERT_ERR_BADARG;
  if (op == 1 && (!msg || !out)) return PERT_ERR_BADARG;
  if (N == 0) return PERT_OK;
  long long total = N * (long long)H;
  int grid = pert_cdiv(total, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (op == 1) k_segreduce_bwd<true><<<grid, 256, 0, st>>>(dout, msg, out, rowptr, perm, dmsg, total, H);
  else k_segreduce_bwd<false><<<grid, 256, 0, st>>>(dout, msg, out, rowptr, perm, dmsg, total, H);
  PERT_LAUNCH_CHECK();
  return PERT_OK;
}

}  // extern "C"
