// Shared-memory-staged variant of the fused TransformerConv kernels (see tconv.cu for the math).
//
// Why: the per-row gather kernels are bound by L2->SM traffic, latency and instruction issue, not by HBM: every
// k/v row is fetched once per out-edge (~3x), L1 catches only ~35-50 % of that, and the predicated 4-edge register
// blocks cost 19 M instructions per launch at cfg2 (r1 ncu: 70 % of issue cycles without an eligible warp, DRAM at
// 10 % with bytes == algorithmic bytes).  Batched call graphs are graph-major: all neighbours of a node live within a
// few hundred rows of it.  So a CTA takes a TILE of consecutive nodes, pulls the tile's two operand planes (k and v
// forward / target pass, g and q in the source pass -- contiguous in the plane layout) into shared memory with two
// TMA bulk copies (cp.async.bulk + mbarrier complete_tx: no per-thread loads, no register staging), stages the
// tile's CSR (or CSC) index slices and per-edge scalars next to them, and walks each node's edges one by one
// against shared memory (~30-cycle latency, 32-bit addressing, no padding slots), with the one remaining global
// gather (the interface-table row) prefetched one edge ahead.  Each operand row is read from L2/HBM once per tile;
// neighbours outside the tile (a graph cut by a tile boundary, graphs larger than a tile) fall back to the global
// gather.  Same semantics, same outputs as tconv.cu.
//
// Tile size (nodes) is chosen by the host so that two CTAs of 512 threads fit one SM; when the batch holds equally
// sized graphs the tile is a whole number of graphs (no cut edges).
#include "common.cuh"
#include <type_traits>

namespace {

constexpr int TILE_THREADS = 512;

// ---------------------------------------------------------------- mbarrier / bulk-copy PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
  } while (!ok);
}
// global -> shared bulk copy (TMA, 1-D): dst/src 16-byte aligned, bytes % 16 == 0
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

struct TileArgs {
  const float *q, *k, *v, *s;   // planes, dense rows (ld == H)
  const float* g;               // backward: dL/dout, dense rows
  const int *rowptr, *csr_src, *csr_if, *csr_rpc;
  const int *colptr, *csc_pos, *csc_dst;
  const float *t_if, *t_rpc;
  int n_rpc;
  float* out;       // fwd: out rows; bwd_dst: dq rows (dense)
  float *dk, *dv;   // bwd_src
  float* alpha;     // fwd: written; bwd: read
  float* dsp;       // bwd_dst: written; bwd_src: read
  float *dt_if, *dt_rpc;
  double* bn_acc;   // fwd: optional [2][H] column sums / sums of squares of `out` (BatchNorm statistics), +=
  float* rpc_ws;    // [N][2][RPC_FAST] per-target sums over in-edges of (alpha, ds) by rpc type, or null (see bwd_src)
  int N, tile_nodes, edge_cap;
  float inv_sqrt_c;
};

// Gradient of the (tiny, hot) rpc-type table without per-edge atomics: for an edge t -> i of rpc type b the table row
// receives de = alpha_t g_i + ds_t q_i, so  dT_rpc[b] = sum_i (A_ib g_i + S_ib q_i)  with the per-target scalars
// A_ib = sum_{t in in(i), rpc(t)=b} alpha_t and S_ib likewise over ds_t.  The target pass accumulates A, S in the lanes
// (lane b of the node's group owns type b) and writes 2 x RPC_FAST floats per node; the source pass -- which has the
// g and q tiles in shared memory anyway -- finishes with a [RPC_FAST x T] x [T x H] product per tile.  With the
// per-edge shared-memory atomics the source pass took 52 us at the BASELINE cfg2 shape, without them 28 us.
constexpr int RPC_FAST = 8;   // fast path for n_rpc <= 8 (the reference data has a handful of rpc types)

// dynamic smem layout shared by the three kernels (NA per-edge 4-byte arrays, kernel specific):
//   [bar 16 B][tile A: T*H][tile B: T*H][rpc table (source pass only): n_rpc*H][row/col ptr slice: T+1][NA x ecap]
// attribute ids are staged packed: interface id | rpc id << 22
#define PACK_ID(a, b) ((a) | ((b) << 22))
#define ID_IF(x) ((x) & 0x3fffff)
#define ID_RPC(x) ((int)((unsigned)(x) >> 22))
struct Smem {
  uint64_t* bar;
  float *ta, *tb, *rpc;
  int *ptr, *e0, *e1;
  float *f0, *f1;
};
__device__ __forceinline__ Smem carve_smem(unsigned char* base, int T, int H, int n_rpc, int ecap) {
  Smem s;
  s.bar = reinterpret_cast<uint64_t*>(base);
  float* f = reinterpret_cast<float*>(base + 16);
  s.ta = f; f += (size_t)T * H;
  s.tb = f; f += (size_t)T * H;
  s.rpc = f; f += (size_t)n_rpc * H;
  s.ptr = reinterpret_cast<int*>(f); f += ((T + 1 + 3) / 4) * 4;
  s.e0 = reinterpret_cast<int*>(f); f += ecap;
  s.e1 = reinterpret_cast<int*>(f); f += ecap;
  s.f0 = f; f += ecap;       // only the first NA arrays are backed by memory (see smem_bytes)
  s.f1 = f;
  return s;
}
static size_t smem_bytes(int T, int H, int n_rpc, int ecap, int n_edge_arrays) {
  return 16 + sizeof(float) * ((size_t)2 * T * H + (size_t)n_rpc * H + ((T + 1 + 3) / 4) * 4 +
                               (size_t)n_edge_arrays * ecap);
}

// stage the two operand tiles + the node-pointer slice; returns after the pointers are visible (tiles: mbar_wait)
template <int H>
__device__ __forceinline__ void stage_tiles(const Smem& S, const float* pa, const float* pb, const int* nodeptr,
                                            int n0, int nt, int tid) {
  if (tid == 0) {
    mbar_init(S.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t tile_bytes = (uint32_t)nt * H * 4;
    mbar_expect_tx(S.bar, 2 * tile_bytes);
    bulk_g2s(S.ta, pa + (size_t)n0 * H, tile_bytes, S.bar);
    bulk_g2s(S.tb, pb + (size_t)n0 * H, tile_bytes, S.bar);
  }
  for (int x = tid; x <= nt; x += TILE_THREADS) S.ptr[x] = __ldg(nodeptr + n0 + x);
  __syncthreads();
}

// ---------------------------------------------------------------- explicit shared-space accessors (32-bit addresses:
// keeps the compiler from falling back to generic loads + cluster-address checks inside the edge loops)
__device__ __forceinline__ float4 lds4s(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ int ldsi(uint32_t a) {
  int v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float ldsf(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void stsf(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
template <int LPR>
__device__ __forceinline__ float gsum_full(float v) {   // butterfly inside each LPR-lane group, whole warp converged
#pragma unroll
  for (int off = LPR >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
struct SAddr {   // shared-space byte addresses of the staged arrays
  uint32_t ta, tb, rpc, ptr, e0, e1, f0, f1;
};
__device__ __forceinline__ SAddr saddr_of(const Smem& S) {
  SAddr s;
  s.ta = smem_u32(S.ta); s.tb = smem_u32(S.tb); s.rpc = smem_u32(S.rpc); s.ptr = smem_u32(S.ptr);
  s.e0 = smem_u32(S.e0); s.e1 = smem_u32(S.e1); s.f0 = smem_u32(S.f0); s.f1 = smem_u32(S.f1);
  return s;
}

// All lane groups of a warp walk their nodes' edges in lockstep (trip count = the warp's max degree, finished groups
// contribute zeros), so shuffles use the full mask and no per-group branch divergence bookkeeping is generated.

// ============================================================== forward
template <int LPR, bool HAS_E>
__global__ void __launch_bounds__(TILE_THREADS, 2) k_tile_fwd(TileArgs a) {
  constexpr int H = 4 * LPR;
  constexpr int GPW = 32 / LPR;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int T = a.tile_nodes;
  const Smem S = carve_smem(smem_raw, T, H, 0, a.edge_cap);
  const int n0 = blockIdx.x * T;
  const int nt = min(T, a.N - n0);
  const int tid = threadIdx.x;
  stage_tiles<H>(S, a.k, a.v, a.rowptr, n0, nt, tid);
  const int e_lo = S.ptr[0];
  const int ne_s = min(S.ptr[nt] - e_lo, a.edge_cap);
  int bad = 0;   // a staged neighbour outside this tile
  for (int x = tid; x < ne_s; x += TILE_THREADS) {
    const int nb_id = __ldg(a.csr_src + e_lo + x);
    S.e0[x] = nb_id;
    bad |= (unsigned)(nb_id - n0) >= (unsigned)nt;
    if (HAS_E) S.e1[x] = PACK_ID(__ldg(a.csr_if + e_lo + x), __ldg(a.csr_rpc + e_lo + x));
  }
  // fast variant of the edge loops when every edge of the tile is staged and every neighbour row is in the tile
  // (whole-graph tiles): no global-memory fallbacks are compiled into it
  const bool all_in = (__syncthreads_or(bad) == 0) && (S.ptr[nt] - e_lo <= a.edge_cap);
  mbar_wait(S.bar, 0);

  const SAddr sa = saddr_of(S);
  const int lane = tid & 31, lig = lane % LPR, grp = lane / LPR;
  constexpr int GPC = TILE_THREADS / LPR;  // lane groups per CTA
  const float qscale = a.inv_sqrt_c * 1.4426950408889634f;   // logits kept in log2 units: exp(x) = 2^(x*log2 e)
  const uint32_t lane4 = lig * 16;
  const int g0 = (tid >> 5) * GPW;          // first group of this warp
  // q / skip rows of the next node are requested one node ahead
  int loc = g0 + grp;
  float4 q_n = f4zero(), s_n = f4zero();
  if (loc < nt) {
    q_n = ldg4(a.q + (size_t)(n0 + loc) * H + lig * 4);
    if (a.s) s_n = ldg4(a.s + (size_t)(n0 + loc) * H + lig * 4);
  }
  float4 bsum = f4zero(), bsq = f4zero();   // this lane's 4 columns over its nodes (fused BatchNorm statistics)
  auto run = [&](auto fast_c) {
    constexpr bool FAST = decltype(fast_c)::value;
    for (; g0 + (loc - g0 - grp) < nt; loc += GPC) {   // warp-uniform: the warp's first group still has a node
      const bool valid = loc < nt;
      const int i = n0 + loc;
      const float4 q = f4scale(qscale, q_n);
      const float4 skip = s_n;
      if (loc + GPC < nt) {
        q_n = ldg4(a.q + (size_t)(i + GPC) * H + lig * 4);
        if (a.s) s_n = ldg4(a.s + (size_t)(i + GPC) * H + lig * 4);
      }
      const int p0 = valid ? ldsi(sa.ptr + loc * 4) : 0;
      const int p1 = valid ? ldsi(sa.ptr + loc * 4 + 4) : 0;
      const int deg = p1 - p0;
      const int degmax = __reduce_max_sync(0xffffffffu, deg);
      float4 acc = f4zero();
      float m = -INFINITY, Z = 0.f;
      // software pipeline over edges: ids + table rows of edge t+1 are requested before edge t is consumed
      int j = n0, id = 0;
      float4 eif = f4zero(), erp = f4zero();
      auto fetch = [&](int p, bool on) {
        j = n0;
        id = 0;
        if (on) {
          const int le = p - e_lo;
          if (FAST || le < ne_s) {
            j = ldsi(sa.e0 + le * 4);
            if (HAS_E) id = ldsi(sa.e1 + le * 4);
          } else {
            j = __ldg(a.csr_src + p);
            if (HAS_E) id = PACK_ID(__ldg(a.csr_if + p), __ldg(a.csr_rpc + p));
          }
        }
        if (HAS_E) {
          eif = ldg4(a.t_if + (size_t)ID_IF(id) * H + lig * 4);
          erp = ldg4(a.t_rpc + ID_RPC(id) * H + lig * 4);
        }
      };
      fetch(p0, 0 < deg);
      for (int t = 0; t < degmax; ++t) {
        const bool on = t < deg;
        const int p = p0 + t;
        const int cj = j;
        const float4 e = f4add(eif, erp);
        fetch(p + 1, t + 1 < deg);
        float4 kk, vv;
        const unsigned sl = (unsigned)(cj - n0);
        if (FAST || sl < (unsigned)nt) {
          kk = lds4s(sa.ta + sl * (H * 4) + lane4);
          vv = lds4s(sa.tb + sl * (H * 4) + lane4);
        } else {
          kk = ldg4(a.k + (size_t)cj * H + lig * 4);
          vv = ldg4(a.v + (size_t)cj * H + lig * 4);
        }
        if (HAS_E) {
          kk = f4add(kk, e);
          vv = f4add(vv, e);
        }
        const float s = gsum_full<LPR>(f4dot(q, kk));
        if (on && lig == 0) {
          const int le = p - e_lo;
          if (FAST || le < ne_s) stsf(sa.f0 + le * 4, s);
          else a.alpha[p] = s;
        }
        const float mn = on ? fmaxf(m, s) : m;
        const float sc = on ? ex2(m - mn) : 1.f;
        const float pz = on ? ex2(s - mn) : 0.f;
        Z = fmaf(Z, sc, pz);
        acc.x = fmaf(pz, vv.x, acc.x * sc);
        acc.y = fmaf(pz, vv.y, acc.y * sc);
        acc.z = fmaf(pz, vv.z, acc.z * sc);
        acc.w = fmaf(pz, vv.w, acc.w * sc);
        m = mn;
      }
      const float invZ = 1.0f / (Z + 1e-16f);
      if (valid) {
        const float4 o = f4add(f4scale(invZ, acc), skip);
        st4(a.out + (size_t)i * H + lig * 4, o);
        bsum = f4add(bsum, o);
        bsq.x = fmaf(o.x, o.x, bsq.x); bsq.y = fmaf(o.y, o.y, bsq.y);
        bsq.z = fmaf(o.z, o.z, bsq.z); bsq.w = fmaf(o.w, o.w, bsq.w);
      }
      __syncwarp();
      for (int p = p0 + lig; p < p1; p += LPR) {
        const int le = p - e_lo;
        const float s = (FAST || le < ne_s) ? ldsf(sa.f0 + le * 4) : a.alpha[p];
        a.alpha[p] = ex2(s - m) * invZ;
      }
    }
  };
  if (all_in) run(std::true_type{});
  else run(std::false_type{});
  if (a.bn_acc) {
    // BatchNorm statistics of this layer's output, fused: column sums and sums of squares of the tile (per-lane fp32
    // partials over <= a handful of nodes, combined in fp64) -> one fp64 atomic per column and tile into bn_acc, the
    // accumulator k_bn_apply derives mean / rstd from (nodeops.cu).  Saves the separate statistics pass over `out`.
    __syncthreads();                                   // every warp is done with the staged tiles
    double* sc = reinterpret_cast<double*>(S.ta);
    for (int x = tid; x < 2 * H; x += TILE_THREADS) sc[x] = 0.0;
    __syncthreads();
    float vals[8] = {bsum.x, bsum.y, bsum.z, bsum.w, bsq.x, bsq.y, bsq.z, bsq.w};
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      float v = vals[kk];
#pragma unroll
      for (int off = LPR; off < 32; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);   // the warp's groups
      if (grp == 0) atomicAdd(&sc[(kk < 4 ? 0 : H) + lig * 4 + (kk & 3)], (double)v);
    }
    __syncthreads();
    for (int x = tid; x < 2 * H; x += TILE_THREADS) atomicAdd(a.bn_acc + x, sc[x]);
  }
}

// ============================================================== backward, target pass (dq, ds)
template <int LPR, bool HAS_E>
__global__ void __launch_bounds__(TILE_THREADS, 2) k_tile_bwd_dst(TileArgs a) {
  constexpr int H = 4 * LPR;
  constexpr int GPW = 32 / LPR;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int T = a.tile_nodes;
  const Smem S = carve_smem(smem_raw, T, H, 0, a.edge_cap);
  const int n0 = blockIdx.x * T;
  const int nt = min(T, a.N - n0);
  const int tid = threadIdx.x;
  stage_tiles<H>(S, a.k, a.v, a.rowptr, n0, nt, tid);
  const int e_lo = S.ptr[0];
  const int ne_s = min(S.ptr[nt] - e_lo, a.edge_cap);
  int bad = 0;   // a staged neighbour outside this tile
  for (int x = tid; x < ne_s; x += TILE_THREADS) {
    const int nb_id = __ldg(a.csr_src + e_lo + x);
    S.e0[x] = nb_id;
    bad |= (unsigned)(nb_id - n0) >= (unsigned)nt;
    S.f0[x] = __ldg(a.alpha + e_lo + x);
    if (HAS_E) S.e1[x] = PACK_ID(__ldg(a.csr_if + e_lo + x), __ldg(a.csr_rpc + e_lo + x));
  }
  // fast variant of the edge loops when every edge of the tile is staged and every neighbour row is in the tile
  // (whole-graph tiles): no global-memory fallbacks are compiled into it
  const bool all_in = (__syncthreads_or(bad) == 0) && (S.ptr[nt] - e_lo <= a.edge_cap);
  mbar_wait(S.bar, 0);

  const SAddr sa = saddr_of(S);
  const int lane = tid & 31, lig = lane % LPR, grp = lane / LPR;
  constexpr int GPC = TILE_THREADS / LPR;
  const uint32_t lane4 = lig * 16;
  const int g0 = (tid >> 5) * GPW;
  int loc = g0 + grp;
  float4 g_n = f4zero();
  if (loc < nt) g_n = ldg4(a.g + (size_t)(n0 + loc) * H + lig * 4);
  auto run = [&](auto fast_c) {
    constexpr bool FAST = decltype(fast_c)::value;
    for (; g0 + (loc - g0 - grp) < nt; loc += GPC) {
      const bool valid = loc < nt;
      const int i = n0 + loc;
      const float4 g = g_n;
      if (loc + GPC < nt) g_n = ldg4(a.g + (size_t)(i + GPC) * H + lig * 4);
      const int p0 = valid ? ldsi(sa.ptr + loc * 4) : 0;
      const int p1 = valid ? ldsi(sa.ptr + loc * 4 + 4) : 0;
      const int deg = p1 - p0;
      const int degmax = __reduce_max_sync(0xffffffffu, deg);
      // one edge record: source row offsets + e = T_if[a] + T_rpc[b]
      auto edge = [&](int p, bool on, int& j, float& al, float4& e, int& rid) {
        j = n0;
        al = 0.f;
        int id = 0;
        if (on) {
          const int le = p - e_lo;
          if (FAST || le < ne_s) {
            j = ldsi(sa.e0 + le * 4);
            al = ldsf(sa.f0 + le * 4);
            if (HAS_E) id = ldsi(sa.e1 + le * 4);
          } else {
            j = __ldg(a.csr_src + p);
            al = __ldg(a.alpha + p);
            if (HAS_E) id = PACK_ID(__ldg(a.csr_if + p), __ldg(a.csr_rpc + p));
          }
        }
        e = f4zero();
        if (HAS_E)
          e = f4add(ldg4(a.t_if + (size_t)ID_IF(id) * H + lig * 4), ldg4(a.t_rpc + ID_RPC(id) * H + lig * 4));
        rid = ID_RPC(id);
      };
      // ONE pass over the in-edges.  With d_t = <g_i, v_j + e_t> (shifted by the first edge's value c, which cancels
      // exactly: sum_t ds_t = 0), w_t = alpha_t (d_t - c) and dot = sum_t w_t:
      //   ds_t = alpha_t (d_t - c - dot) / sqrt(C)
      //   dq_i = sum_t ds_t (k_j + e_t) = (P - dot Q) / sqrt(C),   P = sum_t w_t (k_j + e_t),  Q = sum_t alpha_t (k_j + e_t)
      // so k_j, v_j and the table rows of an edge are fetched once (the two-pass form fetched the edge record and
      // its table rows twice); ds_t is written by a scalar post-pass with the lanes spread over the node's edges.
      float dot = 0.f, c_shift = 0.f;
      float4 P = f4zero(), Q = f4zero();
      float sumA = 0.f, sumW = 0.f;          // lane b: sums of alpha / w over this node's in-edges of rpc type b
      for (int t = 0; t < degmax; ++t) {
        const bool on = t < deg;
        const int p = p0 + t;
        int j, rid;
        float al;
        float4 e;
        edge(p, on, j, al, e, rid);
        float4 kk, vv;
        const unsigned sl = (unsigned)(j - n0);
        if (FAST || sl < (unsigned)nt) {
          kk = lds4s(sa.ta + sl * (H * 4) + lane4);
          vv = lds4s(sa.tb + sl * (H * 4) + lane4);
        } else {
          kk = ldg4(a.k + (size_t)j * H + lig * 4);
          vv = ldg4(a.v + (size_t)j * H + lig * 4);
        }
        const float da = gsum_full<LPR>(f4dot(g, f4add(vv, e)));
        if (t == 0) c_shift = da;
        const float dc = da - c_shift;
        const float w = al * dc;             // alpha is 0 on finished groups
        dot += w;
        const float4 ke = f4add(kk, e);
        P = f4fma(w, ke, P);
        Q = f4fma(al, ke, Q);
        if (HAS_E && on && lig == rid) { sumA += al; sumW += w; }
        if (on && lig == 0) {
          const int le = p - e_lo;
          if (FAST || le < ne_s) stsf(sa.f1 + le * 4, dc);
          else a.dsp[p] = dc;
        }
      }
      float4 dq;
      dq.x = (P.x - dot * Q.x) * a.inv_sqrt_c; dq.y = (P.y - dot * Q.y) * a.inv_sqrt_c;
      dq.z = (P.z - dot * Q.z) * a.inv_sqrt_c; dq.w = (P.w - dot * Q.w) * a.inv_sqrt_c;
      const float sumS = (sumW - dot * sumA) * a.inv_sqrt_c;
      __syncwarp();
      for (int p = p0 + lig; p < p1; p += LPR) {
        const int le = p - e_lo;
        const float dc = (FAST || le < ne_s) ? ldsf(sa.f1 + le * 4) : a.dsp[p];
        const float al = (FAST || le < ne_s) ? ldsf(sa.f0 + le * 4) : __ldg(a.alpha + p);
        a.dsp[p] = al * (dc - dot) * a.inv_sqrt_c;
      }
      if (valid) st4(a.out + (size_t)i * H + lig * 4, dq);
      if (HAS_E && a.rpc_ws && valid && lig < RPC_FAST) {
        a.rpc_ws[(size_t)i * 2 * RPC_FAST + lig] = sumA;
        a.rpc_ws[(size_t)i * 2 * RPC_FAST + RPC_FAST + lig] = sumS;
      }
    }
  };
  if (all_in) run(std::true_type{});
  else run(std::false_type{});
}

// ============================================================== backward, source pass (dk, dv, table grads)
template <int LPR, bool HAS_E>
__global__ void __launch_bounds__(TILE_THREADS, 2) k_tile_bwd_src(TileArgs a) {
  constexpr int H = 4 * LPR;
  constexpr int GPW = 32 / LPR;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int T = a.tile_nodes;
  const Smem S = carve_smem(smem_raw, T, H, HAS_E ? a.n_rpc : 0, a.edge_cap);
  const int n0 = blockIdx.x * T;
  const int nt = min(T, a.N - n0);
  const int tid = threadIdx.x;
  float* s_drpc = S.rpc;   // privatised gradient of the rpc-type table (few hot rows), flushed once per CTA
  if (HAS_E)
    for (int x = tid; x < a.n_rpc * H; x += TILE_THREADS) s_drpc[x] = 0.f;
  stage_tiles<H>(S, a.g, a.q, a.colptr, n0, nt, tid);   // targets of a node's out-edges live in the same graph
  const int c_lo = S.ptr[0];
  const int ne_s = min(S.ptr[nt] - c_lo, a.edge_cap);
  // per out-edge (CSC order): target, and through the CSR slot its alpha, ds and attribute ids
  int bad = 0;   // a staged neighbour outside this tile
  for (int x = tid; x < ne_s; x += TILE_THREADS) {
    const int p = __ldg(a.csc_pos + c_lo + x);
    const int nb_id = __ldg(a.csc_dst + c_lo + x);
    S.e0[x] = nb_id;
    bad |= (unsigned)(nb_id - n0) >= (unsigned)nt;
    S.f0[x] = __ldg(a.alpha + p);
    S.f1[x] = __ldg(a.dsp + p);
    if (HAS_E) S.e1[x] = PACK_ID(__ldg(a.csr_if + p), __ldg(a.csr_rpc + p));
  }
  // fast variant of the edge loops when every edge of the tile is staged and every neighbour row is in the tile
  // (whole-graph tiles): no global-memory fallbacks are compiled into it
  const bool all_in = (__syncthreads_or(bad) == 0) && (S.ptr[nt] - c_lo <= a.edge_cap);
  mbar_wait(S.bar, 0);

  const SAddr sa = saddr_of(S);
  const int lane = tid & 31, lig = lane % LPR, grp = lane / LPR;
  constexpr int GPC = TILE_THREADS / LPR;
  const uint32_t lane4 = lig * 16;
  const int g0 = (tid >> 5) * GPW;
  auto run = [&](auto fast_c) {
    constexpr bool FAST = decltype(fast_c)::value;
    for (int loc = g0 + grp; g0 + (loc - g0 - grp) < nt; loc += GPC) {
      const bool valid = loc < nt;
      const int jn = n0 + loc;
      const int c0 = valid ? ldsi(sa.ptr + loc * 4) : 0;
      const int c1 = valid ? ldsi(sa.ptr + loc * 4 + 4) : 0;
      const int deg = c1 - c0;
      const int degmax = __reduce_max_sync(0xffffffffu, deg);
      float4 dk = f4zero(), dv = f4zero();
      for (int t = 0; t < degmax; ++t) {
        const bool on = t < deg;
        const int c = c0 + t;
        int i = n0, id = 0;
        float al = 0.f, ds = 0.f;
        if (on) {
          const int le = c - c_lo;
          if (FAST || le < ne_s) {
            i = ldsi(sa.e0 + le * 4); al = ldsf(sa.f0 + le * 4); ds = ldsf(sa.f1 + le * 4);
            if (HAS_E) id = ldsi(sa.e1 + le * 4);
          } else {
            const int p = __ldg(a.csc_pos + c);
            i = __ldg(a.csc_dst + c); al = __ldg(a.alpha + p); ds = __ldg(a.dsp + p);
            if (HAS_E) id = PACK_ID(__ldg(a.csr_if + p), __ldg(a.csr_rpc + p));
          }
        }
        float4 gi, qi;
        const unsigned sl = (unsigned)(i - n0);
        if (FAST || sl < (unsigned)nt) {
          gi = lds4s(sa.ta + sl * (H * 4) + lane4);
          qi = lds4s(sa.tb + sl * (H * 4) + lane4);
        } else {
          gi = ldg4(a.g + (size_t)i * H + lig * 4);
          qi = ldg4(a.q + (size_t)i * H + lig * 4);
        }
        dk = f4fma(ds, qi, dk);
        dv = f4fma(al, gi, dv);
        if (HAS_E && on) {
          const float4 de = f4fma(ds, qi, f4scale(al, gi));
          red4(a.dt_if + (size_t)ID_IF(id) * H + lig * 4, de);
          if (!a.rpc_ws) {                 // general path (n_rpc > RPC_FAST): privatised table, per-edge atomics
            float* prp = s_drpc + ID_RPC(id) * H + lig * 4;
            atomicAdd(prp + 0, de.x);
            atomicAdd(prp + 1, de.y);
            atomicAdd(prp + 2, de.z);
            atomicAdd(prp + 3, de.w);
          }
        }
      }
      if (valid) {
        st4(a.dk + (size_t)jn * H + lig * 4, dk);
        st4(a.dv + (size_t)jn * H + lig * 4, dv);
      }
    }
  };
  if (all_in) run(std::true_type{});
  else run(std::false_type{});
  if (HAS_E && a.rpc_ws) {
    // dT_rpc tile contribution: thread = (column, node slice); the per-target scalars are warp-uniform loads
    constexpr int NSL = TILE_THREADS / H;
    const int col = tid % H, slc = tid / H;
    float acc[RPC_FAST];
#pragma unroll
    for (int b = 0; b < RPC_FAST; ++b) acc[b] = 0.f;
#pragma unroll 2
    for (int loc = slc; loc < nt; loc += NSL) {
      const float4* wp = reinterpret_cast<const float4*>(a.rpc_ws + (size_t)(n0 + loc) * 2 * RPC_FAST);
      const float4 A0 = __ldg(wp), A1 = __ldg(wp + 1), S0 = __ldg(wp + 2), S1 = __ldg(wp + 3);
      const float gv = S.ta[(size_t)loc * H + col], qv = S.tb[(size_t)loc * H + col];
      acc[0] = fmaf(A0.x, gv, fmaf(S0.x, qv, acc[0])); acc[1] = fmaf(A0.y, gv, fmaf(S0.y, qv, acc[1]));
      acc[2] = fmaf(A0.z, gv, fmaf(S0.z, qv, acc[2])); acc[3] = fmaf(A0.w, gv, fmaf(S0.w, qv, acc[3]));
      acc[4] = fmaf(A1.x, gv, fmaf(S1.x, qv, acc[4])); acc[5] = fmaf(A1.y, gv, fmaf(S1.y, qv, acc[5]));
      acc[6] = fmaf(A1.z, gv, fmaf(S1.z, qv, acc[6])); acc[7] = fmaf(A1.w, gv, fmaf(S1.w, qv, acc[7]));
    }
#pragma unroll
    for (int b = 0; b < RPC_FAST; ++b)
      if (b < a.n_rpc && acc[b] != 0.f) atomicAdd(s_drpc + b * H + col, acc[b]);
  }
  if (HAS_E) {
    __syncthreads();
    for (int x = tid; x < a.n_rpc * H; x += TILE_THREADS) {
      const float v = s_drpc[x];
      if (v != 0.f) atomicAdd(a.dt_rpc + x, v);
    }
  }
}

// tile geometry for a launch: nodes per tile and staged-edge capacity within the per-CTA smem budget
// (two CTAs per SM: (233472 / 2) - 1024 reserved).  When the batch holds B equally sized graphs the tile is a
// whole number of graphs, so no edge ever crosses a tile boundary.
struct TileGeom {
  int T, ecap;
  size_t bytes;
};
TileGeom tile_geom(int H, int n_rpc, long long N, long long E, long long B, int n_edge_arrays) {
  const double budget2 = 115712.0, budget1 = 231424.0;
  const double deg = N > 0 ? (double)E / (double)N : 1.0;
  const double per_node = 4.0 * (2.0 * H + 1.0 + n_edge_arrays * deg * 1.02);
  const double fixed = 64.0 + 4.0 * n_rpc * H + 4.0 * n_edge_arrays * 20.0;
  auto fit = [&](double b) { return (long long)((b - fixed) / per_node); };
  long long T = fit(budget2);
  const long long G = (B > 0 && N % B == 0) ? N / B : 0;   // uniform graph size, if any
  if ((G > 0 && T < G) || T < 64) T = fit(budget1);        // wide rows / big graphs: one CTA per SM
  if (G > 0 && G <= T) T = T / G * G;
  if (T > N) T = N;
  if (T < 1) T = 1;
  int ecap = (int)(deg * 1.02 * (double)T) + 20;
  ecap = (ecap + 3) / 4 * 4;
  TileGeom g{(int)T, ecap, smem_bytes((int)T, H, n_rpc, ecap, n_edge_arrays)};
  return g;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return e == cudaSuccess ? 0 : (int)e;
}

template <int LPR>
int launch_fwd(const TileArgs& a0, long long N, long long E, long long B, bool has_e, cudaStream_t st) {
  constexpr int H = 4 * LPR;
  const TileGeom g = tile_geom(H, 0, N, E, B, 3);   // src, packed ids, logit staging
  TileArgs a = a0;
  a.tile_nodes = g.T;
  a.edge_cap = g.ecap;
  const int grid = pert_cdiv(N, g.T);
  int rc;
  if (has_e) {
    if ((rc = set_smem(k_tile_fwd<LPR, true>, g.bytes))) return rc;
    k_tile_fwd<LPR, true><<<grid, TILE_THREADS, g.bytes, st>>>(a);
  } else {
    if ((rc = set_smem(k_tile_fwd<LPR, false>, g.bytes))) return rc;
    k_tile_fwd<LPR, false><<<grid, TILE_THREADS, g.bytes, st>>>(a);
  }
  return PERT_OK;
}
template <int LPR>
int launch_bwd(const TileArgs& a0, long long N, long long E, long long B, bool has_e, cudaStream_t st) {
  constexpr int H = 4 * LPR;
  const TileGeom gd = tile_geom(H, 0, N, E, B, 4);                      // src, ids, alpha, dalpha staging
  const TileGeom gs = tile_geom(H, has_e ? a0.n_rpc : 0, N, E, B, 4);   // dst, ids, alpha, ds
  TileArgs ad = a0, as = a0;
  ad.tile_nodes = gd.T; ad.edge_cap = gd.ecap;
  as.tile_nodes = gs.T; as.edge_cap = gs.ecap;
  int rc;
  // per-target rpc sums (see RPC_FAST) live in caller scratch; without it the general atomics path runs
  float* ws = (has_e && a0.n_rpc <= RPC_FAST && LPR >= RPC_FAST && TILE_THREADS % H == 0) ? a0.rpc_ws : nullptr;
  ad.rpc_ws = as.rpc_ws = ws;
  if (has_e) {
    if ((rc = set_smem(k_tile_bwd_dst<LPR, true>, gd.bytes))) return rc;
    if ((rc = set_smem(k_tile_bwd_src<LPR, true>, gs.bytes))) return rc;
    k_tile_bwd_dst<LPR, true><<<pert_cdiv(N, gd.T), TILE_THREADS, gd.bytes, st>>>(ad);
    k_tile_bwd_src<LPR, true><<<pert_cdiv(N, gs.T), TILE_THREADS, gs.bytes, st>>>(as);
  } else {
    if ((rc = set_smem(k_tile_bwd_dst<LPR, false>, gd.bytes))) return rc;
    if ((rc = set_smem(k_tile_bwd_src<LPR, false>, gs.bytes))) return rc;
    k_tile_bwd_dst<LPR, false><<<pert_cdiv(N, gd.T), TILE_THREADS, gd.bytes, st>>>(ad);
    k_tile_bwd_src<LPR, false><<<pert_cdiv(N, gs.T), TILE_THREADS, gs.bytes, st>>>(as);
  }
  return PERT_OK;
}

}  // namespace

// Entry points used by tconv.cu's C-ABI functions: return PERT_ERR_UNSUPPORTED when the tile path does not apply
// (width, strided planes, oversized rpc table) so the caller falls through to the per-row gather kernels.
int pert_tile_fwd(const float* q, const float* k, const float* v, const float* s, int ld, const int* rowptr,
                  const int* csr_src, const int* csr_if, const int* csr_rpc, const float* t_if, const float* t_rpc,
                  int n_rpc, float* out, int ld_out, float* alpha, long long N, long long E, long long B, int H,
                  double* bn_acc, cudaStream_t st) {
  if (ld != H || ld_out != H || (t_if && ((size_t)n_rpc * H * 4 > 16 * 1024 || n_rpc > 1023)))
    return PERT_ERR_UNSUPPORTED;
  TileArgs a{};
  a.q = q; a.k = k; a.v = v; a.s = s;
  a.rowptr = rowptr; a.csr_src = csr_src; a.csr_if = csr_if; a.csr_rpc = csr_rpc;
  a.t_if = t_if; a.t_rpc = t_rpc; a.n_rpc = n_rpc; a.out = out; a.alpha = alpha; a.bn_acc = bn_acc;
  a.N = (int)N; a.inv_sqrt_c = 1.0f / sqrtf((float)H);
  switch (H) {
    case 32: return launch_fwd<8>(a, N, E, B, t_if != nullptr, st);
    case 64: return launch_fwd<16>(a, N, E, B, t_if != nullptr, st);
    case 128: return launch_fwd<32>(a, N, E, B, t_if != nullptr, st);
    default: return PERT_ERR_UNSUPPORTED;
  }
}

int pert_tile_bwd(const float* g_, int ld_g, const float* q, const float* k, const float* v, int ld, const int* rowptr,
                  const int* csr_src, const int* csr_if, const int* csr_rpc, const int* colptr, const int* csc_pos,
                  const int* csc_dst, const float* t_if, const float* t_rpc, const float* alpha, float* dq, float* dk,
                  float* dv, int ld_d, float* dsp, float* rpc_ws, float* dt_if, float* dt_rpc, int n_rpc, long long N,
                  long long E, long long B, int H, cudaStream_t st) {
  if (ld != H || ld_g != H || ld_d != H || (t_if && ((size_t)n_rpc * H * 4 > 16 * 1024 || n_rpc > 1023)))
    return PERT_ERR_UNSUPPORTED;
  TileArgs a{};
  a.q = q; a.k = k; a.v = v; a.g = g_;
  a.rowptr = rowptr; a.csr_src = csr_src; a.csr_if = csr_if; a.csr_rpc = csr_rpc;
  a.colptr = colptr; a.csc_pos = csc_pos; a.csc_dst = csc_dst;
  a.t_if = t_if; a.t_rpc = t_rpc; a.n_rpc = n_rpc;
  a.out = dq; a.dk = dk; a.dv = dv; a.alpha = const_cast<float*>(alpha); a.dsp = dsp;
  a.dt_if = dt_if; a.dt_rpc = dt_rpc; a.rpc_ws = rpc_ws;
  a.N = (int)N; a.inv_sqrt_c = 1.0f / sqrtf((float)H);
  switch (H) {
    case 32: return launch_bwd<8>(a, N, E, B, t_if != nullptr, st);
    case 64: return launch_bwd<16>(a, N, E, B, t_if != nullptr, st);
    case 128: return launch_bwd<32>(a, N, E, B, t_if != nullptr, st);
    default: return PERT_ERR_UNSUPPORTED;
  }
}
