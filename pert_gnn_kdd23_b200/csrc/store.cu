// Device-side batch assembly from a resident pattern store (SURVEY.md section 8f rows N1 + N4).
//
// The reference builds every sample on the host, per trace, from Python dicts and a pandas MultiIndex:
//   get_entry_data (pert_gnn.py:134-173): for the trace's entry, concatenate ALL runtime-pattern graphs of that entry --
//     get_x (:40-67, the (timestamp, msname) -> 8 resource statistics join, missing-indicator column),
//     get_cat_X (:97-99), get_node_depth (:102-104), get_edge_index (:107-119, pattern node offsets),
//     get_edge_attr (:77-82), get_pattern_num_nodes (:85-94), pattern_probs (:170) --
//   then PyG's DataLoader collates B samples (:201-209: concatenate, offset edge_index by the cumulative node count,
//   batch vector, ptr), and the train loop rebuilds the per-node pattern probability on the host every step with B
//   tiny H2D copies (:220-230, transform_pattern_probs :122-131).
// Here the patterns (CSR-like concatenation), the entry -> (pattern, probability) lists, the resource table (sorted
// (timestamp, ms) keys) and the trace table live in HBM; a batch is a list of trace ids (8 B each) and two kernels
// (one thread per output node / per output edge) write the collated Batch tensors -- bit-identical to what
// get_entry_data + Batch.from_data_list + transform_pattern_probs produce (tests/golden/ref_loop.npz holds the
// reference's own outputs).
//
// Reference quirk reproduced: get_x maps ms -> node id through a dict (ms2nid), so when a microservice occurs several
// times in one pattern only its LAST node receives the resource statistics; earlier duplicates keep [0 x 8, 1]
// (`last_occ` flag, computed when the store is built).  A resourced microservice whose (timestamp, ms) row is missing
// raises KeyError in the reference; here it sets PERT_ERR_RANGE in `status` and the node keeps the missing indicator.
#include "common.cuh"

namespace {

constexpr int NF = 8;   // resource statistics per (timestamp, ms) row; x has NF + 1 columns (pert_gnn.py:44-52)

__device__ __forceinline__ int upper_seg(const int* __restrict__ off, int n, int v) {
  // largest b in [0, n) with off[b] <= v   (off is an exclusive prefix sum, off[n] = total)
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= v) lo = mid;
    else hi = mid;
  }
  return lo;
}

// exclusive prefix sums of the per-trace node / edge / pattern counts (one CTA; B is a batch size)
__global__ void __launch_bounds__(1024) k_store_offsets(PertStore s, const int64_t* __restrict__ trace_ids, int B,
                                                        int* __restrict__ node_off, int* __restrict__ edge_off,
                                                        int* __restrict__ pat_off, int* status) {
  __shared__ int carry[3];
  __shared__ int wsum[3][32];
  if (threadIdx.x < 3) carry[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int base = 0; base < B; base += blockDim.x) {
    const int b = base + threadIdx.x;
    int v[3] = {0, 0, 0};
    if (b < B) {
      int64_t t = trace_ids[b];
      if (t < 0 || t >= s.n_traces) {
        if (status) atomicExch(status, PERT_ERR_RANGE);
        t = 0;
      }
      const int ent = s.trace_entry[t];
      v[0] = s.ent_nodes[ent];
      v[1] = s.ent_edges[ent];
      v[2] = s.ent_ptr[ent + 1] - s.ent_ptr[ent];
    }
    int incl[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      int x = v[k];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += u;
      }
      incl[k] = x;
      if (lane == 31) wsum[k][w] = x;
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        int x = wsum[k][lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int u = __shfl_up_sync(0xffffffffu, x, o);
          if (lane >= o) x += u;
        }
        wsum[k][lane] = x;
      }
    }
    __syncthreads();
    int* outs[3] = {node_off, edge_off, pat_off};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int excl = carry[k] + (w ? wsum[k][w - 1] : 0) + incl[k] - v[k];
      if (b < B) outs[k][b] = excl;
      if (b == B - 1) outs[k][B] = excl + v[k];
    }
    __syncthreads();
    if (threadIdx.x < 3) carry[threadIdx.x] += wsum[threadIdx.x][31];
    __syncthreads();
  }
  if (B == 0 && threadIdx.x == 0) node_off[0] = edge_off[0] = pat_off[0] = 0;
}

// pattern instance of local node / edge index l inside entry `ent`: returns slot k (into ent_pat / ent_prob) and the
// offsets of that pattern inside the trace
__device__ __forceinline__ int find_pattern(const PertStore& s, int ent, int l, bool edges, int& node_base,
                                            int& local) {
  int nb = 0, acc = 0;
  const int k0 = s.ent_ptr[ent], k1 = s.ent_ptr[ent + 1];
  for (int k = k0; k < k1; ++k) {
    const int p = s.ent_pat[k];
    const int nn = s.pat_nptr[p + 1] - s.pat_nptr[p];
    const int sz = edges ? s.pat_eptr[p + 1] - s.pat_eptr[p] : nn;
    if (l < acc + sz || k == k1 - 1) {
      node_base = nb;
      local = l - acc;
      return k;
    }
    acc += sz;
    nb += nn;
  }
  node_base = 0;
  local = 0;
  return k0;
}

__global__ void __launch_bounds__(256) k_store_nodes(PertStore s, const int64_t* __restrict__ trace_ids, int B,
                                                     const int* __restrict__ node_off, PertBatchOut o, int* status) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = node_off[B];
  if (n >= N) return;
  const int b = upper_seg(node_off, B, n);
  int64_t t = trace_ids[b];
  if (t < 0 || t >= s.n_traces) t = 0;
  const int ent = s.trace_entry[t];
  int node_base, local;
  const int k = find_pattern(s, ent, n - node_off[b], false, node_base, local);
  const int p = s.ent_pat[k];
  const int g = s.pat_nptr[p] + local;
  const int64_t ms = s.pat_ms[g];
  o.cat_X[n] = ms;
  o.node_depth[n] = s.pat_depth[g];
  o.pattern_num_nodes[n] = (float)(s.pat_nptr[p + 1] - s.pat_nptr[p]);
  o.rt_probs[n] = s.ent_prob[k];
  o.batch[n] = b;
  // ---- feature join (get_x): statistics of (timestamp, ms) for the LAST node of a resourced ms, else missing indicator
  float f[NF + 1];
#pragma unroll
  for (int c = 0; c < NF; ++c) f[c] = 0.f;
  f[NF] = 1.f;
  if (s.pat_last[g] && ms >= 0 && ms < s.n_ms && s.ms_has_res[ms]) {
    const int64_t key = s.trace_ts[t] * (int64_t)s.n_ms + ms;
    int lo = 0, hi = s.n_res;        // lower bound over the sorted keys
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s.res_keys[mid] < key) lo = mid + 1;
      else hi = mid;
    }
    if (lo < s.n_res && s.res_keys[lo] == key) {
#pragma unroll
      for (int c = 0; c < NF; ++c) f[c] = s.res_vals[(size_t)lo * NF + c];
      f[NF] = 0.f;
    } else if (status) {
      atomicExch(status, PERT_ERR_RANGE);      // the reference raises KeyError here (resource_df.loc)
    }
  }
#pragma unroll
  for (int c = 0; c <= NF; ++c) o.x[(size_t)n * (NF + 1) + c] = f[c];
}

__global__ void __launch_bounds__(256) k_store_edges(PertStore s, const int64_t* __restrict__ trace_ids, int B,
                                                     const int* __restrict__ node_off,
                                                     const int* __restrict__ edge_off, PertBatchOut o) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int E = edge_off[B];
  if (e >= E) return;
  const int b = upper_seg(edge_off, B, e);
  int64_t t = trace_ids[b];
  if (t < 0 || t >= s.n_traces) t = 0;
  const int ent = s.trace_entry[t];
  int node_base, local;
  const int k = find_pattern(s, ent, e - edge_off[b], true, node_base, local);
  const int p = s.ent_pat[k];
  const size_t ge = (size_t)s.pat_eptr[p] + local;
  const int64_t off = (int64_t)node_off[b] + node_base;
  o.edge_index[e] = s.pat_src[ge] + off;
  o.edge_index[(size_t)E + e] = s.pat_dst[ge] + off;
  for (int c = 0; c < s.attr_cols; ++c) o.edge_attr[(size_t)e * s.attr_cols + c] = s.pat_attr[ge * s.attr_cols + c];
}

// per-trace outputs: entry_id, y, ptr (int64), and the concatenated per-pattern probabilities
__global__ void __launch_bounds__(256) k_store_traces(PertStore s, const int64_t* __restrict__ trace_ids, int B,
                                                      const int* __restrict__ node_off,
                                                      const int* __restrict__ pat_off, PertBatchOut o) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > B) return;
  o.ptr[b] = node_off[b];
  if (b == B) return;
  int64_t t = trace_ids[b];
  if (t < 0 || t >= s.n_traces) t = 0;
  const int ent = s.trace_entry[t];
  o.entry_id[b] = ent;
  o.y[b] = s.trace_y[t];
  const int k0 = s.ent_ptr[ent], k1 = s.ent_ptr[ent + 1];
  for (int k = k0; k < k1; ++k) o.pattern_probs[pat_off[b] + (k - k0)] = s.ent_prob[k];
}

}  // namespace

extern "C" {

int pert_store_assemble(const PertStore* s, const int64_t* trace_ids, long long B, long long N, long long E,
                        int* offsets, const PertBatchOut* out, int* status, void* stream) {
  if (!s || !out || B < 0 || N < 0 || E < 0 || (B > 0 && (!trace_ids || !offsets))) return PERT_ERR_BADARG;
  if (!s->ent_ptr || !s->ent_pat || !s->ent_prob || !s->pat_nptr || !s->pat_eptr || !s->trace_entry)
    return PERT_ERR_BADARG;
  if (B == 0) return PERT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int* node_off = offsets;
  int* edge_off = offsets + (B + 1);
  int* pat_off = offsets + 2 * (B + 1);
  k_store_offsets<<<1, 1024, 0, st>>>(*s, trace_ids, (int)B, node_off, edge_off, pat_off, status);
  k_store_traces<<<pert_cdiv(B + 1, 256), 256, 0, st>>>(*s, trace_ids, (int)B, node_off, pat_off, *out);
  if (N > 0) k_store_nodes<<<pert_cdiv(N, 256), 256, 0, st>>>(*s, trace_ids, (int)B, node_off, *out, status);
  if (E > 0) k_store_edges<<<pert_cdiv(E, 256), 256, 0, st>>>(*s, trace_ids, (int)B, node_off, edge_off, *out);
  PERT_LAUNCH_CHECK();
  return PERT_OK;
}

}  // extern "C"
