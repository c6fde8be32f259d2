// PERT-graph and span-graph construction on the GPU.
//
// Span graph (misc.py:190-219, the `--graph_type span` default of pert_gnn.py:32): nodes = sorted unique microservice
// ids (torch.unique(sorted=True, return_inverse=True)), one edge per span row in table order, edge_attr =
// [interface, rpctype]; fully specified, so the device result equals the reference's tensors bit for bit.
//
// The reference builds one PERT graph per runtime pattern on the host with pandas row loops
// (misc.py:221-319, GraphConstruct.get_pert_edge_index):
//   * every caller `um` with c calls becomes a chain of 2c+1 stage nodes (edge stage_s -> stage_{s+1},
//     attr [0,0,1,1]); every microservice that is only called becomes one node                           (:238-255)
//   * per caller, its 2c call/return events are ordered by time (stable: ties keep row order, a row's start before
//     its end); event i is either a call edge  stage_i(um) -> stage_0(dm)  attr [interface, rpctype, 1, 0]  or a
//     return edge  stage_last(dm) -> stage_{i+1}(um)  attr [0, 0, 0, 0]                                    (:271-302)
//   * sorted_span_id = the microservice of every node, root_nid = stage_0 of the root microservice          (:304-311)
// Here one CTA builds one trace entirely in shared memory: de-duplicate the microservice ids (rank sort), count
// calls, place the node blocks, rank every event inside its caller by (time, row, mode), and write nodes and the 4r
// edges straight to their final slots.  No sort passes over global memory; the input rows are read once.
//
// Node numbering.  The reference numbers callers in pandas value_counts() order (count descending, ties in hash-table
// order) and leaves in Python set order -- both implementation-defined.  The device path uses the canonical order
//   callers by (count descending, microservice id ascending), then leaves by microservice id ascending,
// and emits edges as [chain edges in node order | per caller in ascending id: its events in time order] (the second
// part is exactly the reference's order: groupby("um") iterates ascending keys).  tests/ compare with the reference's
// own outputs up to that relabelling (oracle/pert_graph_oracle.py:canonical_form).
#include "common.cuh"

namespace {

struct PgArgs {
  const int64_t* row_ptr;   // [T+1] rows of trace t
  const int64_t *um, *dm, *itf, *rpc, *t0, *t1;   // [R]
  const int64_t* root_ms;   // [T]
  const int64_t* node_ptr;  // [T+1] (build pass)
  int64_t* node_cnt;        // [T]   (count pass)
  int64_t* ms_id;           // [N]
  int64_t* edge_index;      // [2, 4R]
  int64_t* edge_attr;       // [4R, 4]
  int64_t* root_nid;        // [T] global node id
  long long R;              // total rows (edge_index row stride = 4R)
  int max_rows, global_ids, span;
  int* status;
};

// shared memory carve-up for a trace of at most `mr` rows (2*mr microservice slots)
__host__ __device__ inline size_t pg_smem_bytes(int mr) {
  return (size_t)mr * 2 * (sizeof(int64_t) * 2 + sizeof(int) * 5) + (size_t)mr * 2 * sizeof(int) + 64;
}

#pragma nv_diag_suppress 128   // the count instantiation returns before the build half
// MODE 0: node counts (PERT: 2 rows + distinct ids; span: distinct ids)   1: PERT graph   2: span graph
template <int MODE>
__global__ void __launch_bounds__(256) k_pert_graph(PgArgs a) {
  constexpr bool BUILD = MODE != 0;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int t = blockIdx.x;
  const long long r0 = a.row_ptr[t];
  const int r = (int)(a.row_ptr[t + 1] - r0);
  if (r < 0 || r > a.max_rows) {
    if (threadIdx.x == 0) {
      if (a.status) atomicExch(a.status, PERT_ERR_RANGE);
      if (!BUILD) a.node_cnt[t] = 0;
    }
    return;
  }
  const int mr = a.max_rows, m2 = 2 * r;
  int64_t* ms = reinterpret_cast<int64_t*>(smem_raw);      // [2mr] um | dm of every row
  int64_t* u = ms + 2 * mr;                                // [2mr] distinct ids, ascending
  int* first = reinterpret_cast<int*>(u + 2 * mr);         // [2mr] first occurrence flag, later: node base
  int* cnt = first + 2 * mr;                               // [2mr] calls made by u[d]
  int* base = cnt + 2 * mr;                                // [2mr] first node of u[d] (trace-local)
  int* cbase = base + 2 * mr;                              // [2mr] first chain-edge slot of u[d]
  int* gbase = cbase + 2 * mr;                             // [2mr] first event-edge slot of u[d]
  int* ui = gbase + 2 * mr;                                // [mr] index of um[i] in u
  int* di = ui + mr;                                       // [mr] index of dm[i] in u
  __shared__ int D_s;

  for (int i = threadIdx.x; i < r; i += blockDim.x) {
    ms[i] = a.um[r0 + i];
    ms[r + i] = a.dm[r0 + i];
  }
  if (threadIdx.x == 0) D_s = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < m2; i += blockDim.x) {
    const int64_t v = ms[i];
    int f = 1;
    for (int j = 0; j < i; ++j) f &= (ms[j] != v);
    first[i] = f;
    if (f) atomicAdd(&D_s, 1);
  }
  __syncthreads();
  const int D = D_s;
  if (!BUILD) {
    if (threadIdx.x == 0) a.node_cnt[t] = a.span ? (long long)D : 2LL * r + D;
    return;
  }
  for (int i = threadIdx.x; i < m2; i += blockDim.x) {
    if (!first[i]) continue;
    const int64_t v = ms[i];
    int rk = 0;
    for (int j = 0; j < m2; ++j) rk += (first[j] && ms[j] < v);
    u[rk] = v;
  }
  for (int d = threadIdx.x; d < D; d += blockDim.x) cnt[d] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < m2; i += blockDim.x) {
    const int64_t v = ms[i];
    int lo = 0, hi = D - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (u[mid] < v) lo = mid + 1;
      else hi = mid;
    }
    if (i < r) {
      ui[i] = lo;
      atomicAdd(&cnt[lo], 1);
    } else {
      di[i - r] = lo;
    }
  }
  __syncthreads();
  if (MODE == 2) {
    // span graph (misc.py:190-219): nodes = sorted unique ids (torch.unique), one edge per row in table order
    const long long n0 = a.node_ptr[t];
    const long long goff = a.global_ids ? n0 : 0;
    for (int d = threadIdx.x; d < D; d += blockDim.x) a.ms_id[n0 + d] = u[d];
    for (int i = threadIdx.x; i < r; i += blockDim.x) {
      const long long e = r0 + i;
      a.edge_index[e] = goff + ui[i];
      a.edge_index[a.R + e] = goff + di[i];
      reinterpret_cast<longlong2*>(a.edge_attr)[e] = make_longlong2(a.itf[r0 + i], a.rpc[r0 + i]);
    }
    if (threadIdx.x == 0) {
      const int64_t rm = a.root_ms[t];
      int lo = 0, hi = D - 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (u[mid] < rm) lo = mid + 1;
        else hi = mid;
      }
      if (D > 0 && u[lo] == rm) {
        a.root_nid[t] = n0 + lo;
      } else {
        a.root_nid[t] = -1;
        if (a.status) atomicExch(a.status, PERT_ERR_RANGE);
      }
    }
    return;
  }
  // block placement: callers by (count desc, id asc), then leaves by id asc.  u is ascending, so id order = index order.
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const int c = cnt[d];
    int nb = 0, cb = 0, gb = 0;
    for (int e = 0; e < D; ++e) {
      const int ce = cnt[e];
      bool before;
      if (c > 0) before = ce > c || (ce == c && e < d);
      else before = ce > 0 || e < d;
      if (before) {
        nb += ce > 0 ? 2 * ce + 1 : 1;
        cb += 2 * ce;
      }
      if (e < d) gb += 2 * ce;
    }
    base[d] = nb;
    cbase[d] = cb;
    gbase[d] = 2 * r + gb;
  }
  __syncthreads();
  const long long n0 = a.node_ptr[t];
  const long long e0 = 4 * r0, ES = 4 * a.R;
  const long long goff = a.global_ids ? n0 : 0;
  int64_t* src = a.edge_index;
  int64_t* dst = a.edge_index + ES;
  // nodes + chain edges: one warp per microservice block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int d = warp; d < D; d += nw) {
    const int c = cnt[d], nb = base[d], cb = cbase[d];
    const int sz = c > 0 ? 2 * c + 1 : 1;
    const int64_t v = u[d];
    for (int s = lane; s < sz; s += 32) a.ms_id[n0 + nb + s] = v;
    for (int s = lane; s < 2 * c; s += 32) {
      const long long e = e0 + cb + s;
      src[e] = goff + nb + s;
      dst[e] = goff + nb + s + 1;
      reinterpret_cast<longlong4*>(a.edge_attr)[e] = make_longlong4(0, 0, 1, 1);
    }
  }
  // call / return edges: rank of each event inside its caller by (time, row, mode)
  for (int i = threadIdx.x; i < r; i += blockDim.x) {
    const int g = ui[i], dd = di[i];
    const int64_t ts = a.t0[r0 + i], te = a.t1[r0 + i];
    int rs = 0, re = 0;
    for (int j = 0; j < r; ++j) {
      if (ui[j] != g) continue;
      const int64_t sj = a.t0[r0 + j], ej = a.t1[r0 + j];
      // event (time, 2*row + mode) strictly before (ts, 2i) / (te, 2i+1)
      rs += (sj < ts || (sj == ts && j < i)) + (ej < ts || (ej == ts && j < i));
      re += (sj < te || (sj == te && j <= i)) + (ej < te || (ej == te && j < i));
    }
    const int ub = base[g], db = base[dd], dlast = db + 2 * cnt[dd];
    long long e = e0 + gbase[g] + rs;
    src[e] = goff + ub + rs;
    dst[e] = goff + db;
    reinterpret_cast<longlong4*>(a.edge_attr)[e] = make_longlong4(a.itf[r0 + i], a.rpc[r0 + i], 1, 0);
    e = e0 + gbase[g] + re;
    src[e] = goff + dlast;
    dst[e] = goff + ub + re + 1;
    reinterpret_cast<longlong4*>(a.edge_attr)[e] = make_longlong4(0, 0, 0, 0);
  }
  if (threadIdx.x == 0) {
    const int64_t rm = a.root_ms[t];
    int lo = 0, hi = D - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (u[mid] < rm) lo = mid + 1;
      else hi = mid;
    }
    if (D > 0 && u[lo] == rm) {
      a.root_nid[t] = n0 + base[lo];
    } else {
      a.root_nid[t] = -1;                     // the reference raises KeyError (stages[self.root_span])
      if (a.status) atomicExch(a.status, PERT_ERR_RANGE);
    }
  }
}

int pg_launch(int mode, const PgArgs& a, long long T, cudaStream_t st) {
  if (T <= 0) return 0;
  if (a.max_rows <= 0 || a.max_rows > PERT_PERT_GRAPH_MAX_ROWS) return PERT_ERR_UNSUPPORTED;
  const size_t smem = pg_smem_bytes(a.max_rows);
  auto fn = mode == 0 ? k_pert_graph<0> : (mode == 1 ? k_pert_graph<1> : k_pert_graph<2>);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  fn<<<(unsigned)T, 256, smem, st>>>(a);
  PERT_LAUNCH_CHECK();
  return 0;
}

int pg_count(int span, const int64_t* row_ptr, long long T, const int64_t* um, const int64_t* dm, int max_rows,
             int64_t* node_cnt, int* status, void* stream) {
  if (!row_ptr || !um || !dm || !node_cnt || T < 0) return PERT_ERR_BADARG;
  PgArgs a{};
  a.row_ptr = row_ptr;
  a.um = um;
  a.dm = dm;
  a.node_cnt = node_cnt;
  a.max_rows = max_rows;
  a.span = span;
  a.status = status;
  return pg_launch(0, a, T, (cudaStream_t)stream);
}

}  // namespace

extern "C" int pert_pert_graph_count(const int64_t* row_ptr, long long T, const int64_t* um, const int64_t* dm,
                                     int max_rows, int64_t* node_cnt, int* status, void* stream) {
  return pg_count(0, row_ptr, T, um, dm, max_rows, node_cnt, status, stream);
}
extern "C" int pert_span_graph_count(const int64_t* row_ptr, long long T, const int64_t* um, const int64_t* dm,
                                     int max_rows, int64_t* node_cnt, int* status, void* stream) {
  return pg_count(1, row_ptr, T, um, dm, max_rows, node_cnt, status, stream);
}

extern "C" int pert_pert_graph_build(const int64_t* row_ptr, long long T, long long R, const int64_t* um,
                                     const int64_t* dm, const int64_t* interface, const int64_t* rpctype,
                                     const int64_t* t_start, const int64_t* t_end, const int64_t* root_ms,
                                     const int64_t* node_ptr, int max_rows, int global_ids, int64_t* ms_id,
                                     int64_t* edge_index, int64_t* edge_attr, int64_t* root_nid, int* status,
                                     void* stream) {
  if (!row_ptr || !um || !dm || !interface || !rpctype || !t_start || !t_end || !root_ms || !node_ptr || !ms_id ||
      !edge_index || !edge_attr || !root_nid || T < 0 || R < 0)
    return PERT_ERR_BADARG;
  PgArgs a{};
  a.row_ptr = row_ptr;
  a.um = um;
  a.dm = dm;
  a.itf = interface;
  a.rpc = rpctype;
  a.t0 = t_start;
  a.t1 = t_end;
  a.root_ms = root_ms;
  a.node_ptr = node_ptr;
  a.ms_id = ms_id;
  a.edge_index = edge_index;
  a.edge_attr = edge_attr;
  a.root_nid = root_nid;
  a.R = R;
  a.max_rows = max_rows;
  a.global_ids = global_ids;
  a.status = status;
  return pg_launch(1, a, T, (cudaStream_t)stream);
}

extern "C" int pert_span_graph_build(const int64_t* row_ptr, long long T, long long R, const int64_t* um,
                                     const int64_t* dm, const int64_t* interface, const int64_t* rpctype,
                                     const int64_t* root_ms, const int64_t* node_ptr, int max_rows, int global_ids,
                                     int64_t* ms_id, int64_t* edge_index, int64_t* edge_attr, int64_t* root_nid,
                                     int* status, void* stream) {
  if (!row_ptr || !um || !dm || !interface || !rpctype || !root_ms || !node_ptr || !ms_id || !edge_index ||
      !edge_attr || !root_nid || T < 0 || R < 0)
    return PERT_ERR_BADARG;
  PgArgs a{};
  a.row_ptr = row_ptr;
  a.um = um;
  a.dm = dm;
  a.itf = interface;
  a.rpc = rpctype;
  a.root_ms = root_ms;
  a.node_ptr = node_ptr;
  a.ms_id = ms_id;
  a.edge_index = edge_index;
  a.edge_attr = edge_attr;
  a.root_nid = root_nid;
  a.R = R;
  a.max_rows = max_rows;
  a.global_ids = global_ids;
  a.span = 1;
  a.status = status;
  return pg_launch(2, a, T, (cudaStream_t)stream);
}
