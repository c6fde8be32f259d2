"""Integer/index oracles (numpy, bit-exact targets).  TEST INFRASTRUCTURE.

* ``collate`` restates PyG ``Batch.from_data_list`` for the reference's Data
  schema (reference pert_gnn.py:163-173 builds the Data; :201-209 collates).
* ``build_index`` is the numpy definition of the CSR/CSC layout the CUDA
  index-construction kernels must reproduce bit-exactly (there is no such
  layout in the reference -- PyG works on COO; this is new).
* ``dfs_min_depth`` / ``node_depth_tensor`` restate the reference's level
  index: misc.py:52-63 (DFS.dfs_min_node_depth), :107-111 (build_adj_list),
  :113-136 (get_node_depth), :155-175 (inf->0, normalise by max) and the
  ``torch.tensor(node_depth, dtype=torch.long)`` truncation at :215/:368.
  This row is PINNED: tests/golden/node_depth_*.npz were produced by running
  the reference's own ``misc.DFS`` (oracle/gen_golden.py).
"""
import numpy as np


# ---------------------------------------------------------------- CSR / CSC
def build_index(edge_index, num_nodes):
    """edge_index: int64 [2,E] (row0 = source j, row1 = target i).

    Returns dict of int32 arrays:
      rowptr [N+1]  CSR by target;   perm [E] original edge id at CSR slot p
                    (stable: increasing edge id inside a target segment)
      csr_src [E]   source of the edge at slot p
      colptr [N+1]  CSC by source;   csc_pos [E] CSR slot of the edge at CSC
                    slot c (stable: increasing edge id inside a source segment)
      csc_dst [E]   target of the edge at CSC slot c
    """
    src = np.asarray(edge_index[0], dtype=np.int64)
    dst = np.asarray(edge_index[1], dtype=np.int64)
    E = src.shape[0]
    perm = np.argsort(dst, kind="stable")
    rowptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.add.at(rowptr, dst + 1, 1)
    rowptr = np.cumsum(rowptr)
    inv = np.empty(E, dtype=np.int64)
    inv[perm] = np.arange(E)
    cperm = np.argsort(src, kind="stable")
    colptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.add.at(colptr, src + 1, 1)
    colptr = np.cumsum(colptr)
    return {
        "rowptr": rowptr.astype(np.int32),
        "perm": perm.astype(np.int32),
        "csr_src": src[perm].astype(np.int32),
        "colptr": colptr.astype(np.int32),
        "csc_pos": inv[cperm].astype(np.int32),
        "csc_dst": dst[cperm].astype(np.int32),
    }


def graph_ptr(batch, num_graphs):
    """ptr [B+1] from a (sorted) PyG ``batch`` vector: node range of each graph."""
    counts = np.bincount(np.asarray(batch, dtype=np.int64), minlength=num_graphs)
    return np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)


# ---------------------------------------------------------------- levels
def dfs_min_depth(edge_index, num_nodes, root):
    """Raw min hop depth from ``root`` over out-edges; -1 where unreachable.

    Restates misc.py:59-63 (relaxing DFS) + :107-111 (adjacency by edge order);
    iterative (explicit stack) so deep chains do not hit Python's recursion limit;
    the fixed point of the relaxation is the BFS distance, independent of order.
    """
    src = np.asarray(edge_index[0], dtype=np.int64)
    dst = np.asarray(edge_index[1], dtype=np.int64)
    adj = [[] for _ in range(num_nodes)]
    for s, d in zip(src.tolist(), dst.tolist()):
        adj[s].append(d)
    INF = float("inf")
    depth = [INF] * num_nodes
    stack = [(int(root), 0)]
    while stack:
        v, d = stack.pop()
        if depth[v] > d:
            depth[v] = d
            for nb in reversed(adj[v]):
                stack.append((nb, d + 1))
    return np.array([-1 if x == INF else int(x) for x in depth], dtype=np.int32)


def node_depth_tensor(min_depth):
    """misc.py:159-175 + :215/:368: inf->0, divide by max (or 1), float->int64
    truncation.  Result [N,1] int64 with values in {0,1} (1 only at the deepest level)."""
    d = np.asarray(min_depth, dtype=np.float64).copy()
    d[d < 0] = 0.0
    norm = d.max() if d.size and d.max() > 0 else 1.0
    nd = np.array([d / norm]).T
    return nd.astype(np.int64)        # torch.tensor(float ndarray, dtype=long) truncates toward zero


def level_order(ptr, level):
    """Level-major node order inside each graph (layout key):
    order = stable argsort of (graph, level); level -1 (unreachable) sorts last."""
    ptr = np.asarray(ptr, dtype=np.int64)
    lv = np.asarray(level, dtype=np.int64)
    n = lv.shape[0]
    graph = np.repeat(np.arange(len(ptr) - 1), np.diff(ptr))
    key_level = np.where(lv < 0, np.iinfo(np.int32).max, lv)
    order = np.lexsort((np.arange(n), key_level, graph))
    return order.astype(np.int32)


# ---------------------------------------------------------------- collate
def collate(data_list):
    """PyG Batch.from_data_list for dicts of numpy arrays / torch tensors.

    Rules: cat along dim 0 for every key except keys containing
    'index' (cat along last dim, incremented by the cumulative node count
    ``x.shape[0]``); 0-dim values are stacked; adds ``batch`` [N] and ``ptr`` [B+1].
    """
    out = {}
    keys = list(data_list[0].keys())
    n_nodes = [int(np.asarray(d["x"]).shape[0]) for d in data_list]
    offs = np.concatenate([[0], np.cumsum(n_nodes)])
    for k in keys:
        vals = [np.asarray(d[k]) for d in data_list]
        if vals[0].ndim == 0:
            out[k] = np.stack(vals)
        elif "index" in k:
            out[k] = np.concatenate([v + offs[i] for i, v in enumerate(vals)], axis=-1)
        else:
            out[k] = np.concatenate(vals, axis=0)
    out["batch"] = np.repeat(np.arange(len(data_list)), n_nodes).astype(np.int64)
    out["ptr"] = offs.astype(np.int64)
    return out
