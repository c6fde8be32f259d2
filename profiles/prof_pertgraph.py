"""Times PERT-graph construction (row N2) on the GPU against the CPU oracle restatement of misc.py:221-319.
Usage (GPU box):  python profiles/prof_pertgraph.py [T]    -> one JSON line
Timed region (e2e): H2D of the concatenated span rows + count + build + level index + node_depth, i.e.
pertgraph.build_pert_graphs_flat from host arrays."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pert_graph_oracle as O  # noqa: E402
from pert_gnn_kdd23_b200 import pertgraph  # noqa: E402
from pert_gnn_kdd23_b200.synthetic import make_span_tables  # noqa: E402


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    raw = make_span_tables(3, T, n_ms=400, calls=(2, 60), anomalies=False)
    tables, roots = [], []
    for tab in raw:
        root = pertgraph.get_root_ms(tab)
        keep = pertgraph.drop_wrong_edges(tab, root)
        tables.append({k: tab[k][keep] for k in tab})
        roots.append(root)
    rows = sum(len(t["um"]) for t in tables)
    host_cols = np.stack([np.concatenate([t[c] for t in tables]) for c in pertgraph.COLUMNS])
    host_ptr = np.concatenate([[0], np.cumsum([len(t["um"]) for t in tables])]).astype(np.int64)
    for _ in range(3):
        pg = pertgraph.build_pert_graphs_flat(host_cols, host_ptr, roots).check()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        pg = pertgraph.build_pert_graphs_flat(host_cols, host_ptr, roots)
    torch.cuda.synchronize()
    gpu_s = (time.perf_counter() - t0) / reps
    # kernels alone (inputs resident): count + build
    dev = torch.device("cuda")
    cols = torch.from_numpy(np.stack([np.concatenate([t[c] for t in tables]) for c in pertgraph.COLUMNS])).to(dev)
    rp = torch.from_numpy(np.concatenate([[0], np.cumsum([len(t["um"]) for t in tables])]).astype(np.int64)).to(dev)
    rm = torch.tensor(roots, dtype=torch.int64, device=dev)
    from pert_gnn_kdd23_b200 import _lib
    N = int(pg.node_ptr[-1])
    node_ptr = torch.from_numpy(pg.node_ptr).to(dev)
    out = [torch.empty(N, dtype=torch.int64, device=dev), torch.empty(2, 4 * rows, dtype=torch.int64, device=dev),
           torch.empty(4 * rows, 4, dtype=torch.int64, device=dev), torch.empty(T, dtype=torch.int64, device=dev)]
    cnt = torch.empty(T, dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    mr = max(len(t["um"]) for t in tables)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def kernels():
        _lib.call("pert_pert_graph_count", rp.data_ptr(), T, cols[0].data_ptr(), cols[1].data_ptr(), mr,
                  cnt.data_ptr(), status.data_ptr(), _lib.stream())
        _lib.call("pert_pert_graph_build", rp.data_ptr(), T, rows, *(cols[c].data_ptr() for c in range(6)),
                  rm.data_ptr(), node_ptr.data_ptr(), mr, 0, *(o.data_ptr() for o in out), status.data_ptr(),
                  _lib.stream())
    for _ in range(3):
        kernels()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(20):
        kernels()
    ev[1].record()
    torch.cuda.synchronize()
    k_ms = ev[0].elapsed_time(ev[1]) / 20
    byts = rows * 6 * 8 + N * 8 + 4 * rows * (16 + 32) + T * 24
    # CPU oracle on a bounded sample
    S = min(T, 256)
    t0 = time.perf_counter()
    for c, r in zip(tables[:S], roots[:S]):
        O.pert_graph(c["um"], c["dm"], c["interface"], c["rpctype"], c["timestamp"], c["endTimestamp"], r)
    cpu_s = (time.perf_counter() - t0) / S
    print(json.dumps({"traces": T, "span_rows": rows, "pert_nodes": N, "pert_edges": 4 * rows,
                      "e2e_traces_per_s": T / gpu_s, "e2e_ms": gpu_s * 1e3, "kernels_ms": k_ms,
                      "kernels_traces_per_s": T / (k_ms * 1e-3), "kernel_alg_bytes": byts,
                      "kernel_GBps": byts / (k_ms * 1e-3) / 1e9, "cpu_oracle_traces_per_s": 1 / cpu_s,
                      "cpu_sample": f"{S} traces, 1 core, python loops (oracle/pert_graph_oracle.py)"}))


if __name__ == "__main__":
    main()
