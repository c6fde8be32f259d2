"""CPU tests: collation rules, CSR/CSC definition, level index against the reference-generated golden vectors."""
import glob
import os

import numpy as np
import torch

from oracle import index_oracle
from pert_gnn_kdd23_b200.data import Batch, Data, DataLoader
from pert_gnn_kdd23_b200.synthetic import bfs_min_depth, make_data_list

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_node_depth_golden_from_reference():
    """tests/golden/node_depth_*.npz were produced by the reference's own misc.DFS / get_node_features
    (oracle/gen_golden.py): the restatements must reproduce them bit-exactly."""
    files = sorted(glob.glob(os.path.join(GOLD, "node_depth_*.npz")))
    assert len(files) >= 9
    for f in files:
        z = np.load(f)
        ei, n, root = z["edge_index"], int(z["num_nodes"]), int(z["root"])
        d = index_oracle.dfs_min_depth(ei, n, root)
        assert np.array_equal(d, z["min_depth"]), f
        assert np.array_equal(index_oracle.node_depth_tensor(d), z["node_depth"]), f
        assert np.array_equal(bfs_min_depth(ei, n, root).astype(np.int32), z["min_depth"]), f   # generator's BFS


def test_build_index_known_answer():
    #            e0     e1     e2     e3     e4
    ei = np.array([[2, 0, 2, 1, 0], [1, 1, 0, 1, 2]])
    r = index_oracle.build_index(ei, 4)
    assert r["rowptr"].tolist() == [0, 1, 4, 5, 5]
    assert r["perm"].tolist() == [2, 0, 1, 3, 4]            # stable inside target 1: e0, e1, e3
    assert r["csr_src"].tolist() == [2, 2, 0, 1, 0]
    assert r["colptr"].tolist() == [0, 2, 3, 5, 5]
    # CSC order: src0: e1,e4 ; src1: e3 ; src2: e0,e2 -> their CSR slots
    assert r["csc_pos"].tolist() == [2, 4, 3, 1, 0]
    assert r["csc_dst"].tolist() == [1, 2, 1, 1, 0]


def test_collate_rules_match_pyg_semantics():
    dl = make_data_list(1, 5, patterns=2, edge_attr_cols=4)
    b = Batch.from_data_list(dl)
    ref = index_oracle.collate([{k: (v.numpy() if torch.is_tensor(v) else v) for k, v in d.items()} for d in dl])
    for k, v in ref.items():
        assert np.array_equal(b[k].numpy(), v), k
    assert b.num_graphs == 5 and b.y.shape == (5,) and b.entry_id.shape == (5,)
    assert b.edge_index.shape[0] == 2 and b.pattern_probs.shape == (10, 1)
    # edge_index offsets: every edge stays inside its graph's node range
    g_of_src = b.batch[b.edge_index[0]]
    g_of_dst = b.batch[b.edge_index[1]]
    assert torch.equal(g_of_src, g_of_dst)
    assert np.array_equal(index_oracle.graph_ptr(b.batch.numpy(), 5), b.ptr.numpy().astype(np.int32))


def test_data_container_and_loader():
    d = Data(x=torch.zeros(3, 2), edge_index=torch.tensor([[0, 1], [1, 2]]), y=torch.tensor(5), foo=torch.ones(3, 1))
    assert d.num_nodes == 3 and d.num_edges == 2 and "foo" in d and d.foo.shape == (3, 1)
    d2 = d.to("cpu")
    assert d2 is not d and torch.equal(d2.x, d.x)
    dl = make_data_list(1, 10)
    loader = DataLoader(dl, batch_size=4, shuffle=False)
    assert len(loader.dataset) == 10
    sizes = [b.num_graphs for b in loader]
    assert sizes == [4, 4, 2]
    b = next(iter(loader))
    slab = b.pin_memory()
    for k in b.keys():
        assert torch.equal(slab[k], b[k]), k
    assert slab.h2d_bytes >= sum(v.numel() * v.element_size() for v in b.to_dict().values() if torch.is_tensor(v))


def test_level_order():
    ptr = np.array([0, 3, 7])
    level = np.array([0, 2, 1, 1, 0, -1, 1])
    order = index_oracle.level_order(ptr, level)
    assert order.tolist() == [0, 2, 1, 4, 3, 6, 5]


def test_shard_by_edges_balances_power_law_sizes():
    """Data-parallel partition by edge count: complete, disjoint, deterministic, and far better
    balanced than an equal-count split on power-law graph sizes."""
    from pert_gnn_kdd23_b200.data import shard_by_edges
    from pert_gnn_kdd23_b200.synthetic import make_data_list

    dl = make_data_list(3)[:256]                     # cfg3: truncated-Pareto graph sizes
    world = 8
    parts = shard_by_edges(dl, world)
    assert parts == shard_by_edges(dl, world)
    flat = sorted(i for p in parts for i in p)
    assert flat == list(range(len(dl)))
    edges = [sum(dl[i].num_edges for i in p) for p in parts]
    naive = [sum(d.num_edges for d in dl[r * 32:(r + 1) * 32]) for r in range(world)]
    assert max(edges) - min(edges) <= max(d.num_edges for d in dl)
    assert (max(edges) / (sum(edges) / world)) <= (max(naive) / (sum(naive) / world))
    assert max(edges) / (sum(edges) / world) < 1.05
    assert shard_by_edges(dl[:3], 4)[3] == []        # more ranks than graphs: empty shards allowed
