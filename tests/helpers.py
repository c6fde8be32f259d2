"""Shared helpers for the parity tests (oracle = checker, CUDA path = thing under test)."""
import numpy as np
import torch

from oracle.model_oracle import OracleSAGEDeterministic
from pert_gnn_kdd23_b200.data import Batch
from pert_gnn_kdd23_b200.synthetic import make_data_list, model_args

RTOL = 1e-4   # BASELINE.json north_star: outputs within 1e-4 relative on fp32


def rel_err(a, b):
    """max |a-b| / max|b| -- 'relative' in the sense of the tensor's scale."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = b.abs().max().clamp_min(1e-30)
    return float((a - b).abs().max() / denom)


def assert_close(a, b, rtol=RTOL, what=""):
    e = rel_err(a, b)
    assert e <= rtol, f"{what}: rel err {e:.3e} > {rtol:.1e}"


def make_batch(cfg_id, num_graphs=None, seed=None, patterns=1, edge_attr_cols=2):
    return Batch.from_data_list(make_data_list(cfg_id, num_graphs=num_graphs, seed=seed, patterns=patterns,
                                               edge_attr_cols=edge_attr_cols))


def forward_args(b):
    return (b.x, b.cat_X, b.edge_index, b.edge_attr, b.pattern_num_nodes, b.rt_probs, b.entry_id, b.batch)


def make_models(cfg_id, seed=0, dtype=torch.float32):
    """(oracle on CPU, CUDA model) with identical weights (copied, never relying on RNG order)."""
    from pert_gnn_kdd23_b200.model import SAGEDeterministic

    torch.manual_seed(seed)
    oracle = OracleSAGEDeterministic(*model_args(cfg_id)).to(dtype)
    model = SAGEDeterministic(*model_args(cfg_id))
    model.load_state_dict({k: v.float() for k, v in oracle.state_dict().items()})
    return oracle, model.cuda()
