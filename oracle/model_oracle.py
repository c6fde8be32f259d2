"""Vectorised pure-torch CPU restatement of the reference model (TEST INFRASTRUCTURE).

PARITY UNPINNED for the arithmetic (see oracle/__init__.py): torch_geometric
2.4.0 is absent, so every PyG call made by the reference is replaced by the
PyG 2.4.0 semantics enumerated in SURVEY.md section 8c, written with the same
ATen ops PyG lowers to (index_select / scatter_reduce_('amax',
include_self=False) / scatter_add_ / Linear).

Follows, line by line:
  reference model.py:11-68   constructor (layer counts, submodule names)
  reference model.py:70-74   reset_parameters (convs + bns only)
  reference model.py:76-114  forward
Module / parameter names are identical to the reference so ``state_dict``
keys match (``convs.{i}.lin_{key,query,value,edge,skip}``, ``bns.{i}``,
``local_linear``, ``global_linear{1,2}``, ``cat_embedding.{i}``,
``entry_embeds``, ``interface_embeds``, ``rpctype_embeds``).  The reference's
``edge_linear = Linear(-1, 2H)`` (model.py:68) is lazy and never used: in PyG 2.4.0 its
weight stays an UninitializedParameter (not restated) while its bias IS a real [2H]
parameter (uninitialised memory there; zeros here) that never receives a gradient,
so ``edge_linear.bias`` is kept for ``state_dict`` / ``parameters()`` parity.
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# PyG 2.4.0 utils restated
# --------------------------------------------------------------------------
def scatter(src, index, dim_size, reduce):
    """torch_geometric.utils.scatter (2.4.0) along dim 0.

    'sum': zeros.scatter_add_;  'max': zeros.scatter_reduce_('amax',
    include_self=False) -> rows that receive nothing stay 0.
    """
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    size = (dim_size,) + tuple(src.shape[1:])
    if reduce == "sum":
        return src.new_zeros(size).scatter_add_(0, idx, src)
    if reduce == "max":
        return src.new_zeros(size).scatter_reduce_(0, idx, src, reduce="amax", include_self=False)
    raise ValueError(reduce)


def segment_softmax(src, index, num_nodes):
    """torch_geometric.utils.softmax (2.4.0), index path (ptr=None).

    max on src.detach(); exp of the shifted logits; denominator + 1e-16.
    """
    src_max = scatter(src.detach(), index, num_nodes, "max")
    out = (src - src_max.index_select(0, index)).exp()
    out_sum = scatter(out, index, num_nodes, "sum") + 1e-16
    return out / out_sum.index_select(0, index)


def global_add_pool(x, batch, size=None):
    """torch_geometric.nn.global_add_pool: zeros[B,H].scatter_add_(0, batch, x);
    B = int(batch.max()) + 1 unless given (reference call site model.py:107)."""
    if size is None:
        size = int(batch.max()) + 1 if batch.numel() > 0 else 0
    return scatter(x, batch, size, "sum")


class PygLinear(torch.nn.Linear):
    """torch_geometric.nn.Linear(in, out): same math and default init as
    nn.Linear (kaiming_uniform(a=sqrt(5)) weight, U(-1/sqrt(in), 1/sqrt(in)) bias)."""


# --------------------------------------------------------------------------
# TransformerConv(heads=1, concat=True, beta=False, dropout=0, edge_dim=2H,
#                 bias=True, root_weight=True), aggr='add', source_to_target
# --------------------------------------------------------------------------
class OracleTransformerConv(torch.nn.Module):
    def __init__(self, in_channels, out_channels, heads=1, edge_dim=None):
        super().__init__()
        assert heads == 1, "reference only instantiates heads=1 (model.py:26-51)"
        self.in_channels, self.out_channels, self.heads, self.edge_dim = (
            in_channels, out_channels, heads, edge_dim)
        self.lin_key = PygLinear(in_channels, out_channels)
        self.lin_query = PygLinear(in_channels, out_channels)
        self.lin_value = PygLinear(in_channels, out_channels)
        self.lin_edge = PygLinear(edge_dim, out_channels, bias=False) if edge_dim is not None else None
        self.lin_skip = PygLinear(in_channels, out_channels)

    def reset_parameters(self):
        for lin in (self.lin_key, self.lin_query, self.lin_value, self.lin_edge, self.lin_skip):
            if lin is not None:
                lin.reset_parameters()

    def forward(self, x, edge_index, edge_attr=None, return_alpha=False):
        C = self.out_channels
        n = x.size(0)
        src, dst = edge_index[0], edge_index[1]
        q = self.lin_query(x)
        k = self.lin_key(x)
        v = self.lin_value(x)
        # MessagePassing.__collect__: *_i <- index by edge_index[1], *_j <- edge_index[0]
        q_i = q.index_select(0, dst)
        k_j = k.index_select(0, src)
        v_j = v.index_select(0, src)
        if self.lin_edge is not None:
            e = self.lin_edge(edge_attr)
            k_j = k_j + e
        alpha = (q_i * k_j).sum(dim=-1) / math.sqrt(C)
        alpha = segment_softmax(alpha, dst, n)
        msg = v_j
        if self.lin_edge is not None:
            msg = msg + e
        msg = msg * alpha.view(-1, 1)
        out = scatter(msg, dst, n, "sum")          # aggr='add'
        out = out + self.lin_skip(x)               # root_weight
        if return_alpha:
            return out, alpha
        return out


class OracleSAGEDeterministic(torch.nn.Module):
    """reference model.py:10-114 with PyG calls replaced by the restatements above."""

    def __init__(self, in_channels, cat_dims, entry_id_max, interface_id_max,
                 rpctype_id_max, hidden_channels, num_layers, dropout):
        super().__init__()
        H = hidden_channels
        self.convs = torch.nn.ModuleList()
        self.convs.append(OracleTransformerConv(in_channels + H, H, heads=1, edge_dim=2 * H))
        self.bns = torch.nn.ModuleList()
        self.bns.append(torch.nn.BatchNorm1d(H))
        for _ in range(num_layers - 2):                       # model.py:35
            self.convs.append(OracleTransformerConv(H, H, heads=1, edge_dim=2 * H))
            self.bns.append(torch.nn.BatchNorm1d(H))
        self.convs.append(OracleTransformerConv(H, H, heads=1, edge_dim=2 * H))
        self.local_linear = PygLinear(H, 1)
        self.global_linear1 = PygLinear(2 * H, H)
        self.global_linear2 = PygLinear(H, 1)
        self.cat_embedding = torch.nn.ModuleList(
            [torch.nn.Embedding(n, H) for n in cat_dims])
        self.dropout = dropout
        self.entry_embeds = torch.nn.Embedding(entry_id_max + 1, H)
        self.interface_embeds = torch.nn.Embedding(interface_id_max + 1, H)
        self.rpctype_embeds = torch.nn.Embedding(rpctype_id_max + 1, H)
        self.edge_linear = torch.nn.Module()                  # model.py:68: lazy Linear(-1, 2H), never used
        self.edge_linear.bias = torch.nn.Parameter(torch.zeros(2 * H))

    def reset_parameters(self):
        for conv in self.convs:
            conv.reset_parameters()
        for bn in self.bns:
            bn.reset_parameters()

    def forward(self, x, cat_X, edge_index, edge_attr, pattern_num_nodes,
                pattern_probs, entry_id, batch, relu_masks=None):
        """``relu_masks`` (test aid, not in the reference): dict {'bn{i}': [N,H] bool, 'head': [B,H] bool}; when given,
        every ReLU is replaced by multiplication with that fixed mask, i.e. the network is evaluated and differentiated
        on ONE prescribed linear piece (the piece another fp32 implementation landed on) instead of the piece this
        precision happens to select for arguments within rounding distance of zero."""
        relu = (lambda t, key: F.relu(t)) if relu_masks is None else (lambda t, key: t * relu_masks[key].to(t.dtype))
        cat_embeds = 0
        for i, emb in enumerate(self.cat_embedding):
            cat_embeds = cat_embeds + emb(cat_X[:, i])
        x = torch.cat([x, cat_embeds], dim=1)
        edge_embeds = torch.cat(
            [self.interface_embeds(edge_attr[:, 0]), self.rpctype_embeds(edge_attr[:, 1])], dim=1)
        for i, conv in enumerate(self.convs[:-1]):
            x = conv(x, edge_index, edge_embeds)
            x = self.bns[i](x)
            x = relu(x, f"bn{i}")
            x = F.dropout(x, p=self.dropout, training=self.training)
        x = self.convs[-1](x, edge_index, edge_embeds)
        local_predict = self.local_linear(x)
        x = x * pattern_probs / pattern_num_nodes
        mean_x = global_add_pool(x, batch)
        g = torch.cat([mean_x, self.entry_embeds(entry_id)], dim=1)
        g = self.global_linear2(relu(self.global_linear1(g), "head"))
        return g, local_predict


def torch_quantile_loss(y_test, y_hat, tau):
    """reference pert_gnn.py:191-193 (pinball loss)."""
    e = y_test - y_hat
    return torch.mean(torch.maximum(tau * e, (tau - 1) * e))


# --------------------------------------------------------------------------
# closed-form backward of one TransformerConv (what the CUDA kernels compute);
# checked against autograd of OracleTransformerConv in tests.
# --------------------------------------------------------------------------
def tconv_backward_closed_form(q, k, v, e, src, dst, alpha, g):
    """Given per-node q,k,v [N,C], per-edge e [E,C], alpha [E] and g = dL/d(out - skip)
    returns (dq, dk, dv, de) following SURVEY.md section 8c.  Pure torch, fp64-capable."""
    C = q.size(1)
    n = q.size(0)
    inv = 1.0 / math.sqrt(C)
    g_i = g.index_select(0, dst)
    q_i = q.index_select(0, dst)
    kj_e = k.index_select(0, src) + e
    vj_e = v.index_select(0, src) + e
    dalpha = (g_i * vj_e).sum(-1)
    dot = scatter(alpha * dalpha, dst, n, "sum").index_select(0, dst)
    ds = alpha * (dalpha - dot)
    dq = scatter(ds.view(-1, 1) * kj_e * inv, dst, n, "sum")
    dk = scatter(ds.view(-1, 1) * q_i * inv, src, n, "sum")
    dv = scatter(alpha.view(-1, 1) * g_i, src, n, "sum")
    de = alpha.view(-1, 1) * g_i + ds.view(-1, 1) * q_i * inv
    return dq, dk, dv, de
