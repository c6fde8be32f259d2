"""CPU tests pinning the oracle: hand-computed known answers, two independent restatements agreeing, fp64
gradcheck, closed-form backward (what the CUDA kernels implement) == autograd of the oracle."""
import math

import numpy as np
import torch

from oracle import loop_oracle, model_oracle
from oracle.model_oracle import OracleSAGEDeterministic, OracleTransformerConv


def _conv(din, c, edge_dim, seed=0, dtype=torch.float64):
    torch.manual_seed(seed)
    return OracleTransformerConv(din, c, heads=1, edge_dim=edge_dim).to(dtype)


def test_kat_two_nodes_one_edge():
    """alpha = 1 for a single incoming edge -> out_1 = v_0 + e + r_1 ; node 0 (no in-edges) -> out_0 = r_0."""
    conv = _conv(3, 4, 2)
    x = torch.tensor([[1.0, 2.0, 3.0], [-1.0, 0.5, 2.0]], dtype=torch.float64)
    ei = torch.tensor([[0], [1]])
    ea = torch.tensor([[0.3, -0.7]], dtype=torch.float64)
    out, alpha = conv(x, ei, ea, return_alpha=True)
    v0 = conv.lin_value(x[0])
    e = conv.lin_edge(ea[0])
    assert torch.allclose(alpha, torch.ones(1, dtype=torch.float64), atol=1e-15)
    assert torch.allclose(out[1], v0 + e + conv.lin_skip(x[1]), atol=1e-12)
    assert torch.allclose(out[0], conv.lin_skip(x[0]), atol=1e-12)


def test_kat_star_equal_logits():
    """identical sources and edge features -> equal logits -> alpha = 1/deg, out = v + e + r."""
    conv = _conv(3, 4, 2, seed=1)
    deg = 5
    x = torch.cat([torch.tensor([[0.2, -0.4, 1.0]]).repeat(deg, 1), torch.tensor([[1.0, 1.0, 1.0]])]).double()
    ei = torch.tensor([list(range(deg)), [deg] * deg])
    ea = torch.tensor([[0.5, 0.25]]).repeat(deg, 1).double()
    out, alpha = conv(x, ei, ea, return_alpha=True)
    assert torch.allclose(alpha, torch.full((deg,), 1.0 / deg, dtype=torch.float64), atol=1e-14)
    assert torch.allclose(out[deg], conv.lin_value(x[0]) + conv.lin_edge(ea[0]) + conv.lin_skip(x[deg]), atol=1e-12)


def test_kat_hand_numbers():
    """Fully hand-computed 1-channel case: weights chosen so q=x, k=2x, v=3x, r=x/2, e=attr."""
    conv = OracleTransformerConv(1, 1, heads=1, edge_dim=1).double()
    with torch.no_grad():
        conv.lin_query.weight.fill_(1.0); conv.lin_query.bias.zero_()
        conv.lin_key.weight.fill_(2.0); conv.lin_key.bias.zero_()
        conv.lin_value.weight.fill_(3.0); conv.lin_value.bias.zero_()
        conv.lin_skip.weight.fill_(0.5); conv.lin_skip.bias.zero_()
        conv.lin_edge.weight.fill_(1.0)
    x = torch.tensor([[1.0], [2.0], [3.0]], dtype=torch.float64)
    ei = torch.tensor([[0, 1], [2, 2]])
    ea = torch.tensor([[0.0], [1.0]], dtype=torch.float64)
    # s_0 = q2*(k0+e0) = 3*(2+0) = 6 ; s_1 = 3*(4+1) = 15 ; C=1
    a0 = math.exp(6 - 15) / (math.exp(6 - 15) + 1 + 1e-16)
    a1 = 1.0 / (math.exp(6 - 15) + 1 + 1e-16)
    expect2 = a0 * (3 + 0) + a1 * (6 + 1) + 1.5
    out, alpha = conv(x, ei, ea, return_alpha=True)
    assert abs(float(alpha[0]) - a0) < 1e-15 and abs(float(alpha[1]) - a1) < 1e-15
    assert abs(float(out[2, 0]) - expect2) < 1e-12
    assert abs(float(out[0, 0]) - 0.5) < 1e-15 and abs(float(out[1, 0]) - 1.0) < 1e-15


def test_very_negative_logits_are_max_shifted():
    conv = _conv(2, 2, 2, seed=3)
    with torch.no_grad():
        for lin in (conv.lin_query, conv.lin_key):
            lin.weight.mul_(300.0)
    x = torch.randn(6, 2, dtype=torch.float64)
    ei = torch.tensor([[0, 1, 2, 3], [5, 5, 5, 5]])
    ea = torch.randn(4, 2, dtype=torch.float64)
    out, alpha = conv(x, ei, ea, return_alpha=True)
    assert torch.isfinite(out).all() and abs(float(alpha.sum()) - 1.0) < 1e-12


def test_duplicate_edges_each_count():
    conv = _conv(3, 4, 2, seed=4)
    x = torch.randn(3, 3, dtype=torch.float64)
    ea = torch.randn(1, 2, dtype=torch.float64)
    o1 = conv(x, torch.tensor([[0], [2]]), ea)
    o2 = conv(x, torch.tensor([[0, 0], [2, 2]]), ea.repeat(2, 1))
    assert torch.allclose(o1, o2, atol=1e-12)       # two identical edges: alpha .5/.5 of the same message
    o3, a3 = conv(x, torch.tensor([[0, 0, 1], [2, 2, 2]]), torch.randn(3, 2, dtype=torch.float64), return_alpha=True)
    assert abs(float(a3.sum()) - 1.0) < 1e-12


def test_loop_oracle_agrees_with_vectorised_oracle():
    rng = np.random.default_rng(0)
    n, E, din, c, de = 7, 15, 3, 4, 2
    conv = _conv(din, c, de, seed=5)
    x = torch.randn(n, din, dtype=torch.float64)
    src = rng.integers(0, n, E)
    dst = rng.integers(0, n - 1, E)        # node n-1 isolated
    ea = torch.randn(E, de, dtype=torch.float64)
    out, alpha = conv(x, torch.tensor(np.stack([src, dst])), ea, return_alpha=True)
    g = lambda t: None if t is None else t.detach().tolist()
    lo, la = loop_oracle.tconv_forward_loops(
        x.tolist(), src.tolist(), dst.tolist(), ea.tolist(),
        g(conv.lin_query.weight), g(conv.lin_query.bias), g(conv.lin_key.weight), g(conv.lin_key.bias),
        g(conv.lin_value.weight), g(conv.lin_value.bias), g(conv.lin_edge.weight),
        g(conv.lin_skip.weight), g(conv.lin_skip.bias))
    assert np.allclose(np.array(lo), out.detach().numpy(), atol=1e-12)
    assert np.allclose(np.array(la), alpha.detach().numpy(), atol=1e-12)


def test_gradcheck_conv_fp64():
    conv = _conv(3, 4, 2, seed=6)
    x = torch.randn(6, 3, dtype=torch.float64, requires_grad=True)
    ei = torch.tensor([[0, 1, 2, 3, 0, 4], [1, 2, 3, 4, 4, 5]])
    ea = torch.randn(6, 2, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(lambda a, b: conv(a, ei, b), (x, ea), eps=1e-6, atol=1e-6)


def test_closed_form_backward_matches_autograd():
    """The formulas the CUDA backward kernels implement == autograd of the oracle forward."""
    torch.manual_seed(7)
    n, E, C = 9, 30, 8
    q, k, v = (torch.randn(n, C, dtype=torch.float64, requires_grad=True) for _ in range(3))
    e = torch.randn(E, C, dtype=torch.float64, requires_grad=True)
    src, dst = torch.randint(0, n, (E,)), torch.randint(0, n - 2, (E,))
    kj = k[src] + e
    s = (q[dst] * kj).sum(-1) / math.sqrt(C)
    alpha = model_oracle.segment_softmax(s, dst, n)
    out = model_oracle.scatter((v[src] + e) * alpha.view(-1, 1), dst, n, "sum")
    g = torch.randn(n, C, dtype=torch.float64)
    dq, dk, dv, de = torch.autograd.grad(out, (q, k, v, e), g)
    cq, ck, cv, ce = model_oracle.tconv_backward_closed_form(q.detach(), k.detach(), v.detach(), e.detach(), src, dst,
                                                             alpha.detach(), g)
    for a, b in ((dq, cq), (dk, ck), (dv, cv), (de, ce)):
        assert torch.allclose(a, b, atol=1e-10)


def test_model_layer_count_quirk_and_state_dict_keys():
    """num_layers 1 and 2 both give 2 convs + 1 bn; 3 -> 3+2; 5 -> 5+4 (reference model.py:24-52)."""
    for nl, nc in ((1, 2), (2, 2), (3, 3), (5, 5)):
        m = OracleSAGEDeterministic(9, [10], 3, 5, 2, 8, nl, 0.0)
        assert len(m.convs) == nc and len(m.bns) == nc - 1
    keys = set(OracleSAGEDeterministic(9, [10], 3, 5, 2, 8, 3, 0.0).state_dict().keys())
    for k in ("convs.0.lin_key.weight", "convs.0.lin_query.bias", "convs.2.lin_edge.weight", "convs.1.lin_skip.bias",
              "bns.1.running_var", "bns.0.num_batches_tracked", "local_linear.weight", "global_linear1.bias",
              "global_linear2.weight", "cat_embedding.0.weight", "entry_embeds.weight", "interface_embeds.weight",
              "rpctype_embeds.weight"):
        assert k in keys
    assert "convs.0.lin_edge.bias" not in keys
    # product model exposes the same keys
    from pert_gnn_kdd23_b200.model import SAGEDeterministic

    assert set(SAGEDeterministic(9, [10], 3, 5, 2, 8, 3, 0.0).state_dict().keys()) == keys


def test_pinball_loss():
    y = torch.tensor([10.0, 20.0, 30.0])
    yh = torch.tensor([12.0, 20.0, 25.0])
    # e = [-2, 0, 5]; tau=.9 -> max(.9e, -.1e) = [.2, 0, 4.5]
    assert abs(float(model_oracle.torch_quantile_loss(y, yh, 0.9)) - (0.2 + 0 + 4.5) / 3) < 1e-6
