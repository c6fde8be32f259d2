"""Second, independent restatement of one TransformerConv layer: scalar Python
loops over python floats (double precision), no torch ops, no code shared with
oracle/model_oracle.py.  TEST INFRASTRUCTURE; small cases only.

Follows SURVEY.md section 8c formula block (PyG 2.4.0 TransformerConv,
heads=1, edge_dim set, root_weight=True, aggr='add'):
    s_t  = <q_i, k_j + e_t> / sqrt(C)
    m_i  = max over incoming edges (0 when the node has none)
    p_t  = exp(s_t - m_i);  Z_i = sum p_t + 1e-16;  alpha_t = p_t / Z_i
    out_i = sum alpha_t (v_j + e_t) + r_i
"""
import math


def _linear(W, b, x):
    out = []
    for o in range(len(W)):
        acc = 0.0 if b is None else float(b[o])
        for i in range(len(x)):
            acc += float(W[o][i]) * float(x[i])
        out.append(acc)
    return out


def tconv_forward_loops(x, edge_src, edge_dst, edge_feat, Wq, bq, Wk, bk, Wv, bv, We, Ws, bs):
    """x: list[N][Din]; edge_feat: list[E][De]; W*: list[C][.]; returns (out[N][C], alpha[E])."""
    n = len(x)
    C = len(Wq)
    q = [_linear(Wq, bq, xi) for xi in x]
    k = [_linear(Wk, bk, xi) for xi in x]
    v = [_linear(Wv, bv, xi) for xi in x]
    r = [_linear(Ws, bs, xi) for xi in x]
    e = [_linear(We, None, ef) for ef in edge_feat]
    E = len(edge_src)
    s = []
    for t in range(E):
        i, j = edge_dst[t], edge_src[t]
        s.append(sum(q[i][c] * (k[j][c] + e[t][c]) for c in range(C)) / math.sqrt(C))
    m = [None] * n
    for t in range(E):
        i = edge_dst[t]
        m[i] = s[t] if m[i] is None else max(m[i], s[t])
    m = [0.0 if mi is None else mi for mi in m]
    p = [math.exp(s[t] - m[edge_dst[t]]) for t in range(E)]
    Z = [1e-16] * n
    for t in range(E):
        Z[edge_dst[t]] += p[t]
    alpha = [p[t] / Z[edge_dst[t]] for t in range(E)]
    out = [list(ri) for ri in r]
    for t in range(E):
        i, j = edge_dst[t], edge_src[t]
        for c in range(C):
            out[i][c] += alpha[t] * (v[j][c] + e[t][c])
    return out, alpha
