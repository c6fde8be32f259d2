"""ncu driver: the three node-linear GEMMs (fwd / dgrad / wgrad) at a BASELINE config's shape."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pert_gnn_kdd23_b200 import ops
from pert_gnn_kdd23_b200.synthetic import CONFIGS

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
c = CONFIGS[cfg]
H = c["hidden"]
N = c["graphs"] * (c["nodes"] or 48)
K = H
x = torch.randn(N, K, device="cuda")
W = torch.randn(4 * H, K, device="cuda") / 8
bias = torch.randn(4 * H, device="cuda")
y = torch.empty(4, N, H, device="cuda")
Wt = W.t().contiguous()
dx = torch.empty(N, K, device="cuda")
dW = torch.zeros(4 * H, K, device="cuda")
for _ in range(3):
    ops.gemm_nt_raw(x, W, bias, y)
    ops.gemm_nt_raw(y, Wt, None, dx)
    ops.gemm_tn_raw(y, x, dW)
torch.cuda.synchronize()
print("ok")
