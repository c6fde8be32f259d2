"""BatchNorm(+ReLU) forward/backward of libpertgnn at large N against torch fp64 on the GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pert_gnn_kdd23_b200 import ops

def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

for N, H in ((48459, 128), (102400, 128), (51200, 64), (256000, 128)):
    torch.manual_seed(N)
    x = (torch.randn(N, H, device="cuda") * 0.7 + 0.3)
    g = torch.randn(N, H, device="cuda") * (torch.rand(N, 1, device="cuda") ** 4)   # very uneven row magnitudes
    gamma = torch.rand(H, device="cuda") + 0.5
    beta = torch.randn(H, device="cuda") * 0.1
    xd = x.double().requires_grad_()
    gd, bd = gamma.double().requires_grad_(), beta.double().requires_grad_()
    yd = torch.relu(torch.nn.functional.batch_norm(xd, None, None, gd, bd, True, 0.1, 1e-5))
    yd.backward(g.double())
    xc = x.clone().requires_grad_()
    gc, bc = gamma.clone().requires_grad_(), beta.clone().requires_grad_()
    rm, rv, nbt = torch.zeros(H, device="cuda"), torch.ones(H, device="cuda"), torch.zeros((), dtype=torch.long, device="cuda")
    yc = ops.batch_norm(xc, gc, bc, rm, rv, nbt, True, 1e-5, 0.1, relu=True)
    yc.backward(g)
    print(f"N={N} H={H}: y {rel(yc, yd):.2e} dx {rel(xc.grad, xd.grad):.2e} dgamma {rel(gc.grad, gd.grad):.2e} dbeta {rel(bc.grad, bd.grad):.2e}")
