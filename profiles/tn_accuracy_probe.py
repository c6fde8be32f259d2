"""Accuracy probe of the tensor-core weight-gradient GEMM (pert_gemm_tn) against fp64, per plane and per column block,
for operands shaped like conv 0 at H = 128: A = dplanes [R, 4H] (planes of very different magnitude), B = X0 [R, 144]
(embedding columns ~N(0,1), feature columns U(0,1), zero pad).  Prints max|err| / max|ref| per (plane, column block) and
the same for a fp32 torch matmul on the GPU (round-to-nearest FMA reference)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pert_gnn_kdd23_b200 import _lib


def run(R, H, Nc, scales, centered_a=False, seed=0):
    torch.manual_seed(seed)
    A = torch.randn(4, R, H)
    for p, s in enumerate(scales):
        A[p] *= s
    if centered_a:      # columns of the q/k planes sum to ~0 like dk does (softmax shift invariance)
        A[:2] -= A[:2].mean(dim=1, keepdim=True)
    B = torch.zeros(R, Nc)
    B[:, :H] = torch.randn(R, H)
    nf = min(9, Nc - H)
    if nf > 0:
        B[:, H:H + nf] = torch.rand(R, nf)
    Ad, Bd = A.double().permute(1, 0, 2).reshape(R, 4 * H), B.double()
    ref = Ad.t() @ Bd
    Ac, Bc = A.cuda(), B.cuda()
    C = torch.zeros(4 * H, Nc, device="cuda")
    cs = torch.zeros(4 * H, device="cuda")
    _lib.call("pert_gemm_tn", Ac.data_ptr(), H, H, R * H, Bc.data_ptr(), Nc, 0, 0, C.data_ptr(), Nc, cs.data_ptr(), R,
              4 * H, Nc, torch.cuda.current_stream().cuda_stream)
    f32 = (Ac.permute(1, 0, 2).reshape(R, 4 * H).t() @ Bc).cpu().double()
    C = C.cpu().double()
    print(f"R={R} H={H} Nc={Nc} scales={scales} centered={centered_a}")
    for p in range(4):
        rows = slice(p * H, (p + 1) * H)
        for name, cols in (("emb", slice(0, H)), ("feat", slice(H, Nc))):
            if cols.stop <= cols.start:
                continue
            r = ref[rows, cols]
            den = r.abs().max().clamp_min(1e-300)
            e_tc = float((C[rows, cols] - r).abs().max() / den)
            e_32 = float((f32[rows, cols] - r).abs().max() / den)
            print(f"  plane {p} cols {name:4s}: tcgen05 {e_tc:.2e}   torch fp32 {e_32:.2e}   max|ref| {float(den):.2e}")


if __name__ == "__main__":
    run(51200, 128, 144, (1e-4, 1e-4, 1.0, 1.0))
    run(51200, 128, 144, (1e-4, 1e-4, 1.0, 1.0), centered_a=True)
    run(51200, 128, 128, (1e-4, 1e-4, 1.0, 1.0), centered_a=True)
    run(51200, 64, 80, (1e-4, 1e-4, 1.0, 1.0), centered_a=True)
