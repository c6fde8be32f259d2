"""Tiny driver for ncu: launches the graph kernels a few times at a BASELINE config's shapes (default cfg2)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pert_gnn_kdd23_b200 import _lib
from pert_gnn_kdd23_b200.data import Batch
from pert_gnn_kdd23_b200.index import build_index
from pert_gnn_kdd23_b200.synthetic import CONFIGS, make_data_list

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
H = CONFIGS[cfg]["hidden"]
b = Batch.from_data_list(make_data_list(cfg)).to("cuda")
N, E = b.x.size(0), b.edge_index.size(1)
gi = build_index(b.edge_index, N, b.edge_attr, 1024, 8)
st = torch.cuda.current_stream().cuda_stream
p = _lib.ptr
planes = torch.randn(4, N, H, device="cuda")
t_if, t_rpc = torch.randn(1024, H, device="cuda"), torch.randn(8, H, device="cuda")
alpha, dsp = torch.empty(E, device="cuda"), torch.empty(E, device="cuda")
out, g = torch.empty(N, H, device="cuda"), torch.randn(N, H, device="cuda")
dpl = torch.empty(4, N, H, device="cuda")
dt_if, dt_rpc = torch.zeros_like(t_if), torch.zeros_like(t_rpc)
rpc_ws = torch.empty(16 * N, device="cuda")
msg = torch.randn(E, H, device="cuda")
for _ in range(reps):
    _lib.call("pert_segment_reduce_fwd", p(msg), p(gi.rowptr), None, p(out), N, H, 1, st)
    _lib.call("pert_tconv_fwd", p(planes[0]), p(planes[1]), p(planes[2]), p(planes[3]), H, p(gi.rowptr), p(gi.csr_src),
              p(gi.csr_if), p(gi.csr_rpc), p(t_if), p(t_rpc), p(out), H, p(alpha), 8, N, E, b.num_graphs, H, st)
    _lib.call("pert_tconv_bwd", p(g), H, p(planes[0]), p(planes[1]), p(planes[2]), H, p(gi.rowptr), p(gi.csr_src),
              p(gi.csr_if), p(gi.csr_rpc), p(gi.colptr), p(gi.csc_pos), p(gi.csc_dst), p(t_if), p(t_rpc), p(alpha),
              p(dpl[0]), p(dpl[1]), p(dpl[2]), H, p(dsp), p(rpc_ws), p(dt_if), p(dt_rpc), 8, N, E, b.num_graphs, H, st)
torch.cuda.synchronize()
print("ok")
