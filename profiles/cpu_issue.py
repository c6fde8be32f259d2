"""How far ahead of the GPU does the host run?  Wall time to ENQUEUE K fused train steps vs. time until they finish."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pert_gnn_kdd23_b200.data import Batch
from pert_gnn_kdd23_b200.model import SAGEDeterministic
from pert_gnn_kdd23_b200.synthetic import make_data_list, model_args
from pert_gnn_kdd23_b200.train import FlatParams, FusedAdam, fused_train_step

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
K = int(sys.argv[2]) if len(sys.argv) > 2 else 50
torch.manual_seed(0)
model = SAGEDeterministic(*model_args(cfg)).cuda()
opt = FusedAdam(FlatParams(model))
b = Batch.from_data_list(make_data_list(cfg)).to("cuda")
for _ in range(5):
    fused_train_step(model, opt, b)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(K):
    fused_train_step(model, opt, b)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"enqueue {1e3 * (t1 - t0) / K:.3f} ms/step, complete {1e3 * (t2 - t0) / K:.3f} ms/step")

# same loop through the CUDA-graph replay path
from pert_gnn_kdd23_b200.train import GraphedTrainStep  # noqa: E402

gs = GraphedTrainStep(model, opt)
for _ in range(5):
    gs(b)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(K):
    gs(b)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"graphed: enqueue {1e3 * (t1 - t0) / K:.3f} ms/step, complete {1e3 * (t2 - t0) / K:.3f} ms/step, "
      f"replays {gs.replays}, capture_error {gs.capture_error}")
