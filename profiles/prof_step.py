"""Short driver for ncu captures: 3 train steps of the BASELINE cfg (default cfg2) through the public API."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pert_gnn_kdd23_b200.data import Batch
from pert_gnn_kdd23_b200.model import SAGEDeterministic
from pert_gnn_kdd23_b200.synthetic import make_data_list, model_args
from pert_gnn_kdd23_b200.train import FlatParams, FusedAdam, fused_train_step as train_step

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
model = SAGEDeterministic(*model_args(cfg)).cuda()
fp = FlatParams(model)
opt = FusedAdam(fp)
b = Batch.from_data_list(make_data_list(cfg)).to("cuda")
for _ in range(steps):
    loss = train_step(model, opt, b)
torch.cuda.synchronize()
print("loss", float(loss))
