"""A few eager train steps of the benchmark workload for ncu launch lists / --set full captures:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python profiles/prof_step.py [cfg] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pert_gnn_kdd23_b200.data import Batch
from pert_gnn_kdd23_b200.model import SAGEDeterministic
from pert_gnn_kdd23_b200.synthetic import make_data_list, model_args
from pert_gnn_kdd23_b200.train import FlatParams, FusedAdam, fused_train_step

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ng = int(sys.argv[3]) if len(sys.argv) > 3 else None
torch.manual_seed(0)
model = SAGEDeterministic(*model_args(cfg)).cuda()
opt = FusedAdam(FlatParams(model), lr=3e-4)
b = Batch.from_data_list(make_data_list(cfg, num_graphs=ng)).to("cuda")
for _ in range(steps):
    fused_train_step(model, opt, b, 0.5)
torch.cuda.synchronize()
print("done")
