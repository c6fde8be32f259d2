"""Eager fused train steps on a PERT-exact batch (span rows -> GPU PERT graphs -> pattern store -> device-collated batch of
256 traces), for an ncu launch list of the step on real-data-shaped graphs (interface = rpctype = 0 on 3 of 4 edges, runs of
equal cat_X).  Usage: ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python profiles/prof_pert_step.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pert_gnn_kdd23_b200.model import SAGEDeterministic
from pert_gnn_kdd23_b200.store import PatternStore
from pert_gnn_kdd23_b200.synthetic import make_pert_artifacts, model_args
from pert_gnn_kdd23_b200.train import FlatParams, FusedAdam, fused_train_step


def main():
    warm, steps = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2, 3)
    art, info = make_pert_artifacts(seed=3, n_patterns=256, n_entries=64, n_traces=4096, device="cuda")
    store = PatternStore.from_artifacts(art, "cuda")
    batch = store.assemble(list(range(256)))
    torch.manual_seed(0)
    model = SAGEDeterministic(*model_args(2)).cuda()
    opt = FusedAdam(FlatParams(model), lr=1e-3)
    for _ in range(warm):
        fused_train_step(model, opt, batch, 0.5)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("steps")
    for _ in range(steps):
        fused_train_step(model, opt, batch, 0.5)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    print("nodes", batch.x.size(0), "edges", batch.edge_index.size(1), info)


if __name__ == "__main__":
    main()
