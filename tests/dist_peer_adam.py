"""2+ GPU check (run under torchrun, one rank per GPU): PeerAdam (fused all-reduce + Adam over peer memory) ==
NCCL all-reduce + FusedAdam after several data-parallel steps, and the replicas stay bit-identical.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_peer_adam.py
Prints PEER_ADAM_OK on rank 0 on success."""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from pert_gnn_kdd23_b200.data import Batch
from pert_gnn_kdd23_b200.model import SAGEDeterministic
from pert_gnn_kdd23_b200.synthetic import make_data_list, model_args
from pert_gnn_kdd23_b200.train import (DataParallel, FlatParams, FusedAdam, GraphedTrainStep, PeerAdam,
                                       fused_train_step)


def main():
    rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr_)
    dev = torch.device("cuda", lr_)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model_a = SAGEDeterministic(*model_args(1)).to(dev)
    model_b = copy.deepcopy(model_a)
    fa, fb = FlatParams(model_a), FlatParams(model_b)
    opt_a, opt_b = FusedAdam(fa, lr=1e-3), PeerAdam(fb, lr=1e-3)
    dp_a, dp_b = DataParallel(fa), DataParallel(fb)
    dl = make_data_list(1)
    per = 16
    batches = [Batch.from_data_list(dl[(2 * rank + k) * per:(2 * rank + k + 1) * per]).to(dev) for k in range(2)]
    gstep = GraphedTrainStep(model_b, opt_b, 0.5, dp_b)
    for it in range(8):
        b = batches[it % 2]
        la = fused_train_step(model_a, opt_a, b, 0.5, dp_a)
        lb = gstep(b)
        assert abs(float(la) - float(lb)) <= 2e-3 * max(1.0, abs(float(la))), (it, float(la), float(lb))
    opt_b.check()
    assert gstep.capture_error is None, gstep.capture_error
    worst = 0.0
    for (n, pa), (_, pb) in zip(model_a.named_parameters(), model_b.named_parameters()):
        if n.endswith("lin_key.bias") or (n.endswith("lin_skip.bias") and not n.startswith("convs.1.")):
            continue
        d = (pa - pb).abs().max().item() / max(pa.abs().max().item(), 1e-6)
        worst = max(worst, d)
    assert worst < 2e-3, worst
    # replicas bit-identical: every rank applied the same rank-ordered sum
    flat = fb.flat.clone()
    ref = flat.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(flat, ref), "PeerAdam replicas diverged"
    phases = opt_b.phase_times_us()
    opt_b.close()
    # drop-in loop under data parallelism: torch.optim.Adam + DataParallel.all_reduce_module_grads (one all-reduce of the
    # engine's flat gradient buffer) keeps the replicas bit-identical
    from pert_gnn_kdd23_b200.train import train_step

    torch.manual_seed(1)
    model_c = SAGEDeterministic(*model_args(1)).to(dev)
    for p_ in model_c.parameters():
        dist.broadcast(p_.data, 0)
    dp_c = DataParallel(FlatParams(model_c, bind_grads=False))
    opt_c = torch.optim.Adam(model_c.parameters(), lr=1e-3)
    for it in range(4):
        train_step(model_c, opt_c, batches[it % 2], 0.5, dp_c)
    flat_c = torch.cat([p_.detach().reshape(-1) for p_ in model_c.parameters()])
    ref_c = flat_c.clone()
    dist.broadcast(ref_c, 0)
    assert torch.equal(flat_c, ref_c), "drop-in DP replicas diverged"
    # a model on this rank's device while ANOTHER device is current (device guard of the bindings)
    other = (lr_ + 1) % torch.cuda.device_count()
    torch.cuda.set_device(other)
    gp, _ = model_c(batches[0].x, batches[0].cat_X, batches[0].edge_index, batches[0].edge_attr,
                    batches[0].pattern_num_nodes, batches[0].rt_probs, batches[0].entry_id, batches[0].batch)
    assert gp.device == dev and torch.isfinite(gp).all()
    torch.cuda.set_device(lr_)
    dist.barrier()
    if rank == 0:
        print(f"PEER_ADAM_OK world={world} worst_rel={worst:.2e} replays={gstep.replays} phases_us={phases} "
              f"dropin_dp_replicas_identical=True device_guard=True")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
