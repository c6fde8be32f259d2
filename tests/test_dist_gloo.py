"""world_size-2 gloo test (CPU) of the data-parallel plumbing: flat-buffer gradient all-reduce == gradient of the
union batch when the loss is the mean over shards (no BN coupling: eval-mode BN)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.model_oracle import OracleSAGEDeterministic, torch_quantile_loss
    from pert_gnn_kdd23_b200.data import Batch
    from pert_gnn_kdd23_b200.synthetic import make_data_list, model_args
    from pert_gnn_kdd23_b200.train import DataParallel, FlatParams, model_inputs

    torch.manual_seed(0)
    model = OracleSAGEDeterministic(*model_args(1)).eval()
    fp = FlatParams(model)
    dp = DataParallel(fp)
    dl = make_data_list(1, 8)
    shard = Batch.from_data_list(dl[rank * 4:(rank + 1) * 4])
    fp.zero_grad()
    g, _ = model(*model_inputs(shard))
    torch_quantile_loss(shard.y.float(), g.flatten(), 0.5).backward()
    scale = dp.all_reduce_grads()
    grad = fp.grad * scale
    if rank == 0:
        torch.save(grad, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_flat_allreduce_equals_union_batch_gradient(tmp_path):
    out = str(tmp_path / "g.pt")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    from oracle.model_oracle import OracleSAGEDeterministic, torch_quantile_loss
    from pert_gnn_kdd23_b200.data import Batch
    from pert_gnn_kdd23_b200.synthetic import make_data_list, model_args
    from pert_gnn_kdd23_b200.train import FlatParams, model_inputs

    torch.manual_seed(0)
    model = OracleSAGEDeterministic(*model_args(1)).eval()
    fp = FlatParams(model)
    b = Batch.from_data_list(make_data_list(1, 8))
    g, _ = model(*model_inputs(b))
    torch_quantile_loss(b.y.float(), g.flatten(), 0.5).backward()
    got = torch.load(out)
    assert torch.allclose(got, fp.grad, rtol=1e-4, atol=1e-7)


def _worker_torch_opt(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.model_oracle import OracleSAGEDeterministic
    from pert_gnn_kdd23_b200.data import Batch
    from pert_gnn_kdd23_b200.synthetic import make_data_list, model_args
    from pert_gnn_kdd23_b200.train import DataParallel, FlatParams, train_step

    torch.manual_seed(0)
    model = OracleSAGEDeterministic(*model_args(1)).eval()
    dp = DataParallel(FlatParams(model))
    opt = torch.optim.SGD(model.parameters(), lr=0.0)      # zero_grad() defaults to set_to_none=True: unbinds fp.grad
    dl = make_data_list(1, 8)
    shard = Batch.from_data_list(dl[rank * 4:(rank + 1) * 4])
    for _ in range(2):                                     # second step: p.grad are fresh tensors, not views of fp.grad
        train_step(model, opt, shard, 0.5, dp)
    if rank == 0:
        torch.save({n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_train_step_torch_optimizer_averages_the_attached_gradients(tmp_path):
    """the torch-optimizer + DataParallel branch must reduce the gradients attached to the parameters
    (torch.optim's zero_grad(set_to_none=True) drops the views FlatParams bound), not a stale flat buffer."""
    out = str(tmp_path / "g2.pt")
    port = 31500 + os.getpid() % 2000
    mp.spawn(_worker_torch_opt, args=(2, port, out), nprocs=2, join=True)
    from oracle.model_oracle import OracleSAGEDeterministic, torch_quantile_loss
    from pert_gnn_kdd23_b200.data import Batch
    from pert_gnn_kdd23_b200.synthetic import make_data_list, model_args
    from pert_gnn_kdd23_b200.train import model_inputs

    torch.manual_seed(0)
    model = OracleSAGEDeterministic(*model_args(1)).eval()
    b = Batch.from_data_list(make_data_list(1, 8))
    g, _ = model(*model_inputs(b))
    torch_quantile_loss(b.y.float(), g.flatten(), 0.5).backward()
    got = torch.load(out)
    for n, p in model.named_parameters():
        if p.grad is not None:
            assert torch.allclose(got[n], p.grad, rtol=1e-4, atol=1e-7), n
