"""Host-side breakdown of the UNCHANGED reference loop body (bench.py `e2e_dropin`, pert_gnn.py:219-250) at cfg2:
perf_counter around each statement without extra synchronisation (the step's own loss.item() is the only sync), so the
numbers are the CPU issue cost of every phase; the last column is the wait inside item() = GPU work still outstanding.
Usage (GPU box): python profiles/prof_dropin.py [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pert_gnn_kdd23_b200.data import Batch
from pert_gnn_kdd23_b200.model import SAGEDeterministic
from pert_gnn_kdd23_b200.synthetic import make_data_list, model_args
from pert_gnn_kdd23_b200.train import model_inputs, torch_quantile_loss


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    dev = torch.device("cuda")
    hbs = []
    for r in range(4):
        dl = make_data_list(2, num_graphs=256, seed=1002 + r)
        for d in dl:
            d._store.pop("level", None)
            d._store.pop("min_depth", None)
        hbs.append(Batch.from_data_list(dl).pin_memory())
    torch.manual_seed(0)
    model = SAGEDeterministic(*model_args(2)).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=3e-4)
    names = ["to", "zero_grad", "forward", "loss", "backward", "opt.step", "item"]
    acc = [0.0] * len(names)

    def step(hb, rec):
        t = [time.perf_counter()]
        data = hb.to(dev, non_blocking=True); t.append(time.perf_counter())
        opt.zero_grad(); t.append(time.perf_counter())
        gp, _ = model(*model_inputs(data)); t.append(time.perf_counter())
        l = torch_quantile_loss(data.y.float(), gp.flatten(), 0.5); t.append(time.perf_counter())
        l.backward(); t.append(time.perf_counter())
        opt.step(); t.append(time.perf_counter())
        v = l.item(); t.append(time.perf_counter())
        if rec:
            for i in range(len(names)):
                acc[i] += t[i + 1] - t[i]
        return v

    for i in range(10):
        step(hbs[i % 4], False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        step(hbs[i % 4], True)
    torch.cuda.synchronize()
    tot = (time.perf_counter() - t0) / steps
    print(f"step {tot * 1e6:.0f} us; per phase (host issue, us):",
          {n: round(a / steps * 1e6, 1) for n, a in zip(names, acc)})
    # the same under the profiler: top CPU ops
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for i in range(10):
            step(hbs[i % 4], False)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=22, max_name_column_width=60))


if __name__ == "__main__":
    main()
