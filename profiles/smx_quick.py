import sys, json
sys.path.insert(0, '/root/repo')
import bench, torch
from pert_gnn_kdd23_b200.synthetic import CONFIGS
bs = bench.make_batches(2, 0, 1)
b = bs[0].to('cuda')
peak, _ = bench._peaks()
for _ in range(3):
    r = bench.scatter_max_bench(b, 64, peak)
    print(round(r['us_per_launch'],2), round(r['frac'],4), round(r['us_single_launch_after_l2_flush'],2))
