"""Per-kernel micro-benchmarks through the C-ABI at a BASELINE config's shapes (default cfg2):
CUDA-event timing, median of `iters`, cold (512 MB L2 flush before every launch) and warm (back to back).
    python profiles/kbench.py [cfg] [out.json]
Not a bench value: this is the optimisation workbench; bench.py is the contract."""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pert_gnn_kdd23_b200 import _lib, ops
from pert_gnn_kdd23_b200.data import Batch
from pert_gnn_kdd23_b200.index import build_index
from pert_gnn_kdd23_b200.synthetic import CONFIGS, make_data_list

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
out_path = sys.argv[2] if len(sys.argv) > 2 else None
H = CONFIGS[cfg]["hidden"]
PEAK = 6580.9
if os.path.exists("MEASURED_PEAKS.json"):
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]

b = Batch.from_data_list(make_data_list(cfg)).to("cuda")
N, E = b.x.size(0), b.edge_index.size(1)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream()


def timeit(fn, iters=40, cold=True):
    ts = []
    for i in range(iters + 3):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        fn()
        e1.record(st)
        e1.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


results = {}


def report(name, fn, alg_bytes=None, flops=None):
    cold, warm = timeit(fn, cold=True), timeit(fn, cold=False)
    r = {"us_cold": round(cold, 2), "us_warm": round(warm, 2)}
    if alg_bytes:
        r["alg_MB"] = round(alg_bytes / 1e6, 2)
        r["GBs_cold"] = round(alg_bytes / cold / 1e3, 1)
        r["frac_cold"] = round(alg_bytes / cold / 1e3 / PEAK, 3)
        r["frac_warm"] = round(alg_bytes / warm / 1e3 / PEAK, 3)
    if flops:
        r["TFLOPs_cold"] = round(flops / cold / 1e6, 2)
        r["TFLOPs_warm"] = round(flops / warm / 1e6, 2)
    results[name] = r
    print(name, r, flush=True)


gi = build_index(b.edge_index, N, b.edge_attr, 1024, 8)
report("build_index", lambda: build_index(b.edge_index, N, b.edge_attr, 1024, 8), alg_bytes=40 * E + 16 * E)

# ---- segment reduce (metric kernel)
msg = torch.randn(E, H, device="cuda")
outn = torch.empty(N, H, device="cuda")
for op, nm in ((1, "max"), (0, "sum")):
    report(f"segreduce_{nm}",
           lambda: _lib.call("pert_segment_reduce_fwd", msg.data_ptr(), gi.rowptr.data_ptr(), None, outn.data_ptr(), N,
                             H, op, st.cuda_stream),
           alg_bytes=4 * E * H + 4 * (N + 1) + 4 * N * H)

# ---- fused conv
planes = torch.randn(4, N, H, device="cuda")
t_if = torch.randn(1024, H, device="cuda")
t_rpc = torch.randn(8, H, device="cuda")
alpha = torch.empty(E, device="cuda")
p = _lib.ptr
report("tconv_fwd",
       lambda: _lib.call("pert_tconv_fwd", p(planes[0]), p(planes[1]), p(planes[2]), p(planes[3]), H, p(gi.rowptr),
                         p(gi.csr_src), p(gi.csr_if), p(gi.csr_rpc), p(t_if), p(t_rpc), p(outn), H, p(alpha), 8, N, E, b.num_graphs, H,
                         st.cuda_stream),
       alg_bytes=16 * N * H + 12 * E + 4 * (N + 1) + 4 * E)
g = torch.randn(N, H, device="cuda")
dpl = torch.empty(4, N, H, device="cuda")
dsp = torch.empty(E, device="cuda")
dt_if, dt_rpc = torch.zeros_like(t_if), torch.zeros_like(t_rpc)
report("tconv_bwd(dst+src)",
       lambda: _lib.call("pert_tconv_bwd", p(g), H, p(planes[0]), p(planes[1]), p(planes[2]), H, p(gi.rowptr),
                         p(gi.csr_src), p(gi.csr_if), p(gi.csr_rpc), p(gi.colptr), p(gi.csc_pos), p(gi.csc_dst),
                         p(t_if), p(t_rpc), p(alpha), p(dpl[0]), p(dpl[1]), p(dpl[2]), H, p(dsp), p(dt_if), p(dt_rpc),
                         8, N, E, b.num_graphs, H, st.cuda_stream),
       alg_bytes=2 * (16 * N * H + 12 * E + 4 * (N + 1) + 8 * E))

# ---- node linears (fwd / dgrad / wgrad) at layer>0 shape K=H and layer-0 shape K=k0
for K in (H, (9 + H + 7) // 8 * 8):
    x = torch.randn(N, K, device="cuda")
    W = torch.randn(4 * H, K, device="cuda") / 8
    bias = torch.randn(4 * H, device="cuda")
    y = torch.empty(4, N, H, device="cuda")
    fl = 2.0 * N * K * 4 * H
    report(f"linear_fwd_K{K}", lambda: ops.gemm_nt_raw(x, W, bias, y), alg_bytes=4 * N * K + 16 * N * H, flops=fl)
    Wt = W.t().contiguous()
    dx = torch.empty(N, K, device="cuda")
    report(f"linear_dgrad_K{K}", lambda: ops.gemm_nt_raw(y, Wt, None, dx), alg_bytes=4 * N * K + 16 * N * H, flops=fl)
    dW = torch.zeros(4 * H, K, device="cuda")
    report(f"linear_wgrad_K{K}", lambda: ops.gemm_tn_raw(y, x, dW), alg_bytes=4 * N * K + 16 * N * H, flops=fl)
    db = torch.zeros(4 * H, device="cuda")
    report(f"linear_bgrad_K{K}", lambda: ops.colsum_raw(y, db), alg_bytes=16 * N * H)

# ---- batch norm + pool
xo = torch.randn(N, H, device="cuda")
gam, bet = torch.ones(H, device="cuda"), torch.zeros(H, device="cuda")
rm, rv = torch.zeros(H, device="cuda"), torch.ones(H, device="cuda")
nbt = torch.zeros(1, dtype=torch.long, device="cuda")
report("bn_relu_fwd", lambda: ops.batch_norm(xo, gam, bet, rm, rv, nbt, True, relu=True), alg_bytes=3 * 4 * N * H)
xr = xo.clone().requires_grad_()
yb = ops.batch_norm(xr, gam.requires_grad_(), bet.requires_grad_(), rm, rv, nbt, True, relu=True)
report("bn_relu_bwd", lambda: torch.autograd.grad(yb, xr, g, retain_graph=True), alg_bytes=4 * 4 * N * H)

if out_path:
    json.dump({"cfg": cfg, "N": N, "E": E, "H": H, "peak_GBs": PEAK, "kernels": results}, open(out_path, "w"), indent=1)
