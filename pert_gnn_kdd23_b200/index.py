"""Per-batch graph index (CSR by target + CSC by source) built on the GPU.

PyG keeps graphs in COO and lets every ``MessagePassing.propagate`` call scatter
over it (reference model.py:100,104).  Here the batch's ``edge_index`` is turned,
once, into the int32 layout the fused kernels stream over (csrc/index.cu;
numpy definition: oracle/index_oracle.py:build_index) and reused by all layers
and by backward.
"""
from __future__ import annotations

import os
import weakref

import torch

from . import _lib

_CHECK = os.environ.get("PERT_CHECK_INDICES", "0") == "1"


class GraphIndex:
    @property
    def device(self):
        return self._buf.device

    __slots__ = ("N", "E", "rowptr", "perm", "csr_src", "csr_if", "csr_rpc", "colptr", "csc_pos",
                 "csc_dst", "status", "has_attr", "n_if", "n_rpc", "num_graphs", "_buf", "__weakref__")

    def check(self):
        """Synchronising validity check of the ids seen while building (debug aid)."""
        code = int(self.status.item())
        if code != 0:
            _lib.check(code, "graph index")
        return self


def _al(n, a=64):
    return (n + a - 1) // a * a


@_lib.on_device_of
def build_index(edge_index, num_nodes, edge_attr=None, n_if=0, n_rpc=0, check=None):
    """edge_index int64 [2,E] on cuda; edge_attr int64 [E,>=2] (cols 0,1 = interface, rpctype) or None."""
    if not edge_index.is_cuda:
        raise _lib.PertGnnError("build_index needs CUDA tensors (no CPU fallback for the hot path)")
    assert edge_index.dtype == torch.int64 and edge_index.dim() == 2 and edge_index.size(0) == 2
    ei = edge_index.contiguous()
    N, E = int(num_nodes), int(ei.size(1))
    dev = ei.device
    ea = None
    cols = 0
    if edge_attr is not None:
        assert edge_attr.dtype == torch.int64 and edge_attr.dim() == 2 and edge_attr.size(1) >= 2
        ea = edge_attr.contiguous()
        cols = int(ea.size(1))
    # one int32 slab for all outputs (64-element aligned slices => 256-byte aligned)
    sizes = [N + 1, E, E, E, E, N + 1, E, E, 16]
    offs = [0]
    for s in sizes:
        offs.append(offs[-1] + _al(max(s, 1)))
    buf = torch.empty(offs[-1], dtype=torch.int32, device=dev)
    parts = [buf[offs[i]:offs[i] + sizes[i]] for i in range(len(sizes))]
    gi = GraphIndex()
    gi.N, gi.E, gi._buf = N, E, buf
    gi.num_graphs = 0          # hint for the shared-memory tile size (set by the model: batch size)
    (gi.rowptr, gi.perm, gi.csr_src, gi.csr_if, gi.csr_rpc, gi.colptr, gi.csc_pos, gi.csc_dst, status) = parts
    gi.status = status[:1]
    gi.status.zero_()
    gi.has_attr = ea is not None
    gi.n_if, gi.n_rpc = int(n_if), int(n_rpc)
    wbytes = _lib.lib().pert_index_workspace_bytes(N, E)
    ws = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    _lib.call("pert_build_index", _lib.ptr(ei), _lib.ptr(ea), cols, N, E, gi.n_if, gi.n_rpc,
              _lib.ptr(gi.rowptr), _lib.ptr(gi.perm), _lib.ptr(gi.csr_src),
              _lib.ptr(gi.csr_if) if ea is not None else None, _lib.ptr(gi.csr_rpc) if ea is not None else None,
              _lib.ptr(gi.colptr), _lib.ptr(gi.csc_pos), _lib.ptr(gi.csc_dst), _lib.ptr(ws), wbytes,
              _lib.ptr(gi.status), _lib.stream())
    if _CHECK if check is None else check:
        gi.check()
    return gi


@_lib.on_device_of
def graph_ptr(batch, num_graphs):
    """int32 ptr [B+1] from a PyG ``batch`` vector (cuda int64)."""
    assert batch.is_cuda and batch.dtype == torch.int64
    B = int(num_graphs)
    ptr = torch.empty(B + 1, dtype=torch.int32, device=batch.device)
    ws = torch.empty(1 << 16, dtype=torch.uint8, device=batch.device)
    _lib.call("pert_graph_ptr", _lib.ptr(batch.contiguous()), batch.numel(), B, _lib.ptr(ptr), _lib.ptr(ws),
              ws.numel(), None, _lib.stream())
    return ptr


@_lib.on_device_of
def min_depth(gptr, index, roots):
    """Level index: min hop depth from ``roots[g]`` (global node ids, int32 [B]) per graph; -1 unreachable."""
    B = gptr.numel() - 1
    depth = torch.empty(index.N, dtype=torch.int32, device=gptr.device)
    _lib.call("pert_min_depth", _lib.ptr(gptr), B, _lib.ptr(index.colptr), _lib.ptr(index.csc_dst),
              _lib.ptr(roots.to(torch.int32).contiguous()), _lib.ptr(depth), _lib.stream())
    return depth


@_lib.on_device_of
def node_depth(gptr, depth):
    """The tensor the reference stores as ``Data.node_depth`` ([N,1] int64, misc.py:159-175 + the long cast of
    :215/:368) from the raw min-depth of ``min_depth``."""
    B = gptr.numel() - 1
    out = torch.empty(depth.numel(), 1, dtype=torch.int64, device=depth.device)
    _lib.call("pert_node_depth", _lib.ptr(gptr), B, _lib.ptr(depth.contiguous()), _lib.ptr(out), _lib.stream())
    return out


@_lib.on_device_of
def level_order(gptr, depth):
    """Level-major node order inside each graph (int32 [N]; oracle/index_oracle.py:level_order)."""
    B = gptr.numel() - 1
    out = torch.empty(depth.numel(), dtype=torch.int32, device=depth.device)
    _lib.call("pert_level_order", _lib.ptr(gptr), B, _lib.ptr(depth.contiguous()), _lib.ptr(out), _lib.stream())
    return out


# ---- small cache so the 2..5 conv layers of one forward (and repeated calls on the same batch) share one index
_cache = {}


def cached_index(edge_index, num_nodes, edge_attr, n_if, n_rpc):
    key = (edge_index.data_ptr(), edge_index._version, tuple(edge_index.shape), int(num_nodes),
           None if edge_attr is None else (edge_attr.data_ptr(), edge_attr._version, tuple(edge_attr.shape)),
           n_if, n_rpc)
    hit = _cache.get("k")
    if hit is not None and hit[0] == key and hit[1]() is edge_index:
        return hit[2]
    gi = build_index(edge_index, num_nodes, edge_attr, n_if, n_rpc)
    _cache["k"] = (key, weakref.ref(edge_index), gi)
    return gi
