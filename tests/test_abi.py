"""CPU tests of the C-ABI boundary: libpertgnn.so loads here (no GPU needed) and exports every symbol that
include/pertgnn.h declares; the ctypes table mirrors the header; argument validation happens before any CUDA call."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pertgnn.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pert_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from pert_gnn_kdd23_b200 import _lib

    names = _declared()
    assert len(names) >= 20
    h = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(h, n), f"{n} declared in include/pertgnn.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} missing from the ctypes table"
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_host_side_validation():
    from pert_gnn_kdd23_b200 import _lib

    L = _lib.lib()
    assert L.pert_version() >= 2003
    assert L.pert_index_workspace_bytes(10, 20) > 0
    assert L.pert_index_workspace_bytes(-1, 0) == -1
    assert L.pert_tconv_supported_width(64) == 1 and L.pert_tconv_supported_width(65) == 0
    # bad arguments are rejected with PERT_ERR_BADARG before touching the device
    assert L.pert_build_index(None, None, 0, -1, 0, 0, 0, None, None, None, None, None, None, None, None, None, 0,
                              None, None) == -1
    assert L.pert_segment_reduce_fwd(None, None, None, None, 5, 0, 1, None) == -1
    assert L.pert_adam_step(None, None, None, None, 10, 1e-3, .9, .999, 1e-8, 0., 1, 1., None) == -1
    assert L.pert_allreduce_adam(None, None, None, None, 10, 1e-3, .9, .999, 1e-8, 0., 1, 1., None, 0, 2, None,
                                 None, None) == -1
    assert L.pert_node_depth(None, 3, None, None, None) == -1 and L.pert_level_order(None, 3, None, None, None) == -1
    assert L.pert_eval_metrics(None, None, 0.5, 4, None, None) == -1
    assert L.pert_pert_graph_count(None, 1, None, None, 8, None, None, None) == -1
    assert L.pert_pert_graph_build(*([None, 1, 4] + [None] * 8 + [8, 0] + [None] * 6)) == -1
    assert L.pert_peer_exchange_bytes(1000) >= 2 * 1000 * 4 and L.pert_peer_exchange_bytes(-1) == 0
    assert L.pert_peer_open(None, None) == -1 and L.pert_peer_close(None) == 0 and L.pert_peer_free(None) == 0


def test_header_is_plain_c():
    """The boundary is a C header: it must compile as C99 without any CUDA / C++ / torch include."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    r = subprocess.run([gcc, "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-Werror", HEADER],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_product_path_refuses_cpu_tensors():
    import torch

    from pert_gnn_kdd23_b200 import _lib, ops

    with pytest.raises(_lib.PertGnnError):
        ops.linear(torch.zeros(4, 4), torch.zeros(4, 4))
