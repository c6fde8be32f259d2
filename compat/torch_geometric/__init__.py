"""Minimal stand-in for the parts of torch_geometric 2.4.0 that the reference imports (pert_gnn.py:2-3,
model.py:2-7), backed by pert_gnn_kdd23_b200.  Not a general PyG replacement."""
__version__ = "2.4.0+pertgnn.b200"
