from pert_gnn_kdd23_b200.data import Batch, Data  # noqa: F401
