from pert_gnn_kdd23_b200.data import DataLoader  # noqa: F401
