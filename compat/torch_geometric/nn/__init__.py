from pert_gnn_kdd23_b200.nn import Linear, TransformerConv, global_add_pool  # noqa: F401
