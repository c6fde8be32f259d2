"""pert_gnn_kdd23_b200 -- B200-native hot path of PERT-GNN."""
__version__ = "0.1.0"
