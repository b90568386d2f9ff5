"""Drop-in for the dense graph construction of ResGCN/gcn_lib: the kNN graph only."""
