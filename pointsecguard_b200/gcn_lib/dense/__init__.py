from .torch_edge import DenseDilated, DenseDilatedKnnGraph, dense_knn_matrix, pairwise_distance  # noqa: F401
