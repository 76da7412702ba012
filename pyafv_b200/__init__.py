"""pyafv_b200 -- B200-native backend for PyAFV's per-step build / force / step path."""
from .physical_params import PhysicalParams, target_delta
from .finite_voronoi import FiniteVoronoiSimulator
from .parallel_voronoi import ParallelFiniteVoronoiSimulator
from .decomposition import DomainDecomposition, decompose_points
from ._connectutil import tile_pbc, rebuild_connection_matrix, get_cluster_sizes, select_daughter_cluster

__version__ = "0.1.0"

__all__ = ["PhysicalParams", "FiniteVoronoiSimulator", "ParallelFiniteVoronoiSimulator", "target_delta", "tile_pbc",
           "rebuild_connection_matrix", "get_cluster_sizes", "select_daughter_cluster", "DomainDecomposition", "decompose_points"]
