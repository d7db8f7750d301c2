"""naviflow_b200 -- B200-native (sm_100a, fp64 CUDA) SIMPLE hot path behind NaviFlow's plugin API.

Host classes mirror ``naviflow_oo``'s Mesh / Fluid / BoundaryConditionManager / Algorithm / Solver objects;
all arithmetic happens in ``libnaviflow_b200.so`` (C-ABI in include/naviflow_b200.h).  There is no CPU
fallback: using a solver without the built library or without a CUDA device raises.
"""
from .host import (BoundaryConditionManager, FluidProperties, SimulationResult, StructuredMesh, ghia_errors,
                   ghia_table)
from .momentum import GpuJacobiMomentumSolver, GpuMatrixFreeMomentumSolver
from .pressure import (GpuBiCGSTABSolver, GpuCGSolver, GpuGaussSeidelSolver, GpuJacobiSolver,
                       GpuMultiGridSolver)
from .simple import GpuPisoSolver, GpuSimpleSolver, GpuSimplerSolver
from .velocity import GpuVelocityUpdater

__all__ = ["StructuredMesh", "FluidProperties", "BoundaryConditionManager", "SimulationResult", "ghia_errors",
           "ghia_table", "GpuJacobiMomentumSolver", "GpuJacobiSolver", "GpuGaussSeidelSolver",
           "GpuMultiGridSolver", "GpuCGSolver", "GpuBiCGSTABSolver", "GpuVelocityUpdater", "GpuSimpleSolver",
           "GpuPisoSolver", "GpuSimplerSolver", "GpuMatrixFreeMomentumSolver"]
