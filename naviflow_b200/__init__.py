"""naviflow_b200 -- B200-native (sm_100a, fp64 CUDA) SIMPLE hot path behind NaviFlow's plugin API.

Host classes mirror ``naviflow_oo``'s Mesh / Fluid / BoundaryConditionManager / Algorithm / Solver objects;
all arithmetic happens in ``libnaviflow_b200.so`` (C-ABI in include/naviflow_b200.h).  There is no CPU
fallback: using a solver without the built library or without a CUDA device raises.
"""
from .host import (BoundaryConditionManager, FluidProperties, SimulationResult, StructuredMesh, ghia_errors,
                   ghia_table)
from .momentum import (GpuJacobiMomentumSolver, GpuMatrixFreeMomentumSolver, GpuQUICKDiscretization,
                       GpuSecondOrderUpwindDiscretization)
from .pressure import (GpuBiCGSTABSolver, GpuCGSolver, GpuGaussSeidelSolver, GpuGeoMultigridPrecondCGSolver,
                       GpuJacobiSolver, GpuMultiGridSolver)
from .profiler import Profiler, load_profile
from .simple import GpuPisoSolver, GpuSimplecSolver, GpuSimpleSolver, GpuSimplerSolver
from .velocity import GpuVelocityUpdater

__all__ = ["StructuredMesh", "FluidProperties", "BoundaryConditionManager", "SimulationResult", "ghia_errors",
           "ghia_table", "GpuJacobiMomentumSolver", "GpuJacobiSolver", "GpuGaussSeidelSolver",
           "GpuMultiGridSolver", "GpuGeoMultigridPrecondCGSolver", "GpuCGSolver", "GpuBiCGSTABSolver", "GpuVelocityUpdater", "GpuSimpleSolver",
           "GpuPisoSolver", "GpuSimplerSolver", "GpuSimplecSolver", "GpuMatrixFreeMomentumSolver", "GpuQUICKDiscretization",
           "GpuSecondOrderUpwindDiscretization", "Profiler",
           "load_profile"]


def register_with_reference():
    """When ``naviflow_oo`` is importable, registers the Gpu* classes as virtual subclasses of the reference's abstract
    bases (``PressureSolver``, ``MomentumSolver``, ``VelocityUpdater``, ``BaseAlgorithm``: base_pressure_solver.py:4,
    base_momentum_solver.py:8, base_velocity_solver.py, base_algorithm.py:13), so ``isinstance`` checks in user code accept
    them.  Returns the list of (class, base) pairs registered; an empty list when the reference is not installed."""
    try:
        from naviflow_oo.solver.Algorithms.base_algorithm import BaseAlgorithm
        from naviflow_oo.solver.momentum_solver.base_momentum_solver import MomentumSolver
        from naviflow_oo.solver.pressure_solver.base_pressure_solver import PressureSolver
        from naviflow_oo.solver.velocity_solver.base_velocity_solver import VelocityUpdater
    except Exception:
        return []
    pairs = [(c, PressureSolver) for c in (GpuJacobiSolver, GpuGaussSeidelSolver, GpuMultiGridSolver, GpuCGSolver,
                                           GpuBiCGSTABSolver, GpuGeoMultigridPrecondCGSolver)]
    pairs += [(c, MomentumSolver) for c in (GpuJacobiMomentumSolver, GpuMatrixFreeMomentumSolver)]
    pairs += [(GpuVelocityUpdater, VelocityUpdater)]
    pairs += [(c, BaseAlgorithm) for c in (GpuSimpleSolver, GpuPisoSolver, GpuSimplerSolver, GpuSimplecSolver)]
    for cls, base in pairs:
        base.register(cls)
    return pairs
