"""Import the real NaviFlow reference (``/root/reference``) for oracle pinning.

TEST INFRASTRUCTURE.  Only usable in the build container: the GPU box has no
``/root/reference``; ``available()`` is False there and every caller skips.

The reference imports matplotlib / scienceplots / pyamg at module import time
(naviflow_oo/solver/Algorithms/simple.py:7, pressure_solver/multigrid.py:3-4,
pressure_solver/__init__.py:6-7, postprocessing/visualization.py:11).  Those
packages are not installed, so inert stub modules are injected into
``sys.modules`` *only when the real package is missing*.  numpy/scipy are never
stubbed.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("NAVIFLOW_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "naviflow_oo"))


class _Swallow:
    """Object that absorbs any attribute access / call (plotting stubs)."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return self

    def __call__(self, *a, **k):
        return self

    def __iter__(self):
        return iter(())

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __getitem__(self, k):
        return self

    def __setitem__(self, k, v):
        pass


def _stub_module(name, **attrs):
    mod = types.ModuleType(name)
    sw = _Swallow()
    mod.__dict__.update(attrs)
    mod.__getattr__ = lambda attr, _sw=sw: _sw  # PEP 562 module-level getattr
    mod.__path__ = []  # behave like a package so submodule imports resolve
    sys.modules[name] = mod
    return mod


def _missing(name):
    try:
        importlib.import_module(name)
        return False
    except Exception:
        return True


def install_stubs():
    if _missing("matplotlib"):
        _stub_module("matplotlib", use=lambda *a, **k: None, rcParams={})
        for sub in ("pyplot", "animation", "cm", "colors", "backends",
                    "backends.backend_pdf", "ticker", "gridspec", "patches",
                    "lines", "figure", "axes"):
            _stub_module("matplotlib." + sub)
        sys.modules["matplotlib.backends.backend_pdf"].PdfPages = _Swallow()
    if _missing("scienceplots"):
        _stub_module("scienceplots")
    if _missing("pyamg"):
        def _no_pyamg(*a, **k):
            raise ImportError("pyamg is not installed (stub)")
        _stub_module("pyamg", smoothed_aggregation_solver=_no_pyamg,
                     ruge_stuben_solver=_no_pyamg)
    if _missing("mpl_toolkits"):
        _stub_module("mpl_toolkits")
        _stub_module("mpl_toolkits.axes_grid1")


_loaded = None


def load():
    """Return the imported ``naviflow_oo`` package of the real reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _loaded = importlib.import_module("naviflow_oo")
    return _loaded


def ref():
    """Namespace with the reference symbols the oracle is pinned against."""
    load()
    ns = types.SimpleNamespace()
    from naviflow_oo.preprocessing.mesh.structured import StructuredMesh
    from naviflow_oo.constructor.properties.fluid import FluidProperties
    from naviflow_oo.constructor.boundary_conditions import BoundaryConditionManager
    from naviflow_oo.solver.momentum_solver.discretization.power_law import PowerLawDiscretization
    from naviflow_oo.solver.momentum_solver.discretization.quick import QUICKDiscretization
    from naviflow_oo.solver.momentum_solver.discretization.second_order_upwind import SecondOrderUpwindDiscretization
    from naviflow_oo.solver.momentum_solver.jacobi_matrix_solver import JacobiMatrixMomentumSolver
    from naviflow_oo.solver.momentum_solver.matrix_free_momentum import MatrixFreeMomentumSolver
    from naviflow_oo.solver.pressure_solver.helpers.rhs_construction import get_rhs
    from naviflow_oo.solver.pressure_solver.helpers.matrix_free import compute_Ap_product
    from naviflow_oo.solver.pressure_solver.helpers.coeff_matrix import get_coeff_mat
    from naviflow_oo.solver.pressure_solver.helpers import multigrid_helpers
    from naviflow_oo.solver.pressure_solver.jacobi import JacobiSolver
    from naviflow_oo.solver.pressure_solver.gauss_seidel import GaussSeidelSolver
    from naviflow_oo.solver.pressure_solver.multigrid import MultiGridSolver
    from naviflow_oo.solver.pressure_solver.direct import DirectPressureSolver
    from naviflow_oo.solver.pressure_solver.matrix_free_BiCGSTAB import MatrixFreeBiCGSTABSolver
    from naviflow_oo.solver.velocity_solver.standard import StandardVelocityUpdater
    from naviflow_oo.solver.Algorithms.simple import SimpleSolver
    from naviflow_oo.solver.Algorithms.piso import PisoSolver
    from naviflow_oo.solver.Algorithms.simpler import SimplerSolver
    from naviflow_oo.postprocessing.validation import cavity_flow
    ns.__dict__.update(locals())
    del ns.__dict__["ns"]

    class JacobiMatrixMomentumAdapter(JacobiMatrixMomentumSolver):
        """4-tuple -> (field, d, info) so it plugs into SimpleSolver."""

        def solve_u_momentum(self, mesh, fluid, u, v, p, relaxation_factor=0.7,
                             boundary_conditions=None, return_dict=True):
            us, du, norm, field = super().solve_u_momentum(
                mesh, fluid, u, v, p, relaxation_factor, boundary_conditions)
            return us, du, {"rel_norm": norm, "field": field}

        def solve_v_momentum(self, mesh, fluid, u, v, p, relaxation_factor=0.7,
                             boundary_conditions=None, return_dict=True):
            vs, dv, norm, field = super().solve_v_momentum(
                mesh, fluid, u, v, p, relaxation_factor, boundary_conditions)
            return vs, dv, {"rel_norm": norm, "field": field}

    ns.JacobiMatrixMomentumAdapter = JacobiMatrixMomentumAdapter
    return ns
