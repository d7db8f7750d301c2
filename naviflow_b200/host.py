"""Host-side model objects with the reference's interface (naviflow_oo L1/L5 objects).

The GPU plugin classes only use the *interface* of these objects (duck typing), so the reference's own
``StructuredMesh`` / ``FluidProperties`` / ``BoundaryConditionManager`` instances can be passed instead;
these stand-alone versions exist because the reference package is not importable on the GPU box
(matplotlib / pyamg imports).

Reference (paths relative to /root/reference/naviflow_oo):
  StructuredMesh             preprocessing/mesh/structured.py:6-43
  FluidProperties            constructor/properties/fluid.py:4-54
  BoundaryConditionManager   constructor/boundary_conditions.py:84-288
  Ghia tables / errors       postprocessing/validation/cavity_flow.py:29-124, 178-301
"""
from __future__ import annotations

import json
import math
import os

import numpy as np

_LOCATIONS = ("top", "bottom", "left", "right")
_TYPES = ("wall", "velocity", "pressure", "inflow", "outflow", "symmetry")


class StructuredMesh:
    """Uniform mesh; note the reference's spacing dx = L/(nx-1) although arrays are cell based
    (structured.py:27-28)."""

    def __init__(self, nx, ny, length=1.0, height=1.0):
        self.nx, self.ny = int(nx), int(ny)
        self.length, self.height = float(length), float(height)
        self.dx = self.length / (self.nx - 1)
        self.dy = self.height / (self.ny - 1)
        self.x = np.linspace(self.dx / 2, self.length - self.dx / 2, self.nx)
        self.y = np.linspace(self.dy / 2, self.height - self.dy / 2, self.ny)

    def get_dimensions(self):
        return self.nx, self.ny

    def get_cell_sizes(self):
        return self.dx, self.dy


class FluidProperties:
    """rho, mu (= rho U L / Re when only the Reynolds number is given; fluid.py:35-46)."""

    def __init__(self, density=1.0, viscosity=None, reynolds_number=None, characteristic_velocity=1.0,
                 characteristic_length=1.0):
        self.density = density
        self.characteristic_velocity = characteristic_velocity
        self.characteristic_length = characteristic_length
        self.reynolds_number = reynolds_number
        if viscosity is None:
            if reynolds_number is None:
                raise ValueError("Either viscosity or Reynolds number must be provided")
            self.viscosity = density * characteristic_velocity * characteristic_length / reynolds_number
        else:
            self.viscosity = viscosity
            if reynolds_number is None:
                self.reynolds_number = density * characteristic_velocity * characteristic_length / viscosity

    def get_density(self):
        return self.density

    def get_viscosity(self):
        return self.viscosity

    def get_reynolds_number(self):
        return self.reynolds_number


def _edge_ops(conditions):
    """The reference's BC routine as an ordered list of (edge, u_value, v_value) writes: the four default
    walls in its fixed order, then every registered velocity / wall condition in insertion order
    (boundary_conditions.py:180-258)."""
    ops = [("left", 0.0, None), ("right", 0.0, None), ("bottom", 0.0, None), ("top", 0.0, None),
           ("left", None, 0.0), ("right", None, 0.0), ("bottom", None, 0.0), ("top", None, 0.0)]
    for loc, conds in conditions.items():
        for typ, vals in conds.items():
            if typ == "velocity":
                uu, vv = (vals or {}).get("u", 0.0), (vals or {}).get("v", 0.0)
            elif typ == "wall":
                uu, vv = 0.0, 0.0
            else:
                continue
            if loc in _LOCATIONS:
                ops.append((loc, uu, vv))
    return ops


def _apply_ops(ops, u, v, nx, ny):
    """Executes the edge writes with the reference's shape-dependent index rules."""
    for loc, uu, vv in ops:
        if uu is not None:
            if loc == "left":
                u[0, :] = uu
            elif loc == "right":
                if u.shape[0] == nx + 1:
                    u[nx, :] = uu
                elif u.shape[0] == nx and nx > 0:
                    u[nx - 1, :] = uu
            elif loc == "bottom":
                u[:, 0] = uu
            elif loc == "top":
                if u.shape[1] > ny - 1 and ny > 0:
                    u[:, ny - 1] = uu
        if vv is not None:
            if loc == "left":
                v[0, :] = vv
            elif loc == "right":
                if v.shape[0] > nx - 1 and nx > 0:
                    v[nx - 1, :] = vv
            elif loc == "bottom":
                v[:, 0] = vv
            elif loc == "top":
                if v.shape[1] == ny + 1:
                    v[:, ny] = vv
                elif v.shape[1] == ny and ny > 0:
                    v[:, ny - 1] = vv
    return u, v


def conditions_of(bc):
    """Ordered {location: {type: values}} of a BoundaryConditionManager (ours or the reference's) or a dict."""
    if bc is None:
        return {}
    if hasattr(bc, "conditions"):
        return bc.conditions
    if hasattr(bc, "to_dict"):
        return bc.to_dict()
    return dict(bc)


class BoundaryConditionManager:
    """Registry of boundary conditions; insertion order is significant (it decides the lid corners).

    ``items()`` / ``len()`` give the ``{location: {type: values}}`` mapping view: the reference's solvers treat every
    boundary-condition object that is not an instance of their own manager class as such a mapping and rebuild a manager
    from it (jacobi_matrix_solver.py:162-170, standard.py), which is what makes this class usable inside the reference's
    own loop (tests/test_reference_api.py)."""

    def __init__(self):
        self.conditions = {}

    def items(self):
        return self.conditions.items()

    def __len__(self):
        return len(self.conditions)

    def set_condition(self, location, bc_type, values=None):
        loc = getattr(location, "name", location).lower()
        typ = getattr(bc_type, "name", bc_type).lower()
        if loc not in _LOCATIONS:
            raise ValueError(f"Unknown boundary location: {loc.upper()}")
        if typ not in _TYPES:
            raise ValueError(f"Unknown boundary type: {typ.upper()}")
        self.conditions.setdefault(loc, {})[typ] = values or {}

    def get_condition(self, location, bc_type=None):
        loc = getattr(location, "name", location).lower()
        if loc not in self.conditions:
            return None
        if bc_type is None:
            return self.conditions[loc]
        return self.conditions[loc].get(getattr(bc_type, "name", bc_type).lower())

    def apply_velocity_boundary_conditions(self, u, v, nx, ny):
        return _apply_ops(_edge_ops(self.conditions), u, v, nx, ny)

    def to_dict(self):
        return self.conditions

    def get_boundary_types(self):
        out = {}
        for loc, conds in self.conditions.items():
            if conds:
                out[loc] = next(iter(conds.keys()))
        for loc in _LOCATIONS:
            out.setdefault(loc, "wall")
        return out


def boundary_program(bc, nx, ny, nx_arg=None):
    """Evaluates the BC routine symbolically: final constant per edge line / corner of u (nx+1, ny) and
    v (nx, ny+1) when the routine is called with ``(nx_arg, ny)`` (some reference callers pass nx+1, which
    disables the v[nx-1,:] writes).  Returns the fields of nf_bc_program as plain lists; NaN = untouched."""
    nx_arg = nx if nx_arg is None else nx_arg
    conds = conditions_of(bc)
    # run the routine on a NaN-filled 5x5 stand-in whose index arithmetic mirrors the real shapes
    m = 5
    u = np.full((m + 1, m), np.nan)
    v = np.full((m, m + 1), np.nan)
    _apply_ops(_edge_ops(conds), u, v, m + (nx_arg - nx), m)
    return {
        "u_edge": [u[0, 2], u[m, 2], u[2, 0], u[2, m - 1]],
        "u_corner": [u[0, 0], u[0, m - 1], u[m, 0], u[m, m - 1]],
        "v_edge": [v[0, 2], v[m - 1, 2], v[2, 0], v[2, m]],
        "v_corner": [v[0, 0], v[0, m], v[m - 1, 0], v[m - 1, m]],
        "v_right_row": (nx - 1) if not math.isnan(v[m - 1, 2]) else -1,
    }


def practice_b_sides(bc):
    """Bit mask of the boundaries with a registered condition (power_law.py:146-199): 1 left, 2 right,
    4 bottom, 8 top."""
    conds = conditions_of(bc)
    mask = 0
    for bit, loc in ((1, "left"), (2, "right"), (4, "bottom"), (8, "top")):
        if conds.get(loc):
            mask |= bit
    return mask


# ---- Ghia et al. validation (cavity_flow.py) ------------------------------------------------------
_GHIA = None


def ghia_table(reynolds):
    """Ghia, Ghia & Shin (1982) centre-line data as tabulated by the reference (cavity_flow.py:29-124)."""
    global _GHIA
    if _GHIA is None:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ghia_tables.json")) as f:
            _GHIA = json.load(f)
    key = str(int(reynolds))
    if key not in _GHIA:
        raise ValueError(f"no Ghia table for Re={reynolds}; available: {sorted(_GHIA, key=int)}")
    return _GHIA[key]


def ghia_errors(u, v, mesh, reynolds):
    """(infinity-norm error, L2 error) of the centre-line profiles (cavity_flow.py:178-301)."""
    from scipy.interpolate import interp1d
    nx, ny = mesh.get_dimensions()
    dx, dy = mesh.get_cell_sizes()
    t = ghia_table(reynolds)
    x = np.linspace(dx / 2, 1 - dx / 2, nx)
    y = np.linspace(dy / 2, 1 - dy / 2, ny)
    ui = interp1d(y, u[nx // 2, :], kind="cubic", bounds_error=False, fill_value="extrapolate")(np.asarray(t["y"]))
    vi = interp1d(x, v[:, ny // 2], kind="cubic", bounds_error=False, fill_value="extrapolate")(np.asarray(t["x"]))
    ue, ve = ui - np.asarray(t["u"]), vi - np.asarray(t["v"])
    inf = max(np.max(np.abs(ue)), np.max(np.abs(ve)))
    l2 = math.sqrt((np.sum(ue ** 2) + np.sum(ve ** 2)) / (len(ue) + len(ve)))
    return float(inf), float(l2)


class SimulationResult:
    """Minimal result container with the reference's accessors (postprocessing/simulation_result.py:11-370)."""

    def __init__(self, u, v, p, mesh, iterations=0, residuals=None, reynolds=None, wall_time=None):
        self.u, self.v, self.p, self.mesh = u, v, p, mesh
        self.iterations = iterations
        self.residuals = residuals or []
        self.reynolds = reynolds
        self.wall_time = wall_time
        self.histories = {}
        self.infinity_norm_error = None

    def add_history(self, name, values):
        self.histories[name] = list(values)

    def get_history(self, name):
        return self.histories.get(name)

    def get_max_divergence(self):
        dx, dy = self.mesh.get_cell_sizes()
        div = (self.u[1:, :] - self.u[:-1, :]) / dx + (self.v[:, 1:] - self.v[:, :-1]) / dy
        return float(np.max(np.abs(div[1:-1, 1:-1])))

    def save_solution(self, filename):
        """np.savez(u, v, p, x, y, reynolds) like simulation_result.py:296-314."""
        x = getattr(self.mesh, "x", None)
        y = getattr(self.mesh, "y", None)
        np.savez(filename, u=self.u, v=self.v, p=self.p, x=x, y=y, reynolds=self.reynolds)
