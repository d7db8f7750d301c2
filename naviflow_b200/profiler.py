"""Run record of an outer-loop solve -- twin of ``naviflow_oo.utils.profiler.Profiler``.

Reference: utils/profiler.py:17-463 of /root/reference/naviflow_oo.  Same methods (``start``, ``end``,
``start_section`` / ``end_section``, ``set_iterations``, ``set_convergence_info``, ``add_residual_data``,
``set_pressure_solver_info``, ``save``) and the same on-disk layout (:317-443): groups ``simulation`` (with the
sub-group ``mesh_size``), ``performance``, ``convergence``, ``system``, ``algorithm``, ``pressure_solver`` (with
``smoother`` / ``multigrid``), ``momentum_solver`` whose scalars are stored as attributes, and the group
``residual_history`` with one dataset per column (``iteration, wall_time, cpu_time, total_residual, momentum_residual,
pressure_residual, infinity_norm_error``).

The reference writes HDF5 through h5py.  h5py is an optional dependency here: when it is importable the file is that
HDF5 file, otherwise the same tree goes into an ``.npz`` next to it -- attribute ``a`` of group ``g/h`` becomes the entry
``g/h@a``, dataset ``d`` of group ``g`` the entry ``g/d`` -- so the reference's notebooks need a three-line loader instead
of ``h5py.File`` (``load_profile`` below returns the same nested dict for both formats).

The device loop does not return to the host between outer iterations, so per-iteration wall times are interpolated
linearly over each chunk of iterations the host waited for (``add_residual_block``).
"""
from __future__ import annotations

import os
import platform
import time
from datetime import datetime

import numpy as np


def _have_h5py():
    try:
        import h5py  # noqa: F401
        return True
    except Exception:
        return False


class Profiler:
    def __init__(self, algorithm_name, mesh, fluid, algorithm=None):
        self.algorithm_name = algorithm_name
        self.mesh = mesh
        self.fluid = fluid
        self.algorithm = algorithm
        self._start_time = None
        self._start_cpu_time = None
        self._section_start_time = None
        self._section_start_cpu_time = None
        self.initialize()

    def initialize(self):
        """profiler.py:46-89"""
        try:
            import psutil
            mem_total = psutil.virtual_memory().total / (1024 ** 3)
        except Exception:
            mem_total = float("nan")
        self.profiling_data = {
            "total_time": 0.0, "cpu_time": 0.0, "iterations": 0, "memory_usage": [],
            "timestamp": datetime.now().strftime("%Y-%m-%d %H:%M:%S"),
            "system_info": {"platform": platform.platform(), "processor": self._processor(),
                            "python_version": platform.python_version(), "memory_total": mem_total},
            "convergence_info": {"tolerance": None, "final_residual": None, "converged": False, "residual_history": []},
            "detailed_residuals": {"iterations": [], "wall_times": [], "cpu_times": [], "total_residuals": [],
                                   "momentum_residuals": [], "pressure_residuals": [], "infinity_norm_errors": []},
            "pressure_solver_info": {"name": None, "total_inner_iterations": 0, "avg_inner_iterations_per_outer": 0.0,
                                     "max_inner_iterations": 0, "min_inner_iterations": float("inf"),
                                     "inner_iterations_history": [], "convergence_rate": None, "solver_specific": {}},
        }

    @staticmethod
    def _processor():
        """profiler.py:91-131 plus the accelerator the arithmetic runs on."""
        name = platform.processor()
        if platform.system() == "Linux":
            try:
                with open("/proc/cpuinfo") as f:
                    for line in f:
                        if line.startswith("model name"):
                            name = line.split(":")[1].strip()
                            break
            except Exception:
                pass
        try:
            import torch
            if torch.cuda.is_available():
                name = f"{name} + {torch.cuda.get_device_name(torch.cuda.current_device())}"
        except Exception:
            pass
        return name

    # ---- timing (profiler.py:133-178) ----
    def start(self):
        self._start_time = time.time()
        self._start_cpu_time = time.process_time()

    def end(self):
        if self._start_time is not None:
            self.profiling_data["total_time"] = time.time() - self._start_time
            self.profiling_data["cpu_time"] = time.process_time() - self._start_cpu_time

    def start_section(self):
        self._section_start_time = time.time()
        self._section_start_cpu_time = time.process_time()
        return self._section_start_time

    def end_section(self, section_name):
        if self._section_start_time is None:
            return
        wall = time.time() - self._section_start_time
        cpu = time.process_time() - self._section_start_cpu_time
        self.profiling_data[section_name] = self.profiling_data.get(section_name, 0.0) + wall
        self.profiling_data[section_name + "_cpu"] = self.profiling_data.get(section_name + "_cpu", 0.0) + cpu
        self._section_start_time = None
        self._section_start_cpu_time = None

    # ---- records (profiler.py:180-288) ----
    def set_iterations(self, iterations):
        self.profiling_data["iterations"] = iterations

    def set_convergence_info(self, tolerance, final_residual, residual_history, converged=None):
        if converged is None:
            converged = final_residual <= tolerance
        ci = self.profiling_data["convergence_info"]
        ci["tolerance"], ci["final_residual"], ci["residual_history"], ci["converged"] = \
            tolerance, final_residual, residual_history, converged

    def add_residual_data(self, iteration, total_residual, momentum_residual, pressure_residual, infinity_norm_error=None,
                          wall_time=None, cpu_time=None):
        if wall_time is None:
            wall_time = time.time() - self._start_time if self._start_time is not None else 0.0
        if cpu_time is None:
            cpu_time = time.process_time() - self._start_cpu_time if self._start_cpu_time is not None else 0.0
        d = self.profiling_data["detailed_residuals"]
        d["iterations"].append(iteration)
        d["wall_times"].append(wall_time)
        d["cpu_times"].append(cpu_time)
        d["total_residuals"].append(total_residual)
        d["momentum_residuals"].append(momentum_residual)
        d["pressure_residuals"].append(pressure_residual)
        d["infinity_norm_errors"].append(infinity_norm_error)

    def add_residual_block(self, first_iteration, totals, momentum, pressure, t_begin, t_end):
        """Records of a chunk of outer iterations the device ran without returning to the host: wall / cpu times are spread
        evenly over [t_begin, t_end] (both measured from start())."""
        n = len(totals)
        cpu_now = time.process_time() - self._start_cpu_time if self._start_cpu_time is not None else 0.0
        for k in range(n):
            frac = (k + 1) / n
            self.add_residual_data(first_iteration + k, totals[k], momentum[k], pressure[k], None,
                                   wall_time=t_begin + frac * (t_end - t_begin), cpu_time=cpu_now)

    def set_pressure_solver_info(self, solver_name, inner_iterations=None, convergence_rate=None, solver_specific=None):
        info = self.profiling_data["pressure_solver_info"]
        info["name"] = solver_name
        if inner_iterations is not None:
            inner_iterations = list(inner_iterations)
            info["inner_iterations_history"] = inner_iterations
            info["total_inner_iterations"] = sum(inner_iterations)
            if inner_iterations:
                info["avg_inner_iterations_per_outer"] = sum(inner_iterations) / len(inner_iterations)
                info["max_inner_iterations"] = max(inner_iterations)
                info["min_inner_iterations"] = min(inner_iterations)
        if convergence_rate is not None:
            info["convergence_rate"] = convergence_rate
        if solver_specific is not None:
            info["solver_specific"] = solver_specific

    # ---- the file (profiler.py:290-463) ----
    def metadata(self):
        pd = self.profiling_data
        nx, ny = self.mesh.get_dimensions()
        meta = {
            "simulation": {"algorithm": self.algorithm_name, "timestamp": pd["timestamp"],
                           "mesh_size": {"x": nx, "y": ny}, "reynolds_number": self.fluid.get_reynolds_number()},
            "performance": {"total_time": pd["total_time"], "cpu_time": pd["cpu_time"], "iterations": pd["iterations"],
                            "avg_time_per_iteration": pd["total_time"] / pd["iterations"] if pd["iterations"] > 0 else 0},
            "convergence": {"tolerance": pd["convergence_info"]["tolerance"],
                            "final_residual": pd["convergence_info"]["final_residual"],
                            "converged": pd["convergence_info"]["converged"]},
            "system": {"platform": pd["system_info"]["platform"], "processor": pd["system_info"]["processor"],
                       "python_version": pd["system_info"]["python_version"]},
        }
        alg = self.algorithm
        if alg is not None:
            params = {k: getattr(alg, k) for k in ("alpha_p", "alpha_u") if hasattr(alg, k)}
            if params:
                meta["algorithm"] = params
            ps = getattr(alg, "pressure_solver", None)
            if ps is not None:
                d = {"type": ps.__class__.__name__}
                for k in ("tolerance", "max_iterations", "matrix_free"):
                    if hasattr(ps, k):
                        d[k] = getattr(ps, k)
                mgp = {k: getattr(ps, k) for k in ("cycle_type", "pre_smoothing", "post_smoothing") if hasattr(ps, k)}
                sm = {}
                if hasattr(ps, "smoother"):
                    sm["type"] = ps.smoother.__class__.__name__
                if hasattr(ps, "smoother_iterations"):
                    sm["iterations"] = ps.smoother_iterations
                if hasattr(ps, "smoother_omega"):
                    sm["omega"] = ps.smoother_omega
                if sm:
                    d["smoother"] = sm
                if mgp:
                    d["multigrid"] = mgp
                meta["pressure_solver"] = d
            ms = getattr(alg, "momentum_solver", None)
            if ms is not None:
                meta["momentum_solver"] = {"type": ms.__class__.__name__}
        return meta

    def residual_columns(self):
        d = self.profiling_data["detailed_residuals"]
        if not d["iterations"]:
            return {}
        cols = {"iteration": np.asarray(d["iterations"]), "wall_time": np.asarray(d["wall_times"], dtype=float),
                "cpu_time": np.asarray(d["cpu_times"], dtype=float),
                "total_residual": np.asarray(d["total_residuals"], dtype=float),
                "momentum_residual": np.asarray(d["momentum_residuals"], dtype=float),
                "pressure_residual": np.asarray(d["pressure_residuals"], dtype=float)}
        if d["infinity_norm_errors"]:
            cols["infinity_norm_error"] = np.asarray([np.nan if e is None else e for e in d["infinity_norm_errors"]],
                                                     dtype=float)
        return cols

    def save(self, filename=None, profile_dir="results/profiles"):
        if filename is None:
            nx, ny = self.mesh.get_dimensions()
            filename = os.path.join(profile_dir, f"{self.algorithm_name}_Re{int(self.fluid.get_reynolds_number())}"
                                                 f"_mesh{nx}x{ny}_profile.h5")
        os.makedirs(os.path.dirname(os.path.abspath(filename)), exist_ok=True)
        meta, cols = self.metadata(), self.residual_columns()
        if _have_h5py():
            import h5py
            with h5py.File(filename, "w") as f:
                for gname, gdata in meta.items():
                    _store_h5(f.create_group(gname), gdata)
                if cols:
                    g = f.create_group("residual_history")
                    for k, v in cols.items():
                        g.create_dataset(k, data=v)
            return os.path.abspath(filename)
        flat = {}
        for gname, gdata in meta.items():
            _flatten(gname, gdata, flat)
        for k, v in cols.items():
            flat["residual_history/" + k] = v
        path = os.path.splitext(filename)[0] + ".npz"
        np.savez(path, **flat)
        return os.path.abspath(path)


def _attr_value(v):
    return v if isinstance(v, (int, float, str, bool, np.number)) else str(v)


def _store_h5(group, data):
    for k, v in data.items():
        if isinstance(v, dict):
            _store_h5(group.create_group(k), v)
        else:
            group.attrs[k] = _attr_value(v)


def _flatten(prefix, data, out):
    for k, v in data.items():
        if isinstance(v, dict):
            _flatten(prefix + "/" + k, v, out)
        else:
            out[prefix + "@" + k] = np.asarray(_attr_value(v))


def load_profile(path):
    """Nested dict {group: {attr: value, sub-group: {...}}, 'residual_history': {column: array}} of a file written by
    Profiler.save (either format) or by the reference's own Profiler."""
    out = {}
    if path.endswith(".npz"):
        z = np.load(path, allow_pickle=False)
        for key in z.files:
            if "@" in key:
                g, a = key.split("@")
                node = out
                for part in g.split("/"):
                    node = node.setdefault(part, {})
                v = z[key]
                node[a] = v.item() if v.shape == () else v
            else:
                g, d = key.rsplit("/", 1)
                node = out
                for part in g.split("/"):
                    node = node.setdefault(part, {})
                node[d] = z[key]
        return out
    import h5py

    def walk(h, node):
        for k, v in h.attrs.items():
            node[k] = v
        for k, v in h.items():
            if isinstance(v, h5py.Group):
                walk(v, node.setdefault(k, {}))
            else:
                node[k] = v[()]
    with h5py.File(path, "r") as f:
        walk(f, out)
    return out
