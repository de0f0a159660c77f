"""Python face of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

ctypes bindings of oracle/libecmc_oracle.so (built by oracle/Makefile from oracle/ecmc_oracle.c) plus a
numpy/pure-Python restatement of the init-time table construction of the reference:

* InnerPointEstimator.derivative_bound      jellyfysh/estimator/inner_point_estimator.py:108-163
* CellVetoEventHandler.initialize           jellyfysh/event_handler/abstracts/cell_veto_event_handler.py:93-159
* Walker.__init__ / _build_table            jellyfysh/event_handler/walker.py:55-103

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module. The product package jellyfysh_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from jellyfysh_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when the reference checkout exists) with oracle/Makefile."""
    so = os.path.join(_HERE, "libecmc_oracle.so")
    sources = [os.path.join(_HERE, "ecmc_oracle.c"), os.path.join(_HERE, "..", "include", "ecmc.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(src) for src in sources):
        subprocess.run(["make", "-C", _HERE, "-s", "all"], check=True)


def lib() -> C.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    build()
    L = C.CDLL(os.path.join(_HERE, "libecmc_oracle.so"))
    d, i, u32, u64, sz, p = C.c_double, C.c_int, C.c_uint32, C.c_uint64, C.c_size_t, C.c_void_p
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int)
    pot = C.POINTER(abi.EcmcPotential)
    L.orc_time_add.argtypes = [d, d, d, dp, dp]
    L.orc_time_sub.argtypes = [d, d, d, d]
    L.orc_time_sub.restype = d
    L.orc_time_from_float.argtypes = [d, dp, dp]
    L.orc_time_lt.argtypes = [d, d, d, d]
    L.orc_time_lt.restype = i
    L.orc_correct_position_entry.argtypes = [d, d]
    L.orc_correct_position_entry.restype = d
    L.orc_correct_separation_entry.argtypes = [d, d]
    L.orc_correct_separation_entry.restype = d
    L.orc_potential_derivative_batch.argtypes = [pot, i, d, dp, sz, p, p, p]
    L.orc_potential_displacement_batch.argtypes = [pot, i, d, dp, sz, p, p, p, p]
    L.orc_random_doubles.argtypes = [u32, u32, u64, u32, u32, sz, p]
    L.orc_random_words.argtypes = [u32, u32, u64, u32, u32, sz, p]
    L.orc_cells_geometry.argtypes = [i, ip, d, p, p]
    L.orc_position_to_cell.argtypes = [i, ip, d, dp]
    L.orc_cells_translate.argtypes = [i, ip, d, i, i]
    L.orc_cells_relative.argtypes = [i, ip, d, i, i]
    L.orc_nearby_cells.argtypes = [i, ip, i, d, i, ip]
    L.orc_lifting_choose.argtypes = [i, i, p, i, p]
    L.orc_bending_derivative.argtypes = [d, d, i, d, p, p, i, p]
    L.orc_chain_create.argtypes = [C.POINTER(abi.EcmcProgram)]
    L.orc_chain_create.restype = p
    L.orc_chain_destroy.argtypes = [p]
    L.orc_chain_set_positions.argtypes = [p, p, p]
    L.orc_chain_get_positions.argtypes = [p, p]
    L.orc_chain_set_roots.argtypes = [p, p]
    L.orc_chain_get_roots.argtypes = [p, p]
    L.orc_chain_get_state.argtypes = [p, C.POINTER(abi.EcmcChainState)]
    L.orc_chain_set_state.argtypes = [p, C.POINTER(abi.EcmcChainState)]
    L.orc_chain_get_cells.argtypes = [p, p, p, C.POINTER(C.c_int32)]
    L.orc_chain_set_cells.argtypes = [p, p, p, C.c_int32]
    L.orc_chain_get_stats.argtypes = [p, C.POINTER(abi.EcmcStats)]
    L.orc_chain_start.argtypes = [p, u32]
    L.orc_chain_run.argtypes = [p, d, d, C.c_int64, p, C.c_int64]
    L.orc_chain_run.restype = C.c_int64
    _LIB = L
    return L


# ----------------------------------------------------------------------------------------------------------
# scalar helpers
# ----------------------------------------------------------------------------------------------------------
def _vec(values):
    arr = (C.c_double * 3)(0.0, 0.0, 0.0)
    for k, v in enumerate(values):
        arr[k] = float(v)
    return arr


def time_add(q, r, other):
    oq, orr = C.c_double(), C.c_double()
    lib().orc_time_add(q, r, other, C.byref(oq), C.byref(orr))
    return oq.value, orr.value


def time_sub(q1, r1, q2, r2):
    return lib().orc_time_sub(q1, r1, q2, r2)


def time_lt(q1, r1, q2, r2):
    return bool(lib().orc_time_lt(q1, r1, q2, r2))


def time_from_float(t):
    oq, orr = C.c_double(), C.c_double()
    lib().orc_time_from_float(t, C.byref(oq), C.byref(orr))
    return oq.value, orr.value


def lifting_choose(kind, rates, active_index, uniforms):
    """Index of the next active unit after inserting (rate, index, index == active_index) in order into the lifting
    scheme `kind` (abi.LIFTING_*); uniforms = the random() values behind the scheme's uniform draws, in call order."""
    r = np.ascontiguousarray(rates, dtype=np.float64)
    u = np.ascontiguousarray(list(uniforms) + [0.0, 0.0], dtype=np.float64)
    return int(lib().orc_lifting_choose(kind, len(r), r.ctypes.data, active_index, u.ctypes.data))


def bending_derivative(prefactor, equilibrium_angle, direction, speed, separation_one, separation_two):
    s1 = np.ascontiguousarray(separation_one, dtype=np.float64)
    s2 = np.ascontiguousarray(separation_two, dtype=np.float64)
    out = np.empty(3)
    lib().orc_bending_derivative(prefactor, equilibrium_angle, direction, speed, s1.ctypes.data, s2.ctypes.data, len(s1),
                                 out.ctypes.data)
    return out


def _velocity(direction_or_velocity, dimension, speed=1.0):
    if np.isscalar(direction_or_velocity):
        velocity = [0.0] * dimension
        velocity[int(direction_or_velocity)] = float(speed)
        return _vec(velocity)
    return _vec(direction_or_velocity)


def potential_derivative_batch(pot, dimension, system_length, velocity, separations, charges=None):
    """derivative(velocity, separation, charges) of the reference for n separations. velocity: vector or direction."""
    seps = np.ascontiguousarray(separations, dtype=np.float64).reshape(-1, dimension)
    out = np.empty(len(seps), dtype=np.float64)
    ch = None if charges is None else np.ascontiguousarray(charges, dtype=np.float64)
    lib().orc_potential_derivative_batch(C.byref(pot), dimension, system_length, _velocity(velocity, dimension),
                                         len(seps), seps.ctypes.data, None if ch is None else ch.ctypes.data,
                                         out.ctypes.data)
    return out


def potential_displacement_batch(pot, dimension, system_length, velocity, separations, charges=None,
                                 potential_changes=None):
    """displacement(velocity, separation, charges, potential_change) of the reference (a time) for n inputs."""
    seps = np.ascontiguousarray(separations, dtype=np.float64).reshape(-1, dimension)
    out = np.empty(len(seps), dtype=np.float64)
    ch = None if charges is None else np.ascontiguousarray(charges, dtype=np.float64)
    du = None if potential_changes is None else np.ascontiguousarray(potential_changes, dtype=np.float64)
    lib().orc_potential_displacement_batch(C.byref(pot), dimension, system_length, _velocity(velocity, dimension),
                                           len(seps), seps.ctypes.data, None if ch is None else ch.ctypes.data,
                                           None if du is None else du.ctypes.data, out.ctypes.data)
    return out


def potential_derivative(pot, dimension, system_length, velocity, separation, c1=1.0, c2=1.0):
    return float(potential_derivative_batch(pot, dimension, system_length, velocity, [separation], [[c1, c2]])[0])


def potential_displacement(pot, dimension, system_length, velocity, separation, c1=1.0, c2=1.0, potential_change=0.0):
    return float(potential_displacement_batch(pot, dimension, system_length, velocity, [separation], [[c1, c2]],
                                              [potential_change])[0])


def random_doubles(seed, stream, event, slot, first, n):
    out = np.empty(n, dtype=np.float64)
    lib().orc_random_doubles(seed, stream, event, slot, first, n, out.ctypes.data)
    return out


def random_words(seed, stream, event, slot, first, n):
    out = np.empty(n, dtype=np.uint32)
    lib().orc_random_words(seed, stream, event, slot, first, n, out.ctypes.data)
    return out


def _per_side(per_side):
    return (C.c_int * 3)(*(list(per_side) + [1] * (3 - len(per_side))))


def cells_geometry(dimension, per_side, system_length):
    n = int(np.prod(per_side[:dimension]))
    cmin = np.empty((n, dimension))
    cmax = np.empty((n, dimension))
    lib().orc_cells_geometry(dimension, _per_side(per_side), system_length, cmin.ctypes.data, cmax.ctypes.data)
    return cmin, cmax


def position_to_cell(dimension, per_side, system_length, position):
    return lib().orc_position_to_cell(dimension, _per_side(per_side), system_length, _vec(position))


def cells_translate(dimension, per_side, system_length, cell, relative_cell):
    return lib().orc_cells_translate(dimension, _per_side(per_side), system_length, cell, relative_cell)


def cells_relative(dimension, per_side, system_length, cell, reference_cell):
    return lib().orc_cells_relative(dimension, _per_side(per_side), system_length, cell, reference_cell)


def nearby_cells(dimension, per_side, neighbor_layers, system_length, cell):
    out = (C.c_int * 343)()
    n = lib().orc_nearby_cells(dimension, _per_side(per_side), neighbor_layers, system_length, cell, out)
    return [out[k] for k in range(n)]


# ----------------------------------------------------------------------------------------------------------
# init-time tables (restated from the reference, see module docstring)
# ----------------------------------------------------------------------------------------------------------
def inner_point_derivative_bounds(pot, system_length, per_side, neighbor_layers, prefactor=1.5, points_per_side=10,
                                  empirical_bound=float("inf"), target_charge=None, uses_charges=False):
    """Bounds (upper, -lower) per far relative cell and direction, 3D only like the reference's estimator.

    Returns (bounds[n_cells][3][2] with NaN rows for nearby cells, far_cells list in yield order).
    """
    dimension = 3
    cmin, cmax = cells_geometry(dimension, per_side, system_length)
    n_cells = len(cmin)
    nearby_zero = set(nearby_cells(dimension, per_side, neighbor_layers, system_length, 0))
    bounds = np.full((n_cells, dimension, 2), np.nan)
    far_cells = [cell for cell in range(n_cells) if cell not in nearby_zero]
    pps = points_per_side
    charges = None
    if uses_charges:
        tc = 1.0 if target_charge is None else target_charge
        charges = np.tile(np.array([1.0, tc]), ((pps + 1) ** 3, 1))
    half = system_length / 2.0
    for cell in far_cells:
        assert cells_relative(dimension, per_side, system_length, cell, 0) == cell
        lower = [cmin[cell][d] - cmax[0][d] for d in range(dimension)]
        upper = [cmax[cell][d] - cmin[0][d] for d in range(dimension)]
        # inner_point_estimator.py:139-147, same expression order
        axes = []
        for d in range(dimension):
            axes.append(np.array([lower[d] + (upper[d] - lower[d]) * i / pps for i in range(pps + 1)]))
        px, py, pz = np.meshgrid(axes[0], axes[1], axes[2], indexing="ij")
        seps = np.stack([px.ravel(), py.ravel(), pz.ravel()], axis=1)
        # correct_separation, hypercubic_setting.py:172 (np.mod has Python's sign convention)
        seps = np.mod(seps + half, system_length) - half
        for direction in range(dimension):
            der = potential_derivative_batch(pot, dimension, system_length, direction, seps, charges)
            upper_bound = -float("inf")
            lower_bound = float("inf")
            for value in der:  # max/min accumulate exactly like the reference's loop
                upper_bound = max(upper_bound, float(value))
                lower_bound = min(lower_bound, float(value))
            if upper_bound > 0.0:
                upper_bound *= prefactor
            else:
                upper_bound /= prefactor
            if lower_bound > 0.0:
                lower_bound /= prefactor
            else:
                lower_bound *= prefactor
            ub = min(empirical_bound, upper_bound)
            lb = max(-empirical_bound, lower_bound)
            bounds[cell][direction][0] = ub
            bounds[cell][direction][1] = -lb
    return bounds, far_cells


def walker_table(items, rates):
    """Walker.__init__ + _build_table (walker.py:55-103). Returns dict with cell_a, cell_b, rate_a, total, mean."""
    rates = [float(r) for r in rates]
    total_rate = sum(rates)
    mean_rate = total_rate / len(items)
    small, large = [], []
    for item, rate in zip(items, rates):
        entry = [item, rate]
        (large if rate > mean_rate else small).append(entry)
    cell_a, cell_b, rate_a = [], [], []
    while len(small) and len(large):
        s = small.pop()
        l = large.pop()
        cell_a.append(s[0])
        rate_a.append(s[1])
        cell_b.append(l[0])
        l[1] -= mean_rate - s[1]
        if l[1] < mean_rate:
            small.append(l)
        else:
            large.append(l)
    while len(small):
        cell_a.append(small.pop()[0])
        rate_a.append(mean_rate)
        cell_b.append(-1)
    while len(large):
        cell_a.append(large.pop()[0])
        rate_a.append(mean_rate)
        cell_b.append(-1)
    return {"cell_a": np.array(cell_a, dtype=np.int32), "cell_b": np.array(cell_b, dtype=np.int32),
            "rate_a": np.array(rate_a, dtype=np.float64), "total_rate": total_rate, "mean_rate": mean_rate}


def veto_tables(bounds, far_cells, dimension=3):
    """CellVetoEventHandler.initialize second half (cell_veto_event_handler.py:147-158)."""
    upper, lower = [], []
    for direction in range(dimension):
        upper.append(walker_table(far_cells, [max(bounds[c][direction][0], 0.0) for c in far_cells]))
        lower.append(walker_table(far_cells, [max(bounds[c][direction][1], 0.0) for c in far_cells]))
    return {"upper": upper, "lower": lower, "bounds": np.ascontiguousarray(bounds, dtype=np.float64)}


# the program assembly is plain data and shared with the product package
from jellyfysh_b200.program import ProgramBuilder  # noqa: E402,F401


class OracleChain:
    """One Markov chain advanced by the C oracle (orc_chain_* in ecmc_oracle.c)."""

    def __init__(self, builder: ProgramBuilder):
        self._builder = builder
        self._lib = lib()
        self._h = self._lib.orc_chain_create(C.byref(builder.program))
        if not self._h:
            raise RuntimeError("orc_chain_create failed")
        p = builder.program
        self.dimension, self.n = p.dimension, p.n_particles
        self.n_cells, self.max_occ, self.max_surplus = builder.n_cells, p.max_occupants, p.max_surplus

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.orc_chain_destroy(self._h)
            self._h = None

    def set_positions(self, positions, charges=None):
        pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(self.n, self.dimension)
        ch = None if charges is None else np.ascontiguousarray(charges, dtype=np.float64)
        self._lib.orc_chain_set_positions(self._h, pos.ctypes.data, None if ch is None else ch.ctypes.data)

    def positions(self):
        out = np.empty((self.n, self.dimension))
        self._lib.orc_chain_get_positions(self._h, out.ctypes.data)
        return out

    def set_roots(self, roots):
        npr = max(int(self._builder.program.nodes_per_root), 1)
        arr = np.ascontiguousarray(roots, dtype=np.float64).reshape(self.n // npr, self.dimension)
        self._lib.orc_chain_set_roots(self._h, arr.ctypes.data)

    def roots(self):
        npr = max(int(self._builder.program.nodes_per_root), 1)
        out = np.empty((self.n // npr, self.dimension))
        self._lib.orc_chain_get_roots(self._h, out.ctypes.data)
        return out

    def start(self, stream=0):
        self._lib.orc_chain_start(self._h, stream)

    def state(self):
        st = abi.EcmcChainState()
        self._lib.orc_chain_get_state(self._h, C.byref(st))
        return st

    def set_state(self, st):
        self._lib.orc_chain_set_state(self._h, C.byref(st))

    def cells(self):
        occ = np.empty((self.n_cells, self.max_occ), dtype=np.int32)
        sur = np.full(max(self.max_surplus, 1), -1, dtype=np.int32)
        n = C.c_int32()
        self._lib.orc_chain_get_cells(self._h, occ.ctypes.data, sur.ctypes.data, C.byref(n))
        return occ, sur[:n.value].copy()

    def set_cells(self, occ, surplus):
        occ = np.ascontiguousarray(occ, dtype=np.int32)
        sur = np.ascontiguousarray(surplus, dtype=np.int32)
        self._lib.orc_chain_set_cells(self._h, occ.ctypes.data, sur.ctypes.data, len(sur))

    def stats(self):
        s = abi.EcmcStats()
        self._lib.orc_chain_get_stats(self._h, C.byref(s))
        return s.as_dict()

    def run(self, until=(float("inf"), float("inf")), max_events=0, record=0):
        rec = np.zeros(record, dtype=abi.record_dtype()) if record else None
        n = self._lib.orc_chain_run(self._h, until[0], until[1], max_events,
                                    None if rec is None else rec.ctypes.data, record)
        return int(n), (rec[:min(n, record)] if rec is not None else None)
