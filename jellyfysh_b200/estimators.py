"""Init-time derivative bounds of the reference's estimators, evaluated on the device.

`CellVetoEventHandler.initialize` (event_handler/abstracts/cell_veto_event_handler.py:128-158) and `CellBoundingPotential`
(potential/cell_bounding_potential.py) ask their estimator for `derivative_bound(lower_corner, upper_corner, direction,
calculate_lower_bound)` once per far cell and direction; the estimator then calls the potential point by point from Python
-- (P + 1)^3 grid points (InnerPointEstimator, inner_point_estimator.py:139-163), the surface of that grid
(BoundaryPointEstimator, boundary_point_estimator.py:108-174), 1000 random dipoles (DipoleMonteCarloEstimator,
dipole_monte_carlo_estimator.py:100-155) or a dipole aligned with the local gradient at every grid point
(DipoleInnerPointEstimator, dipole_inner_point_estimator.py:100-163). For the merged-image Coulomb potential that is the
whole start-up time of a run.

`accelerate(estimator)` replaces `derivative_bound` of ONE reference estimator instance by a function that builds the same
points with numpy (same expressions, same order of operations, same calls of Python's `random` in the same order for the
Monte Carlo estimator), evaluates the potential for all points of the region in one launch of `ecmc_potential_derivative`,
and applies the reference's max / min, prefactor and empirical-bound rules. The instance stays a reference object (the
event handlers keep calling `charge_correction_factor` etc. on it); `accelerate_activator` finds the estimators of a
factory-built activator before `Mediator.__init__` initialises it.
"""
import random

import numpy as np

from jellyfysh_b200 import compiler, engine


class _Region:
    """What every estimator needs of the box and of its own configuration."""

    def __init__(self, estimator, device):
        from jellyfysh import setting
        from jellyfysh.setting import hypercubic_setting
        if not hypercubic_setting.initialized():
            raise compiler._configuration_error("device estimators need the hypercubic setting")
        self.dimension = int(setting.dimension)
        self.length = float(hypercubic_setting.system_length)
        self.periodic = estimator._correct_separation.__name__ != "<lambda>"
        self.record = compiler.potential_descriptor(estimator._potential)
        self.device = device
        if self.dimension != 3:
            raise compiler._configuration_error("device estimators cover three dimensions")

    def correct(self, points):
        """setting.periodic_boundaries.correct_separation on every row (hypercubic_setting.py: ((s + L/2) % L) - L/2)."""
        if not self.periodic:
            return points
        half = self.length / 2.0
        return np.mod(points + half, self.length) - half

    def derivative(self, direction, points, charges):
        pair = None if charges is None else np.broadcast_to(np.asarray(charges, dtype=np.float64), (len(points), 2))
        return engine.potential_derivative(self.record, 3, self.length, direction, self.correct(points),
                                           None if pair is None else np.ascontiguousarray(pair), device=self.device)


def _signed_bounds(estimator, upper, lower, calculate_lower_bound):
    """inner_point_estimator.py:151-163 / boundary_point_estimator.py:161-174"""
    upper = upper * estimator._prefactor if upper > 0.0 else upper / estimator._prefactor
    lower = lower / estimator._prefactor if lower > 0.0 else lower * estimator._prefactor
    if calculate_lower_bound:
        return [min(estimator._empirical_bound, upper), max(-estimator._empirical_bound, lower)]
    return [min(estimator._empirical_bound, upper)]


def _grid_axes(lower_corner, upper_corner, points_per_side):
    steps = np.arange(points_per_side + 1)
    # lower + (upper - lower) * i / P, the reference's expression
    return [lower_corner[d] + (upper_corner[d] - lower_corner[d]) * steps / points_per_side for d in range(3)]


def _inner_point(estimator, region):
    charges = estimator._charges if estimator._number_charges == 2 else None

    def derivative_bound(lower_corner, upper_corner, direction, calculate_lower_bound=False):
        axes = _grid_axes(lower_corner, upper_corner, estimator._points_per_side)
        points = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, 3)
        values = region.derivative(direction, points, charges)
        return _signed_bounds(estimator, float(values.max()), float(values.min()), calculate_lower_bound)
    return derivative_bound


def _boundary_point(estimator, region):
    charges = estimator._charges if estimator._number_charges == 2 else None

    def derivative_bound(lower_corner, upper_corner, direction, calculate_lower_bound=False):
        p = estimator._points_per_side
        lower, upper = np.asarray(lower_corner, dtype=np.float64), np.asarray(upper_corner, dtype=np.float64)
        # index / P * (upper - lower) + lower, the reference's expression (boundary_point_estimator.py:147-156)
        grid = np.stack(np.meshgrid(*[np.arange(p + 1)] * 3, indexing="ij"), axis=-1).reshape(-1, 3)
        surface = grid[np.any((grid == 0) | (grid == p), axis=1)]
        points = surface / p * (upper - lower) + lower
        values = region.derivative(direction, points, charges)
        return _signed_bounds(estimator, float(values.max()), float(values.min()), calculate_lower_bound)
    return derivative_bound


def _dipole_bounds(estimator, largest, calculate_lower_bound):
    upper = largest * estimator._prefactor
    if calculate_lower_bound:
        return [min(estimator._empirical_bound, upper), max(-estimator._empirical_bound, -upper)]
    return [min(estimator._empirical_bound, upper)]


def _dipole_pair_sum(estimator, region, direction, first, second):
    """|dU/dx(position1, 1, +q) + dU/dx(position2, 1, -q)| for rows of positions, both in one launch."""
    q = estimator._dipole_charge
    points = np.concatenate([first, second])
    charges = np.concatenate([np.tile([1.0, q], (len(first), 1)), np.tile([1.0, -q], (len(second), 1))])
    values = engine.potential_derivative(region.record, 3, region.length, direction, region.correct(points), charges,
                                         device=region.device)
    return values[:len(first)], values[len(first):]


def _dipole_monte_carlo(estimator, region):
    def derivative_bound(lower_corner, upper_corner, direction, calculate_lower_bound=False):
        half = estimator._dipole_separation_over_two
        lower = [lower_corner[d] - half for d in range(3)]
        upper = [upper_corner[d] + half for d in range(3)]
        centers = np.empty((estimator._number_trials, 3))
        directions = np.empty((estimator._number_trials, 3))
        for trial in range(estimator._number_trials):
            # the reference's draws in the reference's order (dipole_monte_carlo_estimator.py:139-141,
            # base/vectors.py:242-262): three uniforms for the centre, then rejection sampling in the unit ball
            centers[trial] = [random.uniform(lower[d], upper[d]) for d in range(3)]
            while True:
                vector = [random.uniform(-1, 1) for _ in range(3)]
                norm = sum(entry * entry for entry in vector) ** 0.5  # base/vectors.py:43
                if 0.0 < norm <= 1.0:
                    break
            directions[trial] = [entry / norm for entry in vector]
        one, two = _dipole_pair_sum(estimator, region, direction, centers + directions * half, centers - directions * half)
        return _dipole_bounds(estimator, max(0.0, float(np.max(np.abs(one + two)))), calculate_lower_bound)
    return derivative_bound


def _dipole_inner_point(estimator, region):
    def derivative_bound(lower_corner, upper_corner, direction, calculate_lower_bound=False):
        half = estimator._dipole_separation_over_two
        lower = [lower_corner[d] - half for d in range(3)]
        upper = [upper_corner[d] + half for d in range(3)]
        axes = _grid_axes(lower, upper, estimator._max_index_per_side)
        centers = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, 3)
        delta = estimator._dipole_separation / 20
        gradient = np.empty_like(centers)
        for d in range(3):
            shift = np.zeros(3)
            shift[d] = delta
            one, two = _dipole_pair_sum(estimator, region, direction, centers + shift, centers - shift)
            gradient[:, d] = (one + two) / delta / 2
        gradient /= np.sqrt(np.sum(gradient * gradient, axis=1))[:, None]
        one, two = _dipole_pair_sum(estimator, region, direction, centers + gradient * half, centers - gradient * half)
        return _dipole_bounds(estimator, max(0.0, float(np.max(np.abs(one + two)))), calculate_lower_bound)
    return derivative_bound


_BUILDERS = (("DipoleMonteCarloEstimator", _dipole_monte_carlo), ("DipoleInnerPointEstimator", _dipole_inner_point),
             ("BoundaryPointEstimator", _boundary_point), ("InnerPointEstimator", _inner_point))


def accelerate(estimator, device=0):
    """Give one reference estimator instance a device-backed `derivative_bound`; returns False for estimator or
    potential types without a device path (the instance is then left alone)."""
    names = compiler._class_names(estimator)
    for name, builder in _BUILDERS:
        if name in names:
            try:
                region = _Region(estimator, device)
            except Exception:  # noqa: BLE001 - e.g. a potential without a device implementation
                return False
            estimator.derivative_bound = builder(estimator, region)
            return True
    return False


def accelerate_activator(activator, device=0):
    """Every estimator reachable from a factory-built TagActivator -- cell-veto handlers (`_estimator`) and cell-bounding
    potentials (`_bounding_potential._estimator`) of all taggers' event handlers. To be called before the activator is
    initialised (Mediator.__init__), which is when the bounds are computed. Returns the number of estimators changed."""
    seen, changed = set(), 0
    for tagger in getattr(activator, "_taggers", []):
        # before Tagger.initialize the tagger holds the one handler it will deep-copy (tagger.py:104,159)
        handlers = list(getattr(tagger, "_event_handlers", None) or [])
        template = getattr(tagger, "_event_handler_to_copy", None)
        if template is not None:
            handlers.append(template)
        for handler in handlers:
            for owner in (handler, getattr(handler, "_bounding_potential", None)):
                estimator = getattr(owner, "_estimator", None)
                if estimator is not None and id(estimator) not in seen:
                    seen.add(id(estimator))
                    changed += bool(accelerate(estimator, device))
    return changed
