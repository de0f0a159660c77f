"""Init-time tables of the cell-veto handler, built with the device's batched potential arithmetic.

What the reference computes once at start-up:
* derivative bounds per far relative cell and direction: CellVetoEventHandler.initialize
  (jellyfysh/event_handler/abstracts/cell_veto_event_handler.py:134-159) with InnerPointEstimator.derivative_bound
  (jellyfysh/estimator/inner_point_estimator.py:108-163): the derivative on an even grid of (P+1)^3 separations
  between the cell pair, max / min, widened by the prefactor, clipped by the empirical bound;
* Walker's alias tables over the far cells (jellyfysh/event_handler/walker.py:55-103).
All derivative evaluations of all cells go to the GPU in one batch per direction (ecmc_potential_derivative);
the max/min reduction and the table pairing are cheap host work. No CPU fallback for the derivative.
"""
import numpy as np

from jellyfysh_b200 import engine


# ---- CuboidPeriodicCells geometry ---------------------------------------------------------------------------
def axis_geometry(n_cells: int, side: float):
    """(cell_min, cell_max) per cell index of one axis: the smallest / largest double x with int(x / side) == index
    (jellyfysh/activator/internal_state/cell_occupancy/cells/cuboid_cells.py:119-131)."""
    cell_min, cell_max = np.zeros(n_cells), np.zeros(n_cells)
    for index in range(n_cells):
        lower, upper = index * side, (index + 1) * side
        if lower > 0.0:
            while int(lower / side) >= index:
                lower = float(np.nextafter(lower, -np.inf))
            while int(lower / side) < index:
                lower = float(np.nextafter(lower, np.inf))
        while int(upper / side) <= index:
            upper = float(np.nextafter(upper, np.inf))
        while int(upper / side) > index:
            upper = float(np.nextafter(upper, -np.inf))
        cell_min[index], cell_max[index] = lower, upper
    return cell_min, cell_max


class CellGeometry:
    """Flat-index view of a periodic cuboid cell system: cell = sum_d index_d * prod_{d' < d} cells_per_side[d']."""

    def __init__(self, dimension, system_length, cells_per_side, neighbor_layers):
        self.dimension = dimension
        self.length = float(system_length)
        self.per_side = [int(c) for c in cells_per_side[:dimension]]
        self.neighbor_layers = int(neighbor_layers)
        self.n_cells = int(np.prod(self.per_side))
        self.cumulative = [int(np.prod(self.per_side[:d])) for d in range(dimension)]
        axes = [axis_geometry(n, self.length / n) for n in self.per_side]
        index = self.cell_indices(np.arange(self.n_cells))
        self.cell_min = np.stack([axes[d][0][index[:, d]] for d in range(dimension)], axis=1)
        self.cell_max = np.stack([axes[d][1][index[:, d]] for d in range(dimension)], axis=1)

    def cell_indices(self, cells):
        cells = np.asarray(cells)
        return np.stack([(cells // self.cumulative[d]) % self.per_side[d] for d in range(self.dimension)], axis=-1)

    def nearby_of_zero(self):
        """Flat indices of the nearby (excluded) cells of cell zero (cuboid_periodic_cells.py:74-100)."""
        width = 2 * self.neighbor_layers + 1
        offsets = np.stack(np.meshgrid(*[np.arange(width) - self.neighbor_layers] * self.dimension, indexing="ij"),
                           axis=-1).reshape(-1, self.dimension)
        wrapped = np.mod(offsets, np.array(self.per_side))
        return sorted(set(int(c) for c in wrapped @ np.array(self.cumulative)))

    def far_cells(self):
        nearby = set(self.nearby_of_zero())
        return [cell for cell in range(self.n_cells) if cell not in nearby]


# ---- estimator -------------------------------------------------------------------------------------------------
def inner_point_derivative_bounds(potential, geometry: CellGeometry, prefactor=1.5, points_per_side=10,
                                  empirical_bound=float("inf"), charges=None, device=0):
    """bounds[n_cells][dimension][2] = (upper bound, -lower bound) for every far relative cell (zeros elsewhere).

    charges: None for potentials without charges, else (1.0, target_charge) as InnerPointEstimator passes them.
    """
    if geometry.dimension != 3:
        raise ValueError("the inner point estimator of the reference covers three dimensions only")
    far = geometry.far_cells()
    points = points_per_side
    lower = geometry.cell_min[far] - geometry.cell_max[0]
    upper = geometry.cell_max[far] - geometry.cell_min[0]
    steps = np.arange(points + 1)
    # lower + (upper - lower) * i / P, same expression order as the reference
    axes = [lower[:, d, None] + (upper[:, d, None] - lower[:, d, None]) * steps[None, :] / points for d in range(3)]
    grid = np.stack(np.broadcast_arrays(axes[0][:, :, None, None], axes[1][:, None, :, None], axes[2][:, None, None, :]),
                    axis=-1).reshape(len(far), -1, 3)
    half = geometry.length / 2.0
    separations = np.mod(grid + half, geometry.length) - half
    flat = separations.reshape(-1, 3)
    pair_charges = None if charges is None else np.tile(np.asarray(charges, dtype=np.float64), (len(flat), 1))
    bounds = np.zeros((geometry.n_cells, 3, 2))
    for direction in range(3):
        derivative = engine.potential_derivative(potential, 3, geometry.length, direction, flat, pair_charges,
                                                 device=device).reshape(len(far), -1)
        upper_bound, lower_bound = derivative.max(axis=1), derivative.min(axis=1)
        upper_bound = np.where(upper_bound > 0.0, upper_bound * prefactor, upper_bound / prefactor)
        lower_bound = np.where(lower_bound > 0.0, lower_bound / prefactor, lower_bound * prefactor)
        bounds[far, direction, 0] = np.minimum(empirical_bound, upper_bound)
        bounds[far, direction, 1] = -np.maximum(-empirical_bound, lower_bound)
    return bounds, far


# ---- Walker ------------------------------------------------------------------------------------------------------
def walker_table(items, rates):
    """Alias table of Walker.__init__ / _build_table (walker.py:55-103): entries (cell_a, rate_a, cell_b), drawn as
    `cell_a if uniform(0, mean) <= rate_a else cell_b`. The pairing order (two stacks, popped from the end) decides
    which entry a random index refers to, so it follows the reference exactly."""
    rates = [float(rate) for rate in rates]
    total_rate = sum(rates)
    mean_rate = total_rate / len(items)
    small = [[item, rate] for item, rate in zip(items, rates) if not rate > mean_rate]
    large = [[item, rate] for item, rate in zip(items, rates) if rate > mean_rate]
    cell_a, cell_b, rate_a = [], [], []
    while small and large:
        low, high = small.pop(), large.pop()
        cell_a.append(low[0])
        rate_a.append(low[1])
        cell_b.append(high[0])
        high[1] -= mean_rate - low[1]
        (small if high[1] < mean_rate else large).append(high)
    for rest in (small, large):
        while rest:
            cell_a.append(rest.pop()[0])
            rate_a.append(mean_rate)
            cell_b.append(-1)
    return {"cell_a": np.array(cell_a, dtype=np.int32), "cell_b": np.array(cell_b, dtype=np.int32),
            "rate_a": np.array(rate_a, dtype=np.float64), "total_rate": total_rate, "mean_rate": mean_rate}


def veto_tables(bounds, far_cells, dimension=3):
    """Upper / lower Walker tables per direction (cell_veto_event_handler.py:147-158) in ProgramBuilder.set_veto form."""
    upper = [walker_table(far_cells, np.maximum(bounds[far_cells, d, 0], 0.0)) for d in range(dimension)]
    lower = [walker_table(far_cells, np.maximum(bounds[far_cells, d, 1], 0.0)) for d in range(dimension)]
    return {"upper": upper, "lower": lower, "bounds": np.ascontiguousarray(bounds, dtype=np.float64)}
