"""Product-side init tables (jellyfysh_b200/tables.py) against the oracle's restatement, which is pinned to the
reference (tests/test_oracle_traces.py::test_init_tables_bit_exact). CPU part: geometry and Walker pairing are exact.
GPU part: derivative bounds evaluated on the device agree with the reference's to 1e-12 and give the same tables."""
import numpy as np
import pytest

import trace_util as tu
from jellyfysh_b200 import abi, tables


@pytest.mark.parametrize("dimension,per_side,length,layers", [(3, [12, 12, 12], 12.699208415745595, 1),
                                                              (3, [4, 5, 4], 1.0, 1), (3, [7, 7, 7], 10.0, 2),
                                                              (2, [13, 13], 12.836, 1)])
def test_cell_geometry_matches_oracle(oracle, dimension, per_side, length, layers):
    geometry = tables.CellGeometry(dimension, length, per_side, layers)
    cmin, cmax = oracle.cells_geometry(dimension, per_side, length)
    assert np.array_equal(geometry.cell_min, cmin) and np.array_equal(geometry.cell_max, cmax)
    assert geometry.nearby_of_zero() == sorted(oracle.nearby_cells(dimension, per_side, layers, length, 0))
    assert len(geometry.far_cells()) == geometry.n_cells - (2 * layers + 1) ** dimension


@pytest.mark.parametrize("name", tu.TRACES)
def test_walker_tables_match_reference(name):
    """The alias tables rebuilt from the reference's bounds are the reference's tables, entry for entry."""
    g = tu.load_trace(name)
    ref = tu.reference_tables(g)
    cps = [int(c) for c in g["meta_cells_per_side"]]
    geometry = tables.CellGeometry(3, float(g["meta_system_length"]), cps, 1)
    ours = tables.veto_tables(np.nan_to_num(ref["bounds"], nan=0.0), geometry.far_cells())
    for kind in ("upper", "lower"):
        for d in range(3):
            for key in ("cell_a", "cell_b", "rate_a"):
                assert np.array_equal(ours[kind][d][key], ref[kind][d][key]), (kind, d, key)
            assert ours[kind][d]["total_rate"] == ref[kind][d]["total_rate"]
            assert ours[kind][d]["mean_rate"] == ref[kind][d]["mean_rate"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["trace_lj_small", "trace_coulomb_small"])
def test_device_estimator_matches_reference_bounds(name):
    g = tu.load_trace(name)
    _, potential, _, veto, use_charge = tu.potentials_of(g)
    cps = [int(c) for c in g["meta_cells_per_side"]]
    geometry = tables.CellGeometry(3, float(g["meta_system_length"]), cps, 1)
    prefactor, points = float(g["meta_estimator"][0]), int(g["meta_estimator"][1])
    bounds, far = tables.inner_point_derivative_bounds(veto, geometry, prefactor=prefactor, points_per_side=points,
                                                       charges=(1.0, 1.0) if use_charge else None)
    ref = g["bounds"]
    assert far == [c for c in range(geometry.n_cells) if not np.isnan(ref[c, 0, 0])]
    scale = np.max(np.abs(ref[far]))
    assert np.max(np.abs(bounds[far] - ref[far])) < 1e-12 * scale
