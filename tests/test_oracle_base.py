"""Pin the oracle's Time arithmetic, periodic boundaries and cell geometry to outputs of the running reference
(tests/golden/base.npz; reference: jellyfysh/base/time.py, jellyfysh/setting/hypercubic_setting.py,
jellyfysh/activator/internal_state/cell_occupancy/cells/cuboid_periodic_cells.py)."""
import numpy as np
import pytest

import kat_replay as kr


def test_time_arithmetic_bit_exact(oracle):
    g = kr.load_npz("base")
    for q, r, dt, added, sub, ff in zip(g["time_q"], g["time_r"], g["time_dt"], g["time_added"], g["time_sub"],
                                        g["time_from_float"]):
        assert oracle.time_add(q, r, dt) == (added[0], added[1])
        assert oracle.time_sub(added[0], added[1], q, r) == sub
        assert oracle.time_from_float(q + r) == (ff[0], ff[1])
    assert oracle.time_add(3.0, 0.5, float("inf")) == (float("inf"), float("inf"))


@pytest.mark.parametrize("geo", range(6))
def test_cells_and_boundaries_bit_exact(oracle, geo):
    g = kr.load_npz("base")
    params = g[f"geo{geo}_params"]
    dim, length, nl = int(params[0]), float(params[1]), int(params[2])
    cps = [int(c) for c in params[3:3 + dim]]
    cmin, cmax = oracle.cells_geometry(dim, cps, length)
    limit = len(g[f"geo{geo}_cell_min"])
    assert np.array_equal(cmin[:limit], g[f"geo{geo}_cell_min"])
    assert np.array_equal(cmax[:limit], g[f"geo{geo}_cell_max"])
    for pos, cell in zip(g[f"geo{geo}_pos"], g[f"geo{geo}_pos_cell"]):
        assert oracle.position_to_cell(dim, cps, length, pos) == cell
    for k, probe in enumerate(g[f"geo{geo}_probe"]):
        assert sorted(oracle.nearby_cells(dim, cps, nl, length, int(probe))) == list(g[f"geo{geo}_nearby"][k])
        rel = int(g[f"geo{geo}_rel"][k])
        assert oracle.cells_translate(dim, cps, length, int(probe), rel) == g[f"geo{geo}_translate"][k]
        assert oracle.cells_relative(dim, cps, length, int(probe), rel) == g[f"geo{geo}_relative"][k]
    lib = oracle.lib()
    for s, so, po in zip(g[f"geo{geo}_sep_in"], g[f"geo{geo}_sep_out"], g[f"geo{geo}_pos_out"]):
        assert lib.orc_correct_separation_entry(s, length) == so
        assert lib.orc_correct_position_entry(s, length) == po
