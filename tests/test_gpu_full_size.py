"""The BASELINE.json configurations at their full sizes, checked through size-independent properties (the oracle would
need hours): bookkeeping invariants of the cell occupancy, determinism, independence of the launch partition, time
ordering, and statistics that must add up. C2: 4096 chains x 1024 Lennard-Jones particles; C3: Coulomb atoms N = 64
and 512; C5: one chain of 65536 particles."""
import numpy as np
import pytest

from jellyfysh_b200 import engine, tables, workloads

pytestmark = pytest.mark.gpu


def check_invariants(eng, length, cells_per_side, tag):
    """Every particle is exactly once active, an occupant of the cell its position lies in, or in the surplus list."""
    positions = eng.download_positions()
    occupants, surplus = eng.cells()
    states = eng.chain_states()
    n_chains, n = positions.shape[:2]
    assert np.all(positions >= 0.0) and np.all(positions < length), tag
    side = length / cells_per_side
    cell_of = np.zeros((n_chains, n), dtype=np.int64)
    for d in range(positions.shape[2]):
        cell_of += (positions[:, :, d] / side).astype(np.int64) * cells_per_side ** d
    for c in range(n_chains):
        occ = occupants[c].ravel()
        present = occ[occ >= 0]
        members = np.concatenate([present, surplus[c], [states[c]["active"]]])
        assert len(members) == n and len(np.unique(members)) == n, (tag, c)
        cells = np.repeat(np.arange(occupants.shape[1]), occupants.shape[2])[occ >= 0]
        assert np.array_equal(cell_of[c, present], cells), (tag, c)
        assert states[c]["active_cell"] == cell_of[c, states[c]["active"]], (tag, c)
    return positions, states


def test_c2_full_size_properties():
    n_chains, n, cells, events = 4096, 1024, 12, 600
    builder, length = workloads.lennard_jones(n_particles=n, cells_per_side=cells)
    start = workloads.lattice_start(n_chains, n, cells, length)
    with engine.Engine(builder, n_chains=n_chains) as one, engine.Engine(builder, n_chains=n_chains) as two:
        for eng in (one, two):
            eng.upload_positions(start)
            eng.start(first_stream=0)
        one.run(max_events=events)
        stats_one = one.sync()
        for _ in range(3):  # the same events in three launches
            two.run(max_events=events // 3)
        stats_two = two.sync()
        # (candidates: with candidate pruning the count of EVALUATED candidates depends on where a launch rebuilt its
        # candidate list; the committed events do not)
        assert {k: v for k, v in stats_one.items() if k != "candidates"} == \
            {k: v for k, v in stats_two.items() if k != "candidates"}
        assert stats_one["events"] == n_chains * events
        assert stats_one["events"] == stats_one["pair_events"] + stats_one["veto_events"] + \
            stats_one["boundary_events"] + stats_one["end_of_chain_events"]
        assert stats_one["capacity_errors"] == 0 and stats_one["bound_violations"] == 0
        assert stats_one["veto_accepted"] <= stats_one["veto_events"]
        positions_one, states_one = check_invariants(one, length, cells, "C2")
        assert np.array_equal(positions_one, two.download_positions())
        assert np.array_equal(states_one, two.chain_states())
        assert np.all(states_one["event_counter"] == events)
        # chains are independent: different streams give different trajectories, the same stream the same one
        assert len(np.unique(states_one["time_r"])) > n_chains // 2
        moved = np.any(positions_one != start, axis=2).sum(axis=1)
        assert np.all(moved >= 1) and np.all(moved <= events + 1)
    # a time-limited run ends every chain exactly at the limit and leaves at most one kept candidate per chain
    with engine.Engine(builder, n_chains=256) as eng:
        eng.upload_positions(start[:256])
        eng.start(first_stream=0)
        eng.run(until=(3.0, 0.25))
        eng.sync()
        states = eng.chain_states()
        assert np.all(states["time_q"] == 3.0) and np.all(states["time_r"] == 0.25)
        assert np.all((states["pending_q"] > 3.0) | ((states["pending_q"] == 3.0) & (states["pending_r"] >= 0.25)))


def test_c2_first_chains_match_a_small_batch():
    """Chain c of the big batch is the same Markov chain as chain c run alone (same stream, same start)."""
    n, cells, events = 1024, 12, 300
    builder, length = workloads.lennard_jones(n_particles=n, cells_per_side=cells)
    start = workloads.lattice_start(512, n, cells, length)
    with engine.Engine(builder, n_chains=512) as big, engine.Engine(builder, n_chains=3) as small:
        big.upload_positions(start)
        big.start(first_stream=100)
        small.upload_positions(start[200:203])
        small.start(first_stream=300)
        big.run(max_events=events)
        small.run(max_events=events)
        big.sync(), small.sync()
        assert np.array_equal(big.download_positions()[200:203], small.download_positions())
        for name in ("active", "direction", "time_q", "time_r", "eoc_next_active", "active_cell"):
            assert np.array_equal(big.chain_states()[name][200:203], small.chain_states()[name])


@pytest.mark.parametrize("n", [64, 512])
def test_c3_coulomb_atoms_properties(n):
    n_chains = 256 if n == 64 else 64
    builder, length = workloads.coulomb_atoms(n_particles=n, points_per_side=4)
    cells = builder.program.cells_per_side[0]
    start = workloads.uniform_start(n_chains, n, length)
    charges = np.ones((n_chains, n))
    with engine.Engine(builder, n_chains=n_chains) as eng:
        eng.upload_positions(start, charges)
        eng.start(first_stream=0)
        eng.run(max_events=400)
        stats = eng.sync()
        assert stats["events"] == n_chains * 400 and stats["capacity_errors"] == 0
        assert stats["veto_events"] > stats["pair_events"] > 0
        # the inverse-power bound with prefactor 1.5837 and the cell bounds hold for like charges
        assert stats["bound_violations"] == 0
        check_invariants(eng, length, cells, "C3 N=%d" % n)


def test_c5_single_large_chain_properties():
    n, cells, events = 65536, 48, 4000
    builder, length = workloads.lennard_jones(n_particles=n, cells_per_side=cells)
    start = workloads.lattice_start(1, n, cells, length)
    with engine.Engine(builder, n_chains=1) as eng:
        eng.upload_positions(start)
        eng.start(first_stream=0)
        records, stats = eng.run_recorded(max_events=events, records_per_chain=events)
        assert stats["events"] == events and stats["capacity_errors"] == 0
        rec = records[0]
        times = rec["time_q"] + rec["time_r"]
        assert np.all(np.diff(times) >= 0.0)  # the scheduler never goes back in time
        assert np.all(rec["n_candidates"] >= 3) and np.all(rec["n_candidates"] <= 27 + 3 + 64)
        pairs = rec[rec["kind"] == 1]
        assert np.all(pairs["accepted"] == 1) and np.all(pairs["new_active"] == pairs["target"])
        check_invariants(eng, length, cells, "C5")


def test_device_tables_for_c2_match_oracle_tables(oracle):
    """The bench's cell-veto tables (estimator on the device) against the oracle's restatement of the reference's."""
    builder, length = workloads.lennard_jones(n_particles=1024, cells_per_side=12)
    potential = builder.program.veto_potential
    bounds, far = oracle.inner_point_derivative_bounds(potential, length, [12] * 3, 1, prefactor=1.5, points_per_side=4)
    ours = builder.tables["bounds"]
    assert far == tables.CellGeometry(3, length, [12] * 3, 1).far_cells()
    scale = np.max(np.abs(bounds[far]))
    assert np.max(np.abs(ours[far] - bounds[far])) < 1e-12 * scale
    reference_tables = oracle.veto_tables(bounds, far)
    for d in range(3):
        same = np.mean(reference_tables["upper"][d]["cell_a"] == builder.tables["upper"][d]["cell_a"])
        assert same > 0.99  # the alias pairing only differs where two bounds are equal to the last bit


def test_separation_histogram_matches_numpy():
    """ecmc_separation_histogram against numpy on the downloaded positions: all pairs of every chain."""
    n_chains, n, cells = 24, 200, 6
    builder, length = workloads.lennard_jones(n_particles=n, cells_per_side=cells, points_per_side=2)
    positions = workloads.lattice_start(n_chains, n, cells, length, jitter=0.2)
    with engine.Engine(builder, n_chains=n_chains) as eng:
        eng.upload_positions(positions)
        eng.start()
        eng.run(max_events=500)
        eng.sync()
        r_max = length * np.sqrt(3.0) / 2.0
        ours = eng.separation_histogram(1000, 0.0, r_max)
        eng.separation_histogram(1000, 0.0, r_max, out=ours)  # accumulates
        current = eng.download_positions()
    expected = np.zeros(1000, dtype=np.int64)
    half = length / 2.0
    iu = np.triu_indices(n, k=1)
    for c in range(n_chains):
        sep = current[c][iu[1]] - current[c][iu[0]]
        sep = np.mod(sep + half, length) - half
        expected += np.histogram(np.sqrt(np.sum(sep * sep, axis=1)), bins=1000, range=(0.0, r_max))[0]
    assert ours.sum() == 2 * n_chains * n * (n - 1) // 2
    # a separation within one ulp of a bin edge may fall on either side (fma vs numpy's rounding): a handful at most
    assert np.abs(ours.astype(np.int64) - 2 * expected).sum() <= 8
    # N > tile size exercises the tiling: 1500 particles, one chain
    builder, length = workloads.lennard_jones(n_particles=1500, cells_per_side=12, points_per_side=2, veto=False)
    positions = workloads.lattice_start(1, 1500, 12, length)
    with engine.Engine(builder, n_chains=1) as eng:
        eng.upload_positions(positions)
        counts = eng.separation_histogram(64, 0.0, length)
    assert counts.sum() == 1500 * 1499 // 2


def test_subset_separation_histogram_matches_numpy():
    """ecmc_separation_histogram_subset: every third particle starting at 1 (the oxygens of water molecules stored as
    H, O, H) against numpy."""
    n_chains, n, cells = 7, 201, 6
    builder, length = workloads.lennard_jones(n_particles=n, cells_per_side=cells, points_per_side=2, veto=False)
    positions = workloads.lattice_start(n_chains, n, cells, length, jitter=0.2)
    with engine.Engine(builder, n_chains=n_chains) as eng:
        eng.upload_positions(positions)
        r_max = length * np.sqrt(3.0) / 2.0
        ours = eng.separation_histogram(500, 0.0, r_max, first=1, stride=3)
    subset = positions[:, 1::3]
    m = subset.shape[1]
    iu = np.triu_indices(m, k=1)
    half = length / 2.0
    expected = np.zeros(500, dtype=np.int64)
    for c in range(n_chains):
        sep = np.mod(subset[c][iu[1]] - subset[c][iu[0]] + half, length) - half
        expected += np.histogram(np.sqrt(np.sum(sep * sep, axis=1)), bins=500, range=(0.0, r_max))[0]
    assert ours.sum() == n_chains * m * (m - 1) // 2
    assert np.abs(ours.astype(np.int64) - expected).sum() <= 4


def test_polarization_and_bond_histograms_match_numpy():
    """ecmc_polarization and ecmc_bond_histograms on water molecules after some events, against numpy on the downloaded
    configuration: the observables of PolarizationOutputHandler (polarization_output_handler.py:76-101 with
    base/node.py:164-188) and BondLengthAndAngleOutputHandler (bond_length_and_angle_output_handler.py:77-103)."""
    import configs
    import trace_util as tu
    from jellyfysh_b200.program import ProgramBuilder
    n_chains, n_molecules = 12, 16
    g = dict(tu.load_trace("trace_water"))
    g["meta_n"] = np.asarray(3 * n_molecules)
    builder = tu.water_builder_of(g, ProgramBuilder)
    length = 10.0
    roots = np.empty((n_chains, n_molecules, 3))
    leaves = np.empty((n_chains, 3 * n_molecules, 3))
    for c in range(n_chains):
        r, l = configs.water_start(n_molecules, length, seed=40 + c)
        roots[c], leaves[c] = r, l.reshape(-1, 3)
    charges = np.tile([0.41, -0.82, 0.41], (n_chains, n_molecules))
    with engine.Engine(builder, n_chains=n_chains) as eng:
        eng.upload_positions(leaves, charges)
        eng.upload_roots(roots)
        eng.start(first_stream=0)
        eng.run(max_events=400)
        eng.sync()
        polarization = eng.polarization()
        other = np.linspace(-1.0, 1.0, 3 * n_molecules)
        polarization_other = eng.polarization(other)
        lengths, angles = eng.bond_histograms(200, (0.5, 1.5), (1.0, 2.8))
        eng.bond_histograms(200, (0.5, 1.5), (1.0, 2.8), out=(lengths, angles))  # accumulates
        now, now_roots = eng.download_positions(), eng.download_roots()
    half = length / 2.0
    closest = np.repeat(now_roots, 3, axis=1) + (np.mod(now - np.repeat(now_roots, 3, axis=1) + half, length) - half)
    assert np.max(np.abs(polarization - np.einsum("cn,cnd->cd", charges, closest))) < 1e-11
    assert np.max(np.abs(polarization_other - np.einsum("n,cnd->cd", other, closest))) < 1e-11
    molecules = now.reshape(n_chains, n_molecules, 3, 3)
    one = np.mod(molecules[:, :, 0] - molecules[:, :, 1] + half, length) - half
    two = np.mod(molecules[:, :, 2] - molecules[:, :, 1] + half, length) - half
    n_one, n_two = np.linalg.norm(one, axis=2), np.linalg.norm(two, axis=2)
    expected_lengths = np.histogram(np.concatenate([n_one.ravel(), n_two.ravel()]), bins=200, range=(0.5, 1.5))[0]
    expected_angles = np.histogram(np.arccos(np.sum(one * two, axis=2) / (n_one * n_two)).ravel(), bins=200, range=(1.0, 2.8))[0]
    assert lengths.sum() == 2 * 2 * n_chains * n_molecules and angles.sum() == 2 * n_chains * n_molecules
    assert np.abs(lengths.astype(np.int64) - 2 * expected_lengths).sum() <= 4
    assert np.abs(angles.astype(np.int64) - 2 * expected_angles).sum() <= 4
