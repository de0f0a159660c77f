"""Whole-chain parity of the CPU oracle with the running reference.

tests/golden/trace_*.npz hold, event by event, what the unmodified reference did when its `random` module read
the slot-keyed Philox stream (tests/golden/ref_recorder.py). The oracle, started from the same configuration
and reading the same stream, must reproduce every event: winner kind, target, acceptance, new active particle
bit-exact -- and, because it follows the reference operation by operation with the same libm, also the event
times and positions bit-exact. Snapshots pin positions, cell occupancy and surplus lists along the way."""
import numpy as np
import pytest

import trace_util as tu


def _init_tables(oracle, g):
    """Init-time tables rebuilt by the oracle's restatement of estimator + Walker."""
    handler, pot, bound, veto, use_charge = tu.potentials_of(g)
    cps = [int(c) for c in g["meta_cells_per_side"]]
    prefactor, points = float(g["meta_estimator"][0]), int(g["meta_estimator"][1])
    bounds, far = oracle.inner_point_derivative_bounds(veto, float(g["meta_system_length"]), cps, 1,
                                                       prefactor=prefactor, points_per_side=points,
                                                       target_charge=1.0 if use_charge else None,
                                                       uses_charges=use_charge)
    return oracle.veto_tables(bounds, far)


@pytest.mark.parametrize("name", tu.TRACES)
def test_init_tables_bit_exact(oracle, name):
    g = tu.load_trace(name)
    ours, ref = _init_tables(oracle, g), tu.reference_tables(g)
    assert np.array_equal(ours["bounds"], ref["bounds"], equal_nan=True)
    for kind in ("upper", "lower"):
        for d in range(3):
            for key in ("cell_a", "cell_b", "rate_a"):
                assert np.array_equal(ours[kind][d][key], ref[kind][d][key]), (kind, d, key)
            assert ours[kind][d]["total_rate"] == ref[kind][d]["total_rate"]
            assert ours[kind][d]["mean_rate"] == ref[kind][d]["mean_rate"]


@pytest.mark.parametrize("name", tu.TRACES)
def test_chain_replay_bit_exact(oracle, name):
    g = tu.load_trace(name)
    records = g["records"]
    chain = oracle.OracleChain(tu.builder_of(g, oracle.ProgramBuilder, tables=_init_tables(oracle, g)))
    chain.set_positions(g["positions0"], tu.charges_of(g))
    chain.start(stream=int(g["seed"][1]))
    done = 0
    snap_events = list(g["snap_event"])
    for k, event in enumerate(snap_events + [len(records)]):
        n, rec = chain.run(max_events=int(event) - done, record=int(event) - done)
        assert n == event - done
        ref = records[done:event]
        assert tu.records_equal_discrete(rec, ref)
        assert np.array_equal(rec["time_q"], ref["time_q"]) and np.array_equal(rec["time_r"], ref["time_r"])
        assert np.array_equal(rec["active_pos"], ref["active_pos"])
        done = int(event)
        if k < len(snap_events):
            assert np.array_equal(chain.positions(), g["snap_positions"][k])
            occ, surplus = chain.cells()
            assert np.array_equal(occ, g["snap_occupants"][k])
            ns = int(g["snap_n_surplus"][k])
            assert sorted(surplus.tolist()) == sorted(g["snap_surplus"][k][:ns].tolist())
            st = chain.state()
            assert (st.active, st.direction) == (int(g["snap_active"][k]), int(g["snap_direction"][k]))
            assert (st.time_q, st.time_r) == tuple(g["snap_time"][k])
    assert np.array_equal(chain.positions(), g["final_positions"])
    assert chain.stats()["capacity_errors"] == 0


@pytest.mark.parametrize("name", tu.CELL_BOUNDING_TRACES)
def test_cell_bounding_chain_replay_bit_exact(oracle, name):
    """Far field through TwoLeafUnitCellBoundingPotentialEventHandler (coulomb_atoms/cell_bounded.ini shape)."""
    g = tu.load_trace(name)
    records = g["records"]
    # the bounds are those of the inner point estimator, rebuilt by the oracle's restatement
    _, _, _, veto, use_charge = tu.potentials_of(g)
    cps = [int(c) for c in g["meta_cells_per_side"]]
    bounds, _ = oracle.inner_point_derivative_bounds(veto, float(g["meta_system_length"]), cps, 1,
                                                     prefactor=float(g["meta_estimator"][0]),
                                                     points_per_side=int(g["meta_estimator"][1]), target_charge=1.0,
                                                     uses_charges=use_charge)
    assert np.array_equal(bounds, g["bounds"], equal_nan=True)
    chain = oracle.OracleChain(tu.builder_of(g, oracle.ProgramBuilder, tables={"bounds": bounds}))
    chain.set_positions(g["positions0"], tu.charges_of(g))
    chain.start(stream=int(g["seed"][1]))
    n, rec = chain.run(max_events=len(records), record=len(records))
    assert n == len(records)
    assert (rec["kind"] == 5).sum() > 1000
    assert tu.records_equal_discrete(rec, records)
    assert np.array_equal(rec["time_q"], records["time_q"]) and np.array_equal(rec["time_r"], records["time_r"])
    assert np.array_equal(rec["active_pos"], records["active_pos"])
    assert np.array_equal(chain.positions(), g["final_positions"])


@pytest.mark.parametrize("name", tu.NO_CELL_TRACES)
def test_no_cells_chain_replay_bit_exact(oracle, name):
    """No cell system (shipped coulomb_atoms/power_bounded.ini, six atoms): the pair factors of the factor type map make
    every other atom a candidate of every event (n_candidates = finite pair candidates + end of chain), there are no
    cell-boundary events."""
    g = tu.load_trace(name)
    records = g["records"]
    chain = oracle.OracleChain(tu.no_cells_builder_of(g, oracle.ProgramBuilder))
    chain.set_positions(g["positions0"], tu.charges_of(g))
    chain.start(stream=int(g["seed"][1]))
    done = 0
    snap_events = list(g["snap_event"])
    for k, event in enumerate(snap_events + [len(records)]):
        n, rec = chain.run(max_events=int(event) - done, record=int(event) - done)
        assert n == event - done
        ref = records[done:event]
        for f in tu.DISCRETE_FIELDS:
            assert np.array_equal(rec[f], ref[f]), (f, done + int(np.nonzero(rec[f] != ref[f])[0][0]))
        assert np.array_equal(rec["time_q"], ref["time_q"]) and np.array_equal(rec["time_r"], ref["time_r"])
        assert np.array_equal(rec["active_pos"], ref["active_pos"])
        done = int(event)
        if k < len(snap_events):
            assert np.array_equal(chain.positions(), g["snap_positions"][k])
            st = chain.state()
            assert (st.active, st.direction) == (int(g["snap_active"][k]), int(g["snap_direction"][k]))
            assert (st.time_q, st.time_r) == tuple(g["snap_time"][k])
    assert (records["kind"] == 3).sum() == 0 and (records["kind"] == 1).sum() > 3000
    assert np.array_equal(chain.positions(), g["final_positions"])
    assert chain.stats()["capacity_errors"] == 0


@pytest.mark.parametrize("name", tu.COMPOSITE_CELL_BOUNDING_TRACES)
def test_composite_cell_bounding_chain_replay_bit_exact(oracle, name):
    """The shipped dipoles/cell_bounded.ini sized for four dipoles: TwoCompositeObjectCellBoundingPotentialEventHandler for
    the objects in cells that are not nearby (two thirds of all events), composite-object Coulomb handlers for the rest,
    root-level 3 x 5 x 7 cells with cell-boundary events of the root."""
    g = tu.load_trace(name)
    records = g["records"]
    chain = oracle.OracleChain(tu.dipole_cell_bounded_builder_of(g, oracle.ProgramBuilder))
    chain.set_positions(g["positions0"], tu.charges_of(g))
    chain.set_roots(g["roots0"])
    chain.start(stream=int(g["seed"][1]))
    done = 0
    snap_events = list(g["snap_event"])
    for k, event in enumerate(snap_events + [len(records)]):
        n, rec = chain.run(max_events=int(event) - done, record=int(event) - done)
        assert n == event - done
        ref = records[done:event]
        for f in tu.DISCRETE_FIELDS:
            assert np.array_equal(rec[f], ref[f]), (f, done + int(np.nonzero(rec[f] != ref[f])[0][0]))
        assert np.array_equal(rec["time_q"], ref["time_q"]) and np.array_equal(rec["time_r"], ref["time_r"])
        assert np.array_equal(rec["active_pos"], ref["active_pos"])
        done = int(event)
        if k < len(snap_events):
            assert np.array_equal(chain.positions(), g["snap_positions"][k])
            assert np.array_equal(chain.roots(), g["snap_roots"][k])
            occ, surplus = chain.cells()
            assert np.array_equal(occ, g["snap_occupants"][k])
            ns = int(g["snap_n_surplus"][k])
            assert sorted(surplus.tolist()) == sorted(g["snap_surplus"][k][:ns].tolist())
    assert (records["kind"] == 5).sum() > 2000 and (records["kind"] == 3).sum() > 100
    assert np.array_equal(chain.positions(), g["final_positions"]) and np.array_equal(chain.roots(), g["final_roots"])


@pytest.mark.parametrize("name", sorted(tu.NO_CELL_MOLECULE_TRACES))
def test_no_cells_composite_chain_replay_bit_exact(oracle, name):
    """Composite point objects without a cell system: the three shipped dipoles/dipole_factors_*.ini (three dipoles;
    composite-object Coulomb factor from the factor type map with inside-first / outside-first / ratio lifting, harmonic
    bond, 1/r^6 repulsion between objects) and the shipped water/single_molecule.ini (bonds and bending only)."""
    g = tu.load_trace(name)
    records = g["records"]
    chain = oracle.OracleChain(tu.NO_CELL_MOLECULE_TRACES[name](g, oracle.ProgramBuilder))
    chain.set_positions(g["positions0"], tu.charges_of(g))
    chain.set_roots(g["roots0"])
    chain.start(stream=int(g["seed"][1]))
    done = 0
    snap_events = list(g["snap_event"])
    for k, event in enumerate(snap_events + [len(records)]):
        n, rec = chain.run(max_events=int(event) - done, record=int(event) - done)
        assert n == event - done
        ref = records[done:event]
        for f in tu.DISCRETE_FIELDS:
            assert np.array_equal(rec[f], ref[f]), (f, done + int(np.nonzero(rec[f] != ref[f])[0][0]))
        assert np.array_equal(rec["time_q"], ref["time_q"]) and np.array_equal(rec["time_r"], ref["time_r"])
        assert np.array_equal(rec["active_pos"], ref["active_pos"])
        done = int(event)
        if k < len(snap_events):
            assert np.array_equal(chain.positions(), g["snap_positions"][k])
            assert np.array_equal(chain.roots(), g["snap_roots"][k])
    assert (records["kind"] == 3).sum() == 0
    assert np.array_equal(chain.positions(), g["final_positions"]) and np.array_equal(chain.roots(), g["final_roots"])
    assert chain.stats()["capacity_errors"] == 0


@pytest.mark.parametrize("name", tu.LEAF_CELL_WATER_TRACES)
def test_leaf_cell_water_chain_replay_bit_exact(oracle, name):
    """The shipped water/coulomb_power_bounded_lj_cell_bounded.ini (twelve molecules in its 6^3 cells, sixteen in 4^3 cells,
    forty in 3^3 cells with a permanently filled surplus list):
    a cell system that stores the oxygens only; the Lennard-Jones factor between oxygens through
    TwoLeafUnitEventHandlerWithPiecewiseConstantBoundingPotential (nearby cells) and
    TwoLeafUnitCellBoundingPotentialEventHandler (all other cells), cell-boundary events of the active oxygen that leave the
    composite-object Coulomb factors, the bonds and the bending factor running."""
    g = tu.load_trace(name)
    records = g["records"]
    chain = oracle.OracleChain(tu.leaf_cell_water_builder_of(g, oracle.ProgramBuilder))
    chain.set_positions(g["positions0"], tu.charges_of(g))
    chain.set_roots(g["roots0"])
    chain.start(stream=int(g["seed"][1]))
    done = 0
    snap_events = list(g["snap_event"])
    for k, event in enumerate(snap_events + [len(records)]):
        n, rec = chain.run(max_events=int(event) - done, record=int(event) - done)
        assert n == event - done
        ref = records[done:event]
        for f in tu.DISCRETE_FIELDS:
            assert np.array_equal(rec[f], ref[f]), (f, done + int(np.nonzero(rec[f] != ref[f])[0][0]))
        assert np.array_equal(rec["time_q"], ref["time_q"]) and np.array_equal(rec["time_r"], ref["time_r"])
        assert np.array_equal(rec["active_pos"], ref["active_pos"])
        done = int(event)
        if k < len(snap_events):
            assert np.array_equal(chain.positions(), g["snap_positions"][k])
            assert np.array_equal(chain.roots(), g["snap_roots"][k])
            occ, surplus = chain.cells()
            assert np.array_equal(occ, g["snap_occupants"][k])
            ns = int(g["snap_n_surplus"][k])
            assert sorted(surplus.tolist()) == sorted(g["snap_surplus"][k][:ns].tolist())
    kinds = np.bincount(records["kind"], minlength=9)
    if name.endswith("_surplus"):  # 3^3 cells: every cell is nearby, 17-18 of the forty oxygens in the surplus
        assert kinds[5] == 0 and kinds[7] > 900 and kinds[3] >= 2 and int(g["snap_n_surplus"].min()) >= 15
    else:
        assert kinds[5] > 1000 and kinds[7] > 40 and kinds[3] >= 3 and kinds[1] > 1500
    assert np.array_equal(chain.positions(), g["final_positions"]) and np.array_equal(chain.roots(), g["final_roots"])
    assert chain.stats()["capacity_errors"] == 0


def test_root_unit_active_mode_replay_bit_exact(oracle):
    """The shipped dipoles/dipole_motion.ini (three dipoles): which unit is active after every event -- the root unit of
    an object or one of its leaves (RootLeafUnitActiveSwitcher) -- and, with the shipped sampling events in between,
    the candidates of the root-unit-active handlers that survive a host control event (their out-states time-slice a
    fresh copy of the objects, the confirmation uses the leaf units of the in-state)."""
    g = tu.load_trace("trace_dipole_motion")
    chain = oracle.OracleChain(tu.dipole_motion_builder_of(g, oracle.ProgramBuilder))
    chain.set_positions(g["positions0"], tu.charges_of(g))
    chain.set_roots(g["roots0"])
    chain.start(stream=int(g["seed"][1]))
    n, rec = chain.run(max_events=len(g["records"]), record=len(g["records"]))
    assert np.array_equal(rec["mode"], g["records"]["reserved"])
    assert (g["records"]["kind"] == 9).sum() > 100 and 0.3 < g["records"]["reserved"].mean() < 0.7
    root = g["records"]["reserved"][:-1] == 1  # events that start with a root unit active
    assert ((g["records"]["kind"][1:] == 1) & root).sum() > 500 and ((g["records"]["kind"][1:] == 7) & root).sum() > 50

    g = tu.load_trace("trace_dipole_motion_sampling")
    records, host = g["records"], g["host_times"]
    chain = oracle.OracleChain(tu.dipole_motion_builder_of(g, oracle.ProgramBuilder))
    chain.set_positions(g["positions0"], tu.charges_of(g))
    chain.set_roots(g["roots0"])
    chain.start(stream=int(g["seed"][1]))
    parts, done = [], 0
    assert len(host) > 100
    for events_before, q, r in host:
        n, rec = chain.run(until=(q, r), record=10000)
        parts.append(rec)
        done += n
        assert done == int(events_before)
    n, rec = chain.run(max_events=len(records) - done, record=10000)
    parts.append(rec)
    ours = np.concatenate(parts)
    assert tu.records_equal_discrete(ours, records) and np.array_equal(ours["mode"], records["reserved"])
    assert np.array_equal(ours["time_q"], records["time_q"]) and np.array_equal(ours["time_r"], records["time_r"])
    assert np.array_equal(ours["active_pos"], records["active_pos"])
    assert np.array_equal(chain.positions(), g["final_positions"]) and np.array_equal(chain.roots(), g["final_roots"])


@pytest.mark.parametrize("name", tu.DIPOLE_TRACES)
def test_composite_chain_replay_bit_exact(oracle, name):
    """C1, the shipped hard_disk_dipoles_cells.ini from the shipped start configuration: composite point objects
    (root units follow their active leaf), leaf-level cells with several occupants, hard-sphere pair events,
    hard-dipole tether events of the factor type map, end of chain drawing (root, child)."""
    g = tu.load_trace(name)
    records = g["records"]
    chain = oracle.OracleChain(tu.dipole_builder_of(g, oracle.ProgramBuilder))
    chain.set_positions(g["positions0"])
    chain.set_roots(g["roots0"])
    chain.start(stream=int(g["seed"][1]))
    done = 0
    snap_events = list(g["snap_event"])
    for k, event in enumerate(snap_events + [len(records)]):
        n, rec = chain.run(max_events=int(event) - done, record=int(event) - done)
        assert n == event - done
        ref = records[done:event]
        for f in tu.DISCRETE_FIELDS:
            assert np.array_equal(rec[f], ref[f]), (f, done + int(np.nonzero(rec[f] != ref[f])[0][0]))
        assert np.array_equal(rec["time_q"], ref["time_q"]) and np.array_equal(rec["time_r"], ref["time_r"])
        assert np.array_equal(rec["active_pos"], ref["active_pos"])
        done = int(event)
        if k < len(snap_events):
            assert np.array_equal(chain.positions(), g["snap_positions"][k])
            assert np.array_equal(chain.roots(), g["snap_roots"][k])
            occ, surplus = chain.cells()
            assert np.array_equal(occ, g["snap_occupants"][k]) and len(surplus) == 0
    assert np.array_equal(chain.positions(), g["final_positions"])
    assert np.array_equal(chain.roots(), g["final_roots"])
    stats = chain.stats()
    assert stats["capacity_errors"] == 0 and stats["bond_events"] > 500 and stats["pair_events"] > 4000


def test_sequential_direction_chain_replay_bit_exact(oracle):
    """The shipped hard_disk_dipoles.ini (no cell system: 160 hard-disk candidates and the tether per event) with GENERAL
    velocities: SingleIndependentActiveSequentialDirectionEndOfChainEventHandler rotates the velocity by 20 degrees at
    every end of chain (:101-122). 5000 events of the running reference from the shipped PDB start configuration:
    every event bit for bit, and at the snapshots the leaf and root positions and the velocities of the active leaf and
    of its root unit -- the root's as the reference accumulates it from velocity changes (abstracts.py:165-227)."""
    g = tu.load_trace("trace_hard_disk_dipoles_sequential")
    records = g["records"]
    chain = oracle.OracleChain(tu.sequential_dipole_builder_of(g, oracle.ProgramBuilder))
    chain.set_positions(g["positions0"])
    chain.set_roots(g["roots0"])
    chain.start(stream=int(g["seed"][1]))
    done = 0
    snap_events = list(g["snap_event"])
    for k, event in enumerate(snap_events + [len(records)]):
        n, rec = chain.run(max_events=int(event) - done, record=int(event) - done)
        assert n == event - done
        ref = records[done:event]
        for f in tu.DISCRETE_FIELDS:
            assert np.array_equal(rec[f], ref[f]), (f, done + int(np.nonzero(rec[f] != ref[f])[0][0]))
        assert np.array_equal(rec["time_q"], ref["time_q"]) and np.array_equal(rec["time_r"], ref["time_r"])
        assert np.array_equal(rec["active_pos"], ref["active_pos"])
        done = int(event)
        if k < len(snap_events):
            assert np.array_equal(chain.positions(), g["snap_positions"][k])
            assert np.array_equal(chain.roots(), g["snap_roots"][k])
            state = chain.state()
            assert state.active == int(g["snap_active"][k])
            assert [state.velocity[0], state.velocity[1]] == g["snap_velocities"][k][0, :2].tolist()
            assert [state.root_velocity[0], state.root_velocity[1]] == g["snap_velocities"][k][1, :2].tolist()
    assert np.array_equal(chain.positions(), g["final_positions"])
    assert np.array_equal(chain.roots(), g["final_roots"])
    stats = chain.stats()
    assert stats["end_of_chain_events"] > 150 and stats["bond_events"] > 1500 and stats["factor_pair_events"] > 2000
    # the velocity really is general: both components non-zero, and its norm drifts from 1 only by rounding
    velocities = g["snap_velocities"][:, 0, :2]
    assert np.all(velocities[1:] != 0.0) and np.max(np.abs(np.linalg.norm(velocities, axis=1) - 1.0)) < 1e-13


@pytest.mark.parametrize("name", tu.WATER_TRACES)
def test_water_chain_replay_bit_exact(oracle, name):
    """C4, the shipped water/coulomb_cell_veto_lj_inverted.ini: composite-object Coulomb pair and cell-veto events with
    inside-first lifting over the six leaf units, Lennard-Jones between the oxygens, harmonic bonds, the bending
    factor with its piecewise constant bounding potential and ratio lifting, root units in root-level cells."""
    g = tu.load_trace(name)
    records = g["records"]
    chain = oracle.OracleChain(tu.water_builder_of(g, oracle.ProgramBuilder))
    chain.set_positions(g["positions0"], g["charges"])
    chain.set_roots(g["roots0"])
    chain.start(stream=int(g["seed"][1]))
    done = 0
    snap_events = list(g["snap_event"])
    for k, event in enumerate(snap_events + [len(records)]):
        n, rec = chain.run(max_events=int(event) - done, record=int(event) - done)
        assert n == event - done
        ref = records[done:event]
        for f in tu.DISCRETE_FIELDS:
            assert np.array_equal(rec[f], ref[f]), (f, done + int(np.nonzero(rec[f] != ref[f])[0][0]))
        assert np.array_equal(rec["time_q"], ref["time_q"]) and np.array_equal(rec["time_r"], ref["time_r"])
        assert np.array_equal(rec["active_pos"], ref["active_pos"])
        done = int(event)
        if k < len(snap_events):
            assert np.array_equal(chain.positions(), g["snap_positions"][k])
            assert np.array_equal(chain.roots(), g["snap_roots"][k])
            occ, surplus = chain.cells()
            assert np.array_equal(occ, g["snap_occupants"][k])
            ns = int(g["snap_n_surplus"][k])
            assert sorted(surplus.tolist()) == sorted(g["snap_surplus"][k][:ns].tolist())
    assert np.array_equal(chain.positions(), g["final_positions"])
    assert np.array_equal(chain.roots(), g["final_roots"])
    assert chain.stats()["capacity_errors"] == 0


def test_time_limit_keeps_candidates(oracle):
    """Stopping at host control times (sampling) keeps the interaction winner: the event sequence is the same
    whether the chain runs in one go or is interrupted, up to the rounding of the extra time slices."""
    g = tu.load_trace("trace_lj_small")
    pb = tu.builder_of(g, oracle.ProgramBuilder)
    free = oracle.OracleChain(pb)
    free.set_positions(g["positions0"])
    free.start(stream=3)
    n_free, rec_free = free.run(until=(2.0, 0.5), record=20000)
    stepped = oracle.OracleChain(pb)
    stepped.set_positions(g["positions0"])
    stepped.start(stream=3)
    parts = []
    for k in range(1, 26):
        t = oracle.time_from_float(0.1 * k)
        _, rec = stepped.run(until=t, record=20000)
        parts.append(rec)
        st = stepped.state()
        assert (st.time_q, st.time_r) == t
    rec_stepped = np.concatenate(parts)
    assert len(rec_stepped) == n_free
    for f in ("kind", "target", "accepted", "new_active"):
        assert np.array_equal(rec_stepped[f], rec_free[f]), f
    assert tu.max_time_error(rec_stepped, rec_free) < 1e-12


def test_sampling_interleaved_replay_bit_exact(oracle):
    """The reference trace with FixedIntervalSamplingEventHandler events in between: running the oracle up to
    each recorded sampling time reproduces the reference bit for bit (time slices at the sampling times, kept
    candidates, no extra draws)."""
    g = tu.load_trace("trace_lj_sampling")
    records, host = g["records"], g["host_times"]
    chain = oracle.OracleChain(tu.builder_of(g, oracle.ProgramBuilder))
    chain.set_positions(g["positions0"])
    chain.start(stream=int(g["seed"][1]))
    parts, done = [], 0
    assert len(host) > 20
    for events_before, q, r in host:
        n, rec = chain.run(until=(q, r), record=10000)
        parts.append(rec)
        done += n
        assert done == int(events_before)
    n, rec = chain.run(max_events=len(records) - done, record=10000)
    parts.append(rec)
    ours = np.concatenate(parts)
    assert tu.records_equal_discrete(ours, records)
    assert np.array_equal(ours["time_q"], records["time_q"]) and np.array_equal(ours["time_r"], records["time_r"])
    assert np.array_equal(ours["active_pos"], records["active_pos"])
