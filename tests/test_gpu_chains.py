"""GPU parity of the event kernel (candidate times -> argmin -> lifting -> commit) through the C ABI.

* against the traces recorded from the running reference (tests/golden/trace_*.npz): every event's winner kind,
  target, acceptance and new active particle bit-exact, event times and positions within 1e-12;
* against the CPU oracle on seeded batches of chains (each chain its own start configuration and random stream);
* host control events (time limits): kept candidates, time slices at the sampling times;
* edge cases: one particle, empty veto cells, several occupants per cell, 2D hard disks, surplus overflow."""
import numpy as np
import pytest

import trace_util as tu
from jellyfysh_b200 import abi, engine
from jellyfysh_b200.program import ProgramBuilder

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def assert_records_match(ours, ref, length, tag=""):
    assert len(ours) == len(ref), (tag, len(ours), len(ref))
    differs = np.zeros(len(ref), dtype=bool)
    for f in tu.DISCRETE_FIELDS:
        differs |= ours[f] != ref[f]
    if differs.any():
        k = int(np.nonzero(differs)[0][0])
        raise AssertionError(f"{tag}: first difference at event {k}: ours {ours[k]} reference {ref[k]}")
    assert tu.max_time_error(ours, ref) < RTOL, (tag, tu.max_time_error(ours, ref))
    assert np.max(np.abs(ours["active_pos"] - ref["active_pos"])) < RTOL * max(1.0, length), tag


@pytest.mark.parametrize("name", tu.TRACES + tu.CELL_BOUNDING_TRACES)
def test_reference_trace_replay(name):
    g = tu.load_trace(name)
    records = g["records"]
    length = float(g["meta_system_length"])
    with engine.Engine(tu.builder_of(g, ProgramBuilder), n_chains=1) as eng:
        eng.upload_positions(g["positions0"][None], None if tu.charges_of(g) is None else tu.charges_of(g)[None])
        eng.start(first_stream=int(g["seed"][1]))
        done = 0
        snaps = list(g["snap_event"])
        for k, event in enumerate(snaps + [len(records)]):
            count = int(event) - done
            rec, stats = eng.run_recorded(max_events=count, records_per_chain=count)
            assert stats["events"] == count and stats["capacity_errors"] == 0
            assert_records_match(rec[0], records[done:event], length, f"{name}[{done}:{event}]")
            done = int(event)
            if k < len(snaps):
                assert np.max(np.abs(eng.download_positions()[0] - g["snap_positions"][k])) < RTOL * max(1.0, length)
                occ, surplus = eng.cells()
                assert np.array_equal(occ[0], g["snap_occupants"][k])
                ns = int(g["snap_n_surplus"][k])
                assert sorted(surplus[0].tolist()) == sorted(g["snap_surplus"][k][:ns].tolist())
                st = eng.chain_states()[0]
                assert (int(st["active"]), int(st["direction"])) == (int(g["snap_active"][k]), int(g["snap_direction"][k]))
                assert abs((st["time_q"] - g["snap_time"][k][0]) + (st["time_r"] - g["snap_time"][k][1])) < RTOL * max(1.0, st["time_q"])
        assert np.max(np.abs(eng.download_positions()[0] - g["final_positions"])) < RTOL * max(1.0, length)
        assert eng.kernel_launches == len(snaps) + 1


def test_sampling_interleaved_trace():
    """Reference run with FixedIntervalSamplingEventHandler events: ecmc_run(until = sampling time) per sample."""
    g = tu.load_trace("trace_lj_sampling")
    records, host = g["records"], g["host_times"]
    with engine.Engine(tu.builder_of(g, ProgramBuilder), n_chains=1) as eng:
        eng.upload_positions(g["positions0"][None])
        eng.start(first_stream=int(g["seed"][1]))
        parts, done = [], 0
        for events_before, q, r in host:
            rec, stats = eng.run_recorded(until=(q, r), records_per_chain=200)
            parts.append(rec[0][:stats["events"]])
            done += stats["events"]
            assert done == int(events_before)
            st = eng.chain_states()[0]
            assert (st["time_q"], st["time_r"]) == (q, r)
        rec, stats = eng.run_recorded(max_events=len(records) - done, records_per_chain=len(records) - done)
        parts.append(rec[0])
        assert_records_match(np.concatenate(parts), records, float(g["meta_system_length"]), "sampling")


def _lj_batch(oracle, n_chains, n=48, cells=4, length=4.6, seed=5, chain_time=1.3, max_occupants=1):
    pot = abi.EcmcPotential.make(abi.POT_LENNARD_JONES, 4.0, 1.0)
    bounds, far = oracle.inner_point_derivative_bounds(pot, length, [cells] * 3, 1, prefactor=1.5, points_per_side=3)
    tables = oracle.veto_tables(bounds, far)
    pb = ProgramBuilder(3, n, length, 1.0, [cells] * 3, 1, max_occupants=max_occupants, max_surplus=n,
                        chain_time=chain_time, seed=seed)
    pb.set_pair(abi.PAIR_TWO_LEAF_UNIT, pot)
    pb.set_veto(pot, tables)
    rng = np.random.default_rng(100 + seed)
    # a jittered lattice keeps all pair distances away from the hard core
    side = int(np.ceil(n ** (1 / 3)))
    grid = np.stack(np.meshgrid(*[np.arange(side)] * 3, indexing="ij"), axis=-1).reshape(-1, 3)[:n]
    positions = np.empty((n_chains, n, 3))
    for c in range(n_chains):
        positions[c] = ((grid + 0.5) * (length / side) + rng.uniform(-0.12, 0.12, size=(n, 3))) % length
    return pb, positions


def _compare_batch_with_oracle(oracle, pb, positions, charges, n_events, first_stream, tag, resync_every=None,
                               roots=None):
    """Run the same seeded chains on the GPU and in the oracle and compare every event and the final state.

    Event chains are chaotic (hard-core collisions amplify a rounding difference of 1e-16 by a constant factor per
    event), so for the configurations where that matters the GPU state is re-seeded from the oracle every
    `resync_every` events through ecmc_upload_positions / _chain_states / _cells: each segment then tests the
    kernel on identical inputs, which is what "event times within 1e-12" can mean for a chaotic system."""
    n_chains = len(positions)
    length = float(pb.program.system_length)
    chains = []
    for c in range(n_chains):
        chain = oracle.OracleChain(pb)
        chain.set_positions(positions[c], None if charges is None else charges[c])
        if roots is not None:
            chain.set_roots(roots[c])
        chain.start(stream=first_stream + c)
        chains.append(chain)
    total = dict.fromkeys(["events", "pair_events", "veto_events", "veto_accepted", "boundary_events",
                           "end_of_chain_events", "candidates", "bond_events", "factor_pair_events"], 0)
    segment = resync_every or n_events
    with engine.Engine(pb, n_chains=n_chains) as eng:
        eng.upload_positions(positions, charges)
        if roots is not None:
            eng.upload_roots(roots)
        eng.start(first_stream=first_stream)
        for begin in range(0, n_events, segment):
            count = min(segment, n_events - begin)
            rec, stats = eng.run_recorded(max_events=count, records_per_chain=count)
            final = eng.download_positions()
            final_roots = eng.download_roots() if roots is not None else None
            occ, surplus = eng.cells()
            states = eng.chain_states()
            for c, chain in enumerate(chains):
                n, ref = chain.run(max_events=count, record=count)
                assert n == count
                assert_records_match(rec[c], ref, length, f"{tag} chain {c} events {begin}+")
                assert np.max(np.abs(final[c] - chain.positions())) < RTOL * max(1.0, length)
                if roots is not None:
                    assert np.max(np.abs(final_roots[c] - chain.roots())) < RTOL * max(1.0, length)
                o_occ, o_sur = chain.cells()
                assert np.array_equal(occ[c], o_occ)
                assert surplus[c].tolist() == o_sur.tolist()
                st = chain.state()
                assert (int(states[c]["active"]), int(states[c]["direction"]), int(states[c]["active_cell"]),
                        int(states[c]["event_counter"]), int(states[c]["eoc_next_active"])) == \
                       (st.active, st.direction, st.active_cell, st.event_counter, st.eoc_next_active)
            for key in total:
                total[key] += stats[key]
            if resync_every:
                eng.upload_positions(np.stack([chain.positions() for chain in chains]), charges)
                if roots is not None:
                    eng.upload_roots(np.stack([chain.roots() for chain in chains]))
                new_states = states.copy()
                for c, chain in enumerate(chains):
                    st = chain.state()
                    for name in new_states.dtype.names:
                        new_states[c][name] = getattr(st, name)
                eng.set_chain_states(new_states)
                eng.set_cells(np.stack([chain.cells()[0] for chain in chains]), [chain.cells()[1] for chain in chains])
    oracle_total = dict.fromkeys(total, 0)
    for chain in chains:
        for key, value in chain.stats().items():
            if key in oracle_total:
                oracle_total[key] += value
    assert total == oracle_total
    return total


def test_lennard_jones_batch_against_oracle(oracle):
    pb, positions = _lj_batch(oracle, n_chains=37)
    stats = _compare_batch_with_oracle(oracle, pb, positions, None, 1500, 1000, "lj")
    assert stats["pair_events"] > 1000 and stats["veto_accepted"] > 0 and stats["end_of_chain_events"] > 0


def test_lennard_jones_2d_against_oracle(oracle):
    """Two-dimensional Lennard-Jones with a cell veto: the Lennard-Jones kernels that hard-wire three dimensions must not
    be picked (pick_kernel), the end of chain rotates the direction through two axes only."""
    pot = abi.EcmcPotential.make(abi.POT_LENNARD_JONES, 4.0, 1.0)
    cells, length, n, n_chains = 5, 6.5, 30, 11
    bounds, far = oracle.inner_point_derivative_bounds(pot, length, [cells] * 2, 1, prefactor=4.0, points_per_side=3)
    pb = ProgramBuilder(2, n, length, 1.0, [cells] * 2, 1, max_occupants=1, max_surplus=n, chain_time=1.3, seed=5)
    pb.set_pair(abi.PAIR_TWO_LEAF_UNIT, pot)
    pb.set_veto(pot, oracle.veto_tables(bounds, far))
    rng = np.random.default_rng(7)
    side = int(np.ceil(n ** 0.5))
    grid = np.stack(np.meshgrid(*[np.arange(side)] * 2, indexing="ij"), axis=-1).reshape(-1, 2)[:n]
    positions = np.stack([((grid + 0.5) * (length / side) + rng.uniform(-0.12, 0.12, size=(n, 2))) % length
                          for _ in range(n_chains)])
    with engine.Engine(pb, n_chains=n_chains) as eng:
        assert "event_kernel" in eng.kernel_name()  # not the batched kernel, which is three-dimensional
    stats = _compare_batch_with_oracle(oracle, pb, positions, None, 1200, 300, "lj 2d", resync_every=300)
    assert stats["pair_events"] > 1000 and stats["veto_events"] > 1000 and stats["end_of_chain_events"] > 100
    assert stats["boundary_events"] > 100


def test_several_occupants_per_cell_against_oracle(oracle):
    pb, positions = _lj_batch(oracle, n_chains=9, n=100, cells=4, length=5.2, seed=9, max_occupants=2)
    _compare_batch_with_oracle(oracle, pb, positions, None, 800, 50, "lj m=3")


def test_coulomb_batch_against_oracle(oracle):
    n, cells, length = 24, 4, 1.0
    mic = abi.EcmcPotential.make(abi.POT_MERGED_IMAGE_COULOMB, 1.0, 3.45, 6, 2)
    ipcb = abi.EcmcPotential.make(abi.POT_INVERSE_POWER_COULOMB_BOUNDING, 1.5837)
    bounds, far = oracle.inner_point_derivative_bounds(mic, length, [cells] * 3, 1, prefactor=1.0, points_per_side=3,
                                                       target_charge=1.0, uses_charges=True)
    tables = oracle.veto_tables(bounds, far)
    pb = ProgramBuilder(3, n, length, 2.0, [cells] * 3, 1, max_occupants=1, max_surplus=n, chain_time=0.78965, seed=21)
    pb.set_pair(abi.PAIR_TWO_LEAF_UNIT_BOUNDING, mic, ipcb, use_charge=True)
    pb.set_veto(mic, tables, use_charge=True, target_charge=1.0)
    rng = np.random.default_rng(77)
    n_chains = 11
    positions = rng.uniform(0.0, length, size=(n_chains, n, 3))
    charges = np.where(rng.random((n_chains, n)) < 0.5, 1.0, -1.0)
    charges[0] = 1.0
    stats = _compare_batch_with_oracle(oracle, pb, positions, charges, 1200, 7, "coulomb")
    assert stats["pair_events"] > 100 and stats["veto_events"] > 1000


@pytest.mark.parametrize("name", tu.NO_CELL_TRACES)
def test_no_cells_reference_trace_replay(oracle, name):
    """No cell system (shipped coulomb_atoms/power_bounded.ini shape): every other atom is a candidate of every event,
    no cell-boundary events; every event of the reference trace on the device.

    Every event of this configuration is a pair event of a few strongly coupled atoms, and half of them lift: the chain
    is chaotic and amplifies the last-bit differences between the device arithmetic and libm by about 2 % per event
    (measured on B200: 1.2e-12 after 500 events). The trace is therefore replayed in stretches of 100 events; at the
    start of each stretch the device takes the state of the oracle, which reproduces the reference bit for bit
    (tests/test_oracle_traces.py::test_no_cells_chain_replay_bit_exact)."""
    g = tu.load_trace(name)
    records = g["records"]
    stretch = 100
    chain = oracle.OracleChain(tu.no_cells_builder_of(g, oracle.ProgramBuilder))
    chain.set_positions(g["positions0"], tu.charges_of(g))
    chain.start(stream=int(g["seed"][1]))
    with engine.Engine(tu.no_cells_builder_of(g, ProgramBuilder), n_chains=1) as eng:
        eng.upload_positions(g["positions0"][None], tu.charges_of(g)[None])
        eng.start(first_stream=int(g["seed"][1]))
        for done in range(0, len(records), stretch):
            count = min(stretch, len(records) - done)
            if done:
                eng.upload_positions(chain.positions()[None], tu.charges_of(g)[None])
                eng.set_chain_states(np.frombuffer(bytes(chain.state()), dtype=abi.chain_state_dtype()))
                occupants, surplus = chain.cells()
                eng.set_cells(occupants[None], [surplus])
            rec, stats = eng.run_recorded(max_events=count, records_per_chain=count)
            assert stats["events"] == count and stats["capacity_errors"] == 0 and stats["boundary_events"] == 0
            assert_records_match(rec[0], records[done:done + count], 1.0, f"{name}[{done}:{done + count}]")
            n, ours = chain.run(max_events=count, record=count)
            assert n == count and tu.records_equal_discrete(ours, records[done:done + count])
            assert np.max(np.abs(eng.download_positions()[0] - chain.positions())) < RTOL


@pytest.mark.parametrize("name", sorted(tu.NO_CELL_MOLECULE_TRACES))
def test_no_cells_composite_reference_trace_replay(oracle, name):
    """Composite point objects without a cell system on the device (molecule_kernel): the three shipped
    dipoles/dipole_factors_*.ini (inside-first, outside-first and ratio lifting of the composite-object Coulomb factor)
    and the shipped water/single_molecule.ini, every event of the reference traces, in resynchronised stretches like
    test_no_cells_reference_trace_replay."""
    g = tu.load_trace(name)
    records = g["records"]
    length = float(g["meta_system_length"])
    stretch = 50
    build = tu.NO_CELL_MOLECULE_TRACES[name]
    chain = oracle.OracleChain(build(g, oracle.ProgramBuilder))
    chain.set_positions(g["positions0"], tu.charges_of(g))
    chain.set_roots(g["roots0"])
    chain.start(stream=int(g["seed"][1]))
    with engine.Engine(build(g, ProgramBuilder), n_chains=1) as eng:
        eng.upload_positions(g["positions0"][None], tu.charges_of(g)[None])
        eng.upload_roots(g["roots0"][None])
        eng.start(first_stream=int(g["seed"][1]))
        for done in range(0, len(records), stretch):
            count = min(stretch, len(records) - done)
            if done:
                eng.upload_positions(chain.positions()[None], tu.charges_of(g)[None])
                eng.upload_roots(chain.roots()[None])
                eng.set_chain_states(np.frombuffer(bytes(chain.state()), dtype=abi.chain_state_dtype()))
                occupants, surplus = chain.cells()
                eng.set_cells(occupants[None], [surplus])
            rec, stats = eng.run_recorded(max_events=count, records_per_chain=count)
            assert stats["events"] == count and stats["capacity_errors"] == 0 and stats["boundary_events"] == 0
            assert_records_match(rec[0], records[done:done + count], length, f"{name}[{done}:{done + count}]")
            n, ours = chain.run(max_events=count, record=count)
            assert n == count and tu.records_equal_discrete(ours, records[done:done + count])
            assert np.max(np.abs(eng.download_positions()[0] - chain.positions())) < RTOL * max(1.0, length)
            assert np.max(np.abs(eng.download_roots()[0] - chain.roots())) < RTOL * max(1.0, length)


@pytest.mark.parametrize("name", tu.LEAF_CELL_WATER_TRACES)
def test_leaf_cell_water_reference_trace_replay(oracle, name):
    """The shipped water/coulomb_power_bounded_lj_cell_bounded.ini on the device (molecule_kernel<..., LEAF_CELLS>): a cell
    system that stores the oxygens only, the Lennard-Jones factor between oxygens through the piecewise-constant-bound
    handler (nearby cells, surplus) and the leaf-level cell-bounding handler (all other cells), cell-boundary events of the
    active oxygen that leave the composite Coulomb factors, bonds and bending running. Every event of the reference traces
    (12 molecules in 6^3 cells, 16 in 4^3, 40 in 3^3 with 17-18 oxygens in the surplus), in stretches that start from the
    oracle's state; occupancy and surplus compared after each."""
    g = tu.load_trace(name)
    records = g["records"]
    length = float(g["meta_system_length"])
    stretch = 50
    build = tu.leaf_cell_water_builder_of
    chain = oracle.OracleChain(build(g, oracle.ProgramBuilder))
    chain.set_positions(g["positions0"], tu.charges_of(g))
    chain.set_roots(g["roots0"])
    chain.start(stream=int(g["seed"][1]))
    boundaries = 0
    with engine.Engine(build(g, ProgramBuilder), n_chains=1) as eng:
        assert "leaf cells" in eng.kernel_name()
        eng.upload_positions(g["positions0"][None], tu.charges_of(g)[None])
        eng.upload_roots(g["roots0"][None])
        eng.start(first_stream=int(g["seed"][1]))
        occupants, surplus = eng.cells()
        ref_occupants, ref_surplus = chain.cells()
        assert np.array_equal(occupants[0], ref_occupants) and sorted(surplus[0].tolist()) == sorted(ref_surplus.tolist())
        assert int(eng.chain_states()[0]["active_cell"]) == chain.state().active_cell
        for done in range(0, len(records), stretch):
            count = min(stretch, len(records) - done)
            if done:
                eng.upload_positions(chain.positions()[None], tu.charges_of(g)[None])
                eng.upload_roots(chain.roots()[None])
                eng.set_chain_states(np.frombuffer(bytes(chain.state()), dtype=abi.chain_state_dtype()))
                ref_occupants, ref_surplus = chain.cells()
                eng.set_cells(ref_occupants[None], [ref_surplus])
            rec, stats = eng.run_recorded(max_events=count, records_per_chain=count)
            assert stats["events"] == count and stats["capacity_errors"] == 0
            boundaries += stats["boundary_events"]
            assert_records_match(rec[0], records[done:done + count], length, f"{name}[{done}:{done + count}]")
            n, ours = chain.run(max_events=count, record=count)
            assert n == count and tu.records_equal_discrete(ours, records[done:done + count])
            assert np.max(np.abs(eng.download_positions()[0] - chain.positions())) < RTOL * length
            assert np.max(np.abs(eng.download_roots()[0] - chain.roots())) < RTOL * length
            occupants, surplus = eng.cells()
            ref_occupants, ref_surplus = chain.cells()
            assert np.array_equal(occupants[0], ref_occupants) and sorted(surplus[0].tolist()) == sorted(ref_surplus.tolist())
            st, ref_st = eng.chain_states()[0], chain.state()
            assert (int(st["active"]), int(st["active_cell"]), int(st["kept_kind"])) == \
                (ref_st.active, ref_st.active_cell, ref_st.kept_kind)
    assert boundaries == int((records["kind"] == 3).sum()) >= 2


def test_root_unit_active_mode_reference_trace_replay(oracle):
    """The shipped dipoles/dipole_motion.ini on the device (molecule_kernel<..., ROOT_MODE>): the independent active unit
    alternates between a leaf unit and the ROOT unit of a dipole (RootLeafUnitActiveSwitcher,
    root_leaf_unit_active_switcher.py:102-228; root-unit-active composite-object and two-leaf handlers). Every event of the
    two reference traces -- without and with the shipped sampling events, which the device meets as time limits with a
    kept candidate -- in stretches that start from the oracle's state (bit-exact with the reference over both traces);
    which unit is active after every event (EcmcEventRecord.mode) is compared as well."""
    g = tu.load_trace("trace_dipole_motion")
    records = g["records"]
    length = float(g["meta_system_length"])
    build = tu.dipole_motion_builder_of
    stretch = 50
    chain = oracle.OracleChain(build(g, oracle.ProgramBuilder))
    chain.set_positions(g["positions0"], tu.charges_of(g))
    chain.set_roots(g["roots0"])
    chain.start(stream=int(g["seed"][1]))

    def resynchronise(eng):
        eng.upload_positions(chain.positions()[None], tu.charges_of(g)[None])
        eng.upload_roots(chain.roots()[None])
        eng.set_chain_states(np.frombuffer(bytes(chain.state()), dtype=abi.chain_state_dtype()))
        occupants, surplus = chain.cells()
        eng.set_cells(occupants[None], [surplus])

    with engine.Engine(build(g, ProgramBuilder), n_chains=1) as eng:
        assert "root mode" in eng.kernel_name()
        eng.upload_positions(g["positions0"][None], tu.charges_of(g)[None])
        eng.upload_roots(g["roots0"][None])
        eng.start(first_stream=int(g["seed"][1]))
        st, ref_st = eng.chain_states()[0], chain.state()
        assert (st["switch_q"], st["switch_r"], st["mode"]) == (ref_st.switch_q, ref_st.switch_r, 0)
        for done in range(0, len(records), stretch):
            count = min(stretch, len(records) - done)
            if done:
                resynchronise(eng)
            rec, stats = eng.run_recorded(max_events=count, records_per_chain=count)
            assert stats["events"] == count and stats["capacity_errors"] == 0
            assert_records_match(rec[0], records[done:done + count], length, f"dipole motion[{done}:{done + count}]")
            assert np.array_equal(rec[0]["mode"], records["reserved"][done:done + count])
            n, ours = chain.run(max_events=count, record=count)
            assert n == count and tu.records_equal_discrete(ours, records[done:done + count])
            assert np.max(np.abs(eng.download_positions()[0] - chain.positions())) < RTOL
            assert np.max(np.abs(eng.download_roots()[0] - chain.roots())) < RTOL
            st, ref_st = eng.chain_states()[0], chain.state()
            for field in ("mode", "active", "eoc_next_active", "event_counter"):
                assert int(st[field]) == int(getattr(ref_st, field)), field
            for field in ("switch_q", "eoc_last_q", "eoc_q"):
                assert st[field] == getattr(ref_st, field), field
            for field in ("switch_r", "eoc_last_r", "eoc_r"):
                assert abs(st[field] - getattr(ref_st, field)) < RTOL, field

    # with sampling events: six time limits per stretch, the candidates that survive them are the device's own
    g = tu.load_trace("trace_dipole_motion_sampling")
    records, host = g["records"], g["host_times"]
    chain = oracle.OracleChain(build(g, oracle.ProgramBuilder))
    chain.set_positions(g["positions0"], tu.charges_of(g))
    chain.set_roots(g["roots0"])
    chain.start(stream=int(g["seed"][1]))
    kept_in_root_mode = 0
    with engine.Engine(build(g, ProgramBuilder), n_chains=1) as eng:
        eng.upload_positions(g["positions0"][None], tu.charges_of(g)[None])
        eng.upload_roots(g["roots0"][None])
        eng.start(first_stream=int(g["seed"][1]))
        done = 0
        for k, (events_before, q, r) in enumerate(host):
            if k and k % 6 == 0:
                resynchronise(eng)
            rec, stats = eng.run_recorded(until=(q, r), records_per_chain=200)
            n, ours = chain.run(until=(q, r), record=200)
            assert stats["events"] == n and done + n == int(events_before)
            assert tu.records_equal_discrete(ours, records[done:done + n])
            if n:
                assert_records_match(rec[0][:n], records[done:done + n], length, f"dipole motion, sample {k}")
                assert np.array_equal(rec[0]["mode"][:n], records["reserved"][done:done + n])
            done += n
            st, ref_st = eng.chain_states()[0], chain.state()
            assert (st["time_q"], st["time_r"]) == (q, r)
            assert (int(st["pending_kind"]), int(st["pending_target"]), int(st["mode"])) == \
                (ref_st.pending_kind, ref_st.pending_target, ref_st.mode)
            kept_in_root_mode += int(ref_st.mode == 1 and ref_st.pending_kind == abi.EVENT_PAIR)
            assert np.max(np.abs(eng.download_positions()[0] - chain.positions())) < RTOL
            assert np.max(np.abs(eng.download_roots()[0] - chain.roots())) < RTOL
    assert len(host) > 100 and kept_in_root_mode > 20


def test_sequential_direction_reference_trace_replay(oracle):
    """General velocities on the device (disk_kernel): the shipped hard_disk_dipoles.ini -- no cell system, 160 hard-disk
    candidates and the tether per event, the velocity rotated by 20 degrees at every end of chain
    (single_independent_active_sequential_direction_end_of_chain_event_handler.py:101-122) -- every one of the 5000 events
    of the running reference from the shipped start configuration, in stretches of 40 events that each start from the
    oracle's state (bit-exact with the reference over the whole trace; hard disks are chaotic). After every stretch the
    leaf and root positions and the velocities of the active leaf and of its root unit are compared."""
    g = tu.load_trace("trace_hard_disk_dipoles_sequential")
    records = g["records"]
    length = float(g["meta_system_length"])
    stretch = 40
    chain = oracle.OracleChain(tu.sequential_dipole_builder_of(g, oracle.ProgramBuilder))
    chain.set_positions(g["positions0"])
    chain.set_roots(g["roots0"])
    chain.start(stream=int(g["seed"][1]))
    totals = dict.fromkeys(["bond_events", "factor_pair_events", "end_of_chain_events"], 0)
    with engine.Engine(tu.sequential_dipole_builder_of(g, ProgramBuilder), n_chains=1) as eng:
        assert "disk_kernel" in eng.kernel_name(record=True)
        eng.upload_positions(g["positions0"][None])
        eng.upload_roots(g["roots0"][None])
        eng.start(first_stream=int(g["seed"][1]))
        for done in range(0, len(records), stretch):
            count = min(stretch, len(records) - done)
            if done:
                eng.upload_positions(chain.positions()[None])
                eng.upload_roots(chain.roots()[None])
                eng.set_chain_states(np.frombuffer(bytes(chain.state()), dtype=abi.chain_state_dtype()))
            rec, stats = eng.run_recorded(max_events=count, records_per_chain=count)
            assert stats["events"] == count and stats["capacity_errors"] == 0
            assert_records_match(rec[0], records[done:done + count], length, f"sequential[{done}:{done + count}]")
            n, ours = chain.run(max_events=count, record=count)
            assert n == count and tu.records_equal_discrete(ours, records[done:done + count])
            assert np.max(np.abs(eng.download_positions()[0] - chain.positions())) < RTOL * max(1.0, length)
            assert np.max(np.abs(eng.download_roots()[0] - chain.roots())) < RTOL * max(1.0, length)
            st, ref = eng.chain_states()[0], chain.state()
            assert int(st["active"]) == ref.active and int(st["eoc_next_active"]) == ref.eoc_next_active
            assert np.max(np.abs(st["velocity"] - np.array(ref.velocity[:]))) < RTOL
            assert np.max(np.abs(st["root_velocity"] - np.array(ref.root_velocity[:]))) < RTOL
            for key in totals:
                totals[key] += stats[key]
    ref_stats = chain.stats()
    assert all(totals[key] == ref_stats[key] for key in totals), (totals, ref_stats)
    assert totals["end_of_chain_events"] > 150 and totals["bond_events"] > 1500 and totals["factor_pair_events"] > 2000


def _sequential_dipole_batch(n_chains, columns=4, rows=7, length=9.0, seed=11, chain_time=0.7, delta_phi=23.0):
    """columns x rows hard-disk dipoles on a jittered lattice, no cell system, general velocities."""
    hs = abi.EcmcPotential.make(abi.POT_HARD_SPHERE, 0.476190476190476)
    tether = abi.EcmcPotential.make(abi.POT_HARD_DIPOLE, 0.952380952380952, 1.047619047619048)
    n_roots = columns * rows

    def build(cls):
        pb = cls(2, 2 * n_roots, length, 1.0, [1, 1], 0, chain_time=chain_time, seed=seed, no_cells=True)
        pb.set_composite(2, bonds=[(0, 1)], bond_potential=tether)
        pb.set_sequential_direction(delta_phi, hs, [(0, 0), (0, 1), (1, 0), (1, 1)])
        return pb
    rng = np.random.default_rng(500 + seed)
    grid = np.stack(np.meshgrid(np.arange(columns), np.arange(rows), indexing="ij"), axis=-1).reshape(-1, 2)
    roots = np.empty((n_chains, n_roots, 2))
    leaves = np.empty((n_chains, n_roots, 2, 2))
    for c in range(n_chains):
        centre = (grid + 0.5) * np.array([length / columns, length / rows]) + rng.uniform(-0.02, 0.02, size=(n_roots, 2))
        angle = rng.uniform(-0.05, 0.05, size=n_roots)
        half = 0.5 * rng.uniform(0.96, 1.04, size=n_roots)
        offset = np.stack([np.cos(angle), np.sin(angle)], axis=1) * half[:, None]
        roots[c] = centre % length
        leaves[c, :, 0] = (centre + offset) % length
        leaves[c, :, 1] = (centre - offset) % length
    return build, roots, leaves.reshape(n_chains, 2 * n_roots, 2)


def test_sequential_direction_batch_against_oracle(oracle):
    """Seeded batch of chains with general velocities against the oracle, re-seeded every 40 events; every chain has its
    own start configuration and random stream (the end of chain draws the next active leaf)."""
    build, roots, leaves = _sequential_dipole_batch(n_chains=7)
    stats = _compare_batch_with_oracle(oracle, build(ProgramBuilder), leaves, None, 1600, 21, "sequential dipoles",
                                       resync_every=40, roots=roots)
    assert stats["factor_pair_events"] > 200 and stats["bond_events"] > 500 and stats["end_of_chain_events"] > 50


def test_sequential_direction_time_limits_keep_candidates(oracle):
    """Host control events between the device events (the shipped file samples the polarization every 10.01): the kept
    candidate, its in-state (both coordinates of the leaf and of the root unit) and the time slices at the sampling
    times against the oracle; between two limits the chains run at most 30 events, so no re-seeding is needed."""
    build, roots, leaves = _sequential_dipole_batch(n_chains=3, seed=12)
    length = 9.0
    chains = []
    for c in range(3):
        chain = oracle.OracleChain(build(oracle.ProgramBuilder))
        chain.set_positions(leaves[c])
        chain.set_roots(roots[c])
        chain.start(stream=40 + c)
        chains.append(chain)
    with engine.Engine(build(ProgramBuilder), n_chains=3) as eng:
        eng.upload_positions(leaves)
        eng.upload_roots(roots)
        eng.start(first_stream=40)
        for k in range(1, 25):
            t = 0.37 * k
            until = (float(np.floor(t)), float(t - np.floor(t)))
            rec, stats = eng.run_recorded(until=until, records_per_chain=64)
            states = eng.chain_states()
            positions, root_positions = eng.download_positions(), eng.download_roots()
            for c, chain in enumerate(chains):
                n, ref = chain.run(until=until, record=64)
                assert stats["events"] >= n
                if n:
                    assert_records_match(rec[c][:n], ref[:n], length, f"limit {k} chain {c}")
                st = chain.state()
                assert (states[c]["time_q"], states[c]["time_r"]) == until
                assert int(states[c]["pending_kind"]) == st.pending_kind and int(states[c]["pending_target"]) == st.pending_target
                assert int(states[c]["event_counter"]) == st.event_counter
                assert np.max(np.abs(positions[c] - chain.positions())) < RTOL * length
                assert np.max(np.abs(root_positions[c] - chain.roots())) < RTOL * length
            # the chains are chaotic: continue from the oracle's state
            eng.upload_positions(np.stack([chain.positions() for chain in chains]))
            eng.upload_roots(np.stack([chain.roots() for chain in chains]))
            eng.set_chain_states(np.concatenate([np.frombuffer(bytes(chain.state()), dtype=abi.chain_state_dtype())
                                                 for chain in chains]))


@pytest.mark.parametrize("name", tu.COMPOSITE_CELL_BOUNDING_TRACES)
def test_composite_cell_bounding_reference_trace_replay(oracle, name):
    """The shipped dipoles/cell_bounded.ini (four dipoles) on the device: composite-object cell-bounding candidates for
    the objects in cells that are not nearby, every event of the reference trace, in resynchronised stretches."""
    g = tu.load_trace(name)
    records = g["records"]
    length = float(g["meta_system_length"])
    stretch = 100
    chain = oracle.OracleChain(tu.dipole_cell_bounded_builder_of(g, oracle.ProgramBuilder))
    chain.set_positions(g["positions0"], tu.charges_of(g))
    chain.set_roots(g["roots0"])
    chain.start(stream=int(g["seed"][1]))
    with engine.Engine(tu.dipole_cell_bounded_builder_of(g, ProgramBuilder), n_chains=1) as eng:
        eng.upload_positions(g["positions0"][None], tu.charges_of(g)[None])
        eng.upload_roots(g["roots0"][None])
        eng.start(first_stream=int(g["seed"][1]))
        for done in range(0, len(records), stretch):
            count = min(stretch, len(records) - done)
            if done:
                eng.upload_positions(chain.positions()[None], tu.charges_of(g)[None])
                eng.upload_roots(chain.roots()[None])
                eng.set_chain_states(np.frombuffer(bytes(chain.state()), dtype=abi.chain_state_dtype()))
                occupants, surplus = chain.cells()
                eng.set_cells(occupants[None], [surplus])
            rec, stats = eng.run_recorded(max_events=count, records_per_chain=count)
            assert stats["events"] == count and stats["capacity_errors"] == 0
            assert_records_match(rec[0], records[done:done + count], length, f"{name}[{done}:{done + count}]")
            n, ours = chain.run(max_events=count, record=count)
            assert n == count and tu.records_equal_discrete(ours, records[done:done + count])
            assert np.max(np.abs(eng.download_positions()[0] - chain.positions())) < RTOL * max(1.0, length)
            occ, surplus = eng.cells()
            assert np.array_equal(occ[0], chain.cells()[0])


def test_no_cells_batch_against_oracle(oracle):
    """The same structure on a batch: 40 atoms of both signs per chain (two passes of pair lanes), 9 chains."""
    n, length = 40, 1.0
    mic = abi.EcmcPotential.make(abi.POT_MERGED_IMAGE_COULOMB, 1.0, 3.45, 6, 2)
    ipcb = abi.EcmcPotential.make(abi.POT_INVERSE_POWER_COULOMB_BOUNDING, 1.5837)
    pb = ProgramBuilder(3, n, length, 2.0, [1, 1, 1], 0, chain_time=0.78965, seed=33, no_cells=True)
    pb.set_pair(abi.PAIR_TWO_LEAF_UNIT_BOUNDING, mic, ipcb, use_charge=True)
    rng = np.random.default_rng(78)
    n_chains = 9
    positions = rng.uniform(0.0, length, size=(n_chains, n, 3))
    charges = np.where(rng.random((n_chains, n)) < 0.5, 1.0, -1.0)
    stats = _compare_batch_with_oracle(oracle, pb, positions, charges, 600, 4, "no cells", resync_every=100)
    assert stats["pair_events"] > 1000 and stats["boundary_events"] == 0 and stats["veto_events"] == 0


def test_coulomb_cell_bounding_batch_against_oracle(oracle):
    """Far field through TwoLeafUnitCellBoundingPotentialEventHandler: one candidate per occupied non-nearby cell
    (coulomb_atoms/cell_bounded.ini shape), charges of both signs, more occupied far cells than one pass holds."""
    n, cells, length = 60, [5, 4, 5], 1.0
    mic = abi.EcmcPotential.make(abi.POT_MERGED_IMAGE_COULOMB, 1.0, 3.45, 6, 2)
    ipcb = abi.EcmcPotential.make(abi.POT_INVERSE_POWER_COULOMB_BOUNDING, 1.5837)
    bounds, _ = oracle.inner_point_derivative_bounds(mic, length, cells, 1, prefactor=1.5, points_per_side=3,
                                                     target_charge=1.0, uses_charges=True)
    pb = ProgramBuilder(3, n, length, 2.0, cells, 1, max_occupants=1, max_surplus=n, chain_time=0.78965, seed=33)
    pb.set_pair(abi.PAIR_TWO_LEAF_UNIT_BOUNDING, mic, ipcb, use_charge=True)
    pb.set_cell_bounding(mic, bounds, use_charge=True, target_charge=1.0)
    rng = np.random.default_rng(78)
    n_chains = 7
    positions = rng.uniform(0.0, length, size=(n_chains, n, 3))
    charges = np.where(rng.random((n_chains, n)) < 0.5, 1.0, -1.0)
    charges[0] = 1.0
    stats = _compare_batch_with_oracle(oracle, pb, positions, charges, 900, 3, "coulomb cell bounding")
    assert stats["pair_events"] > 500 and stats["veto_events"] == 0


def test_hard_disks_2d_against_oracle(oracle):
    """2D hard disks in cells without a cell-veto handler (the potential needs no potential change)."""
    n, cells, length = 30, 5, 10.0
    hs = abi.EcmcPotential.make(abi.POT_HARD_SPHERE, 0.4)
    pb = ProgramBuilder(2, n, length, 1.0, [cells, cells], 1, max_occupants=4, max_surplus=n, chain_time=2.5, seed=4)
    pb.set_pair(abi.PAIR_TWO_LEAF_UNIT, hs)
    rng = np.random.default_rng(8)
    grid = np.stack(np.meshgrid(np.arange(6), np.arange(6), indexing="ij"), axis=-1).reshape(-1, 2)[:n]
    positions = np.stack([(grid + 0.5) * (length / 6) + rng.uniform(-0.3, 0.3, size=(n, 2)) for _ in range(5)])
    stats = _compare_batch_with_oracle(oracle, pb, positions, None, 600, 3, "hard disks", resync_every=40)
    assert stats["pair_events"] > 50 and stats["veto_events"] == 0


def _dipole_batch(n_chains, columns=5, rows=9, length=11.0, seed=3, chain_time=1.0, max_occupants=6):
    """columns x rows hard-disk dipoles lying along x on a jittered lattice (the C1 structure, smaller)."""
    hs = abi.EcmcPotential.make(abi.POT_HARD_SPHERE, 0.476190476190476)
    tether = abi.EcmcPotential.make(abi.POT_HARD_DIPOLE, 0.952380952380952, 1.047619047619048)
    n_roots = columns * rows
    pb = ProgramBuilder(2, 2 * n_roots, length, 1.0, [11, 11], 1, max_occupants=max_occupants, max_surplus=0,
                        chain_time=chain_time, seed=seed)
    pb.set_pair(abi.PAIR_TWO_LEAF_UNIT, hs)
    pb.set_composite(2, bonds=[(0, 1)], bond_potential=tether)
    rng = np.random.default_rng(300 + seed)
    grid = np.stack(np.meshgrid(np.arange(columns), np.arange(rows), indexing="ij"), axis=-1).reshape(-1, 2)
    roots = np.empty((n_chains, n_roots, 2))
    leaves = np.empty((n_chains, n_roots, 2, 2))
    for c in range(n_chains):
        centre = (grid + 0.5) * np.array([length / columns, length / rows]) + rng.uniform(-0.02, 0.02, size=(n_roots, 2))
        angle = rng.uniform(-0.05, 0.05, size=n_roots)
        half = 0.5 * rng.uniform(0.96, 1.04, size=n_roots)
        offset = np.stack([np.cos(angle), np.sin(angle)], axis=1) * half[:, None]
        roots[c] = centre % length
        leaves[c, :, 0] = (centre + offset) % length
        leaves[c, :, 1] = (centre - offset) % length
    return pb, roots, leaves.reshape(n_chains, 2 * n_roots, 2)


def test_hard_disk_dipoles_against_oracle(oracle):
    """Composite point objects (C1 structure): hard-sphere pairs through leaf-level cells with several occupants,
    the hard-dipole tether of the factor type map, root units time-sliced with their active leaf, end of chain drawing
    (root, child). Chaotic: re-seeded from the oracle every 40 events."""
    pb, roots, leaves = _dipole_batch(n_chains=6)
    stats = _compare_batch_with_oracle(oracle, pb, leaves, None, 1200, 5, "dipoles", resync_every=40, roots=roots)
    assert stats["pair_events"] > 300 and stats["bond_events"] > 100 and stats["end_of_chain_events"] > 20


def test_hard_disk_dipoles_reference_trace():
    """The shipped hard_disk_dipoles_cells.ini from the shipped start configuration (tests/golden/
    trace_hard_disk_dipoles.npz, recorded from the running reference): every segment between two snapshots of the
    reference starts from the reference's own state and must reproduce its events."""
    g = tu.load_trace("trace_hard_disk_dipoles")
    records = g["records"]
    length = float(g["meta_system_length"])
    pb = tu.dipole_builder_of(g, ProgramBuilder)
    with engine.Engine(pb, n_chains=1) as eng:
        eng.upload_positions(g["positions0"][None])
        eng.upload_roots(g["roots0"][None])
        eng.start(first_stream=int(g["seed"][1]))
        rec, stats = eng.run_recorded(max_events=120, records_per_chain=120)
        assert_records_match(rec[0], records[:120], length, "dipole trace")
        assert stats["bond_events"] > 5 and stats["capacity_errors"] == 0


def test_hard_disk_dipoles_free_running_horizon():
    """How long a FREE-RUNNING device chain of the shipped hard-disk dipole configuration follows the reference trace
    (no re-seeding from the oracle): hard-disk chains are chaotic, so a last-bit difference in one collision time (the
    device regroups the arithmetic) grows by a factor per collision until a different disk is hit. The test measures the
    horizon -- the first event whose discrete fields differ -- and requires every event before it to agree in time within
    the error growth that horizon implies; it documents the number rather than hiding it behind re-seeding."""
    g = tu.load_trace("trace_hard_disk_dipoles")
    records = g["records"]
    length = float(g["meta_system_length"])
    pb = tu.dipole_builder_of(g, ProgramBuilder)
    n = len(records)
    with engine.Engine(pb, n_chains=1) as eng:
        eng.upload_positions(g["positions0"][None])
        eng.upload_roots(g["roots0"][None])
        eng.start(first_stream=int(g["seed"][1]))
        rec, _ = eng.run_recorded(max_events=n, records_per_chain=n)
    ours = rec[0]
    differs = np.zeros(n, dtype=bool)
    for field in tu.DISCRETE_FIELDS:
        differs |= ours[field] != records[field]
    horizon = int(np.nonzero(differs)[0][0]) if differs.any() else n
    errors = np.abs((ours["time_q"] - records["time_q"]) + (ours["time_r"] - records["time_r"]))[:horizon]
    print(f"free-running horizon: {horizon} of {n} events; time error after 100 events {errors[min(100, horizon - 1)]:.2e}, "
          f"at the horizon {errors[-1]:.2e}")
    assert horizon >= 120
    assert np.all(errors[:120] < 1e-12 * np.maximum(1.0, records["time_q"][:120]))


@pytest.mark.parametrize("name", tu.WATER_TRACES)
def test_water_reference_trace_replay(name):
    """C4: the shipped water/coulomb_cell_veto_lj_inverted.ini recorded from the running reference (composite-object
    Coulomb pair and cell-veto events with inside-first lifting, Lennard-Jones between the oxygens, harmonic bonds,
    bending with ratio lifting, root-level cells, factors kept across root cell-boundary events): every event of the
    trace and the reference's snapshots of leaves, roots and cell occupancy."""
    g = tu.load_trace(name)
    records = g["records"]
    length = float(g["meta_system_length"])
    with engine.Engine(tu.water_builder_of(g, ProgramBuilder), n_chains=1) as eng:
        eng.upload_positions(g["positions0"][None], g["charges"][None])
        eng.upload_roots(g["roots0"][None])
        eng.start(first_stream=int(g["seed"][1]))
        done = 0
        snaps = list(g["snap_event"])
        for k, event in enumerate(snaps + [len(records)]):
            count = int(event) - done
            if count:
                rec, stats = eng.run_recorded(max_events=count, records_per_chain=count)
                assert stats["events"] == count and stats["capacity_errors"] == 0
                assert_records_match(rec[0], records[done:event], length, f"{name}[{done}:{event}]")
            done = int(event)
            if k < len(snaps):
                assert np.max(np.abs(eng.download_positions()[0] - g["snap_positions"][k])) < RTOL * length
                assert np.max(np.abs(eng.download_roots()[0] - g["snap_roots"][k])) < RTOL * length
                occ, surplus = eng.cells()
                assert np.array_equal(occ[0], g["snap_occupants"][k])
        assert np.max(np.abs(eng.download_positions()[0] - g["final_positions"])) < RTOL * length
        assert np.max(np.abs(eng.download_roots()[0] - g["final_roots"])) < RTOL * length


@pytest.mark.parametrize("lifting", [abi.LIFTING_INSIDE_FIRST, abi.LIFTING_OUTSIDE_FIRST, abi.LIFTING_RATIO])
def test_water_batch_against_oracle(oracle, lifting):
    """Seeded water chains (12 molecules, the dense geometry of trace_water_dense with the reference's own cell-veto
    tables) against the oracle, event by event, including surplus molecules (more molecules than a cell holds), for
    each lifting scheme of the composite-object handlers (the oracle's schemes are pinned to the reference's in
    tests/test_oracle_potentials.py)."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import configs
    g = tu.load_trace("trace_water_dense")
    pb = tu.water_builder_of(g, ProgramBuilder)
    pb.program.composite_lifting = lifting
    n_chains, n_roots = 5, int(g["meta_n"]) // 3
    roots = np.empty((n_chains, n_roots, 3))
    leaves = np.empty((n_chains, 3 * n_roots, 3))
    for c in range(n_chains):
        r, l = configs.water_start(n_roots, float(g["meta_system_length"]), seed=40 + c, jitter=0.2 + 0.1 * c)
        roots[c], leaves[c] = r, l.reshape(-1, 3)
    charges = np.tile(g["charges"], (n_chains, 1))
    n_events = 1500 if lifting == abi.LIFTING_INSIDE_FIRST else 700
    stats = _compare_batch_with_oracle(oracle, pb, leaves, charges, n_events, 60, "water", roots=roots)
    assert stats["pair_events"] > n_events // 5 and stats["veto_events"] > n_events and stats["bond_events"] > n_events // 15
    assert stats["factor_pair_events"] > n_events // 15


def test_single_particle_chain(oracle):
    """One particle: only cell-boundary, rejected cell-veto (empty cells) and end-of-chain events."""
    pb, positions = _lj_batch(oracle, n_chains=3, n=1, cells=4, length=4.6, seed=2, chain_time=0.9)
    stats = _compare_batch_with_oracle(oracle, pb, positions, None, 300, 11, "single particle")
    assert stats["pair_events"] == 0 and stats["veto_accepted"] == 0 and stats["boundary_events"] > 0


def test_time_limits_keep_candidates(oracle):
    """Interrupting at host control times must not change the event sequence (kept candidates, no new draws)."""
    pb, positions = _lj_batch(oracle, n_chains=8, seed=6)
    with engine.Engine(pb, n_chains=8) as free, engine.Engine(pb, n_chains=8) as stepped:
        for eng in (free, stepped):
            eng.upload_positions(positions)
            eng.start(first_stream=40)
        rec_free, stats_free = free.run_recorded(until=(2.0, 0.5), records_per_chain=4000)
        parts = [[] for _ in range(8)]
        total = 0
        for k in range(1, 26):
            t = oracle.time_from_float(0.1 * k)
            rec, stats = stepped.run_recorded(until=t, records_per_chain=1000)
            total += stats["events"]
            for c in range(8):
                parts[c].append(rec[c][rec[c]["kind"] != abi.EVENT_NONE])
            st = stepped.chain_states()
            assert np.all(st["time_q"] == t[0]) and np.all(st["time_r"] == t[1])
        assert total == stats_free["events"]
        for c in range(8):
            a = np.concatenate(parts[c])
            b = rec_free[c][rec_free[c]["kind"] != abi.EVENT_NONE]
            assert len(a) == len(b)
            for f in ("kind", "target", "accepted", "new_active"):
                assert np.array_equal(a[f], b[f]), (c, f)
            assert tu.max_time_error(a, b) < RTOL
        # the oracle interrupted at the same times agrees event by event
        chain = oracle.OracleChain(pb)
        chain.set_positions(positions[0])
        chain.start(stream=40)
        ref_parts = []
        for k in range(1, 26):
            _, rec = chain.run(until=oracle.time_from_float(0.1 * k), record=1000)
            ref_parts.append(rec)
        assert_records_match(np.concatenate(parts[0]), np.concatenate(ref_parts), float(pb.program.system_length), "stepped")


def test_split_launches_equal_one_launch(oracle):
    pb, positions = _lj_batch(oracle, n_chains=16, seed=12)
    with engine.Engine(pb, n_chains=16) as one, engine.Engine(pb, n_chains=16) as two:
        for eng in (one, two):
            eng.upload_positions(positions)
            eng.start(first_stream=0)
        one.run(max_events=900)
        s1 = one.sync()
        for _ in range(3):
            two.run(max_events=300)
        s2 = two.sync()
        # (candidates: with candidate pruning the count of EVALUATED candidates depends on where a launch rebuilt its
        # candidate list; the committed events do not)
        assert {k: v for k, v in s1.items() if k != "candidates"} == {k: v for k, v in s2.items() if k != "candidates"}
        assert np.array_equal(one.download_positions(), two.download_positions())
        assert np.array_equal(one.chain_states(), two.chain_states())
        assert two.kernel_launches == 3 and two.kernel_seconds > 0.0


def test_submitted_host_steps_equal_blocking_host_steps(oracle):
    """ecmc_submit_from_host x 3 + ecmc_wait (steps chained through two host buffers, ordered on the device slice by slice)
    against three blocking ecmc_run_from_host calls: the same configurations, the same counters."""
    pb, positions = _lj_batch(oracle, n_chains=600, seed=21)
    buffers = [np.ascontiguousarray(positions.copy()), np.empty_like(positions)]
    try:
        import torch
        pinned = [torch.from_numpy(b).pin_memory() for b in buffers]  # page-locked: the copies really are asynchronous
        buffers = [p.numpy() for p in pinned]
    except ImportError:
        pass
    with engine.Engine(pb, n_chains=600) as blocking, engine.Engine(pb, n_chains=600) as queued:
        current, totals = positions, {}
        for k in range(3):
            current, stats = blocking.run_from_host(current, first_stream=1000 * k, max_events=150)
            for key, value in stats.items():
                totals[key] = totals.get(key, 0) + value
        for k in range(3):
            queued.submit_from_host(buffers[k % 2], first_stream=1000 * k, max_events=150, out=buffers[(k + 1) % 2])
        stats = queued.wait()
        assert np.array_equal(buffers[1], current)
        assert {k: v for k, v in stats.items() if k != "candidates"} == {k: v for k, v in totals.items() if k != "candidates"}
        assert np.array_equal(queued.download_positions(), blocking.download_positions())
        with pytest.raises(ValueError):
            queued.submit_from_host(positions[:10], max_events=5, out=buffers[0])


@pytest.mark.parametrize("fused", [True, False])
def test_sparse_write_back_equals_the_full_copy(oracle, fused):
    """ecmc_submit_from_host_sparse: three steps chained IN PLACE through one pinned buffer (ecmc_host_alloc), the device
    writing only the coordinates of the particles that moved, against three blocking full-copy ecmc_run_from_host calls:
    the buffer ends as the same complete configuration bit for bit; ordinary (pageable) memory is refused.
    fused: the whole step is one launch per chain slice (lj_spec_kernel<HOST>: the kernel reads the pinned buffer itself,
    bins, runs the events and writes every hand-over through) -- the bytes written are one position per hand-over plus
    the last active particle of every chain; otherwise copy -> pack -> start -> events -> write-back of the particles that
    differ from the step's input, whose bytes are exactly those."""
    pb, positions = _lj_batch(oracle, n_chains=600, seed=22)
    buffer = engine.pinned_array(positions.shape)
    buffer[...] = positions
    with engine.Engine(pb, n_chains=600) as blocking, engine.Engine(pb, n_chains=600) as sparse:
        sparse.set_option(engine.Engine.OPTION_FUSED_HOST_STEPS, int(fused))
        current, totals, moved = positions, {}, 0
        for k in range(3):
            after, stats = blocking.run_from_host(current, first_stream=1000 * k, max_events=150)
            moved += int(np.count_nonzero(np.any(after != current, axis=2)))
            current = after
            for key, value in stats.items():
                totals[key] = totals.get(key, 0) + value
        for k in range(3):
            sparse.submit_from_host(buffer, first_stream=1000 * k, max_events=150, out=buffer, sparse=True)
        stats = sparse.wait()
        assert np.array_equal(buffer, current)
        assert {k: v for k, v in stats.items() if k != "candidates"} == {k: v for k, v in totals.items() if k != "candidates"}
        assert np.array_equal(sparse.download_positions(), blocking.download_positions())
        assert 0 < moved < 0.5 * 3 * positions.shape[0] * positions.shape[1]
        if fused:
            hand_overs = totals["pair_events"] + totals["veto_accepted"] + totals["end_of_chain_events"]
            assert moved * 24 <= sparse.host_bytes_written <= (hand_overs + 3 * 600) * 24
        else:
            assert sparse.host_bytes_written == moved * 3 * 8
        pageable = positions.copy()
        with pytest.raises(RuntimeError, match="pinned"):
            sparse.submit_from_host(pageable, max_events=5, out=pageable, sparse=True)


@pytest.mark.parametrize("sparse", [True, False])
def test_continued_host_steps_equal_the_resident_run(oracle, sparse):
    """ECMC_OPTION_CONTINUE_HOST_STEPS: host steps that keep the lifting state of their chains on the device and take only
    the configuration from the host (cell occupancy rebuilt from it every step) against ONE engine that runs the same
    events without leaving the device. While no cell holds two particles the rebuilt occupancy is the one the resident
    chain carries, so positions and chain states agree bit for bit; chains that did put a particle into the surplus are
    only required to have run their events."""
    n_chains, steps, events = 300, 4, 150
    pb, positions = _lj_batch(oracle, n_chains=n_chains, seed=23)
    buffer = engine.pinned_array(positions.shape)
    buffer[...] = positions
    with engine.Engine(pb, n_chains=n_chains) as resident, engine.Engine(pb, n_chains=n_chains) as stepped:
        resident.upload_positions(positions)
        resident.start(first_stream=77)
        clean = np.ones(n_chains, dtype=bool)  # no surplus particle at any step boundary so far
        for k in range(steps):
            resident.run(max_events=events)
            resident.sync()
            if k + 1 < steps:
                clean &= np.array([len(surplus) == 0 for surplus in resident.cells()[1]])
        stepped.set_option(engine.Engine.OPTION_CONTINUE_HOST_STEPS, 1)
        stepped.upload_positions(positions)
        stepped.start(first_stream=77)
        total = 0
        for k in range(steps):
            if sparse:
                stepped.submit_from_host(buffer, first_stream=12345, max_events=events, out=buffer, sparse=True)
                total += stepped.wait()["events"]
            else:
                out, stats = stepped.run_from_host(buffer, first_stream=12345, max_events=events)
                buffer[...] = out
                total += stats["events"]
        assert total == n_chains * steps * events
        assert clean.sum() > n_chains // 2, clean.sum()
        ours, theirs = stepped.chain_states(), resident.chain_states()
        assert np.array_equal(ours["event_counter"], theirs["event_counter"])
        assert np.array_equal(ours["stream"], theirs["stream"])
        for field in ("active", "direction", "time_q", "time_r", "eoc_q", "eoc_r", "eoc_next_active", "active_cell"):
            assert np.array_equal(ours[field][clean], theirs[field][clean]), field
        assert np.array_equal(buffer[clean], resident.download_positions()[clean])


def test_pruned_launches_reach_the_same_state(oracle):
    """The kernel instantiations with and without event records must commit the same events: identical positions, cells
    and chain states bit for bit, the same event counts. (Launches without records skip pair candidates that provably
    cannot win -- ECMC_OPTION_PRUNE_CANDIDATES --; then only the count of finite candidates may differ.)"""
    pb, positions = _lj_batch(oracle, n_chains=64, n=100, cells=4, length=5.2, seed=14)
    with engine.Engine(pb, n_chains=64) as pruned, engine.Engine(pb, n_chains=64) as exact:
        for eng in (pruned, exact):
            eng.upload_positions(positions)
            eng.start(first_stream=7)
        pruned.run(max_events=4000)
        s1 = pruned.sync()
        _, s2 = exact.run_recorded(max_events=4000, records_per_chain=1)
        assert np.array_equal(pruned.download_positions(), exact.download_positions())
        assert np.array_equal(pruned.chain_states(), exact.chain_states())
        occ1, sur1 = pruned.cells()
        occ2, sur2 = exact.cells()
        assert np.array_equal(occ1, occ2) and all(a.tolist() == b.tolist() for a, b in zip(sur1, sur2))
        for key in ("events", "pair_events", "veto_events", "veto_accepted", "boundary_events", "end_of_chain_events",
                    "pair_targets"):
            assert s1[key] == s2[key], key
        assert 0 < s1["candidates"] <= s2["candidates"]
        assert s1["pair_events"] > 10000 and s1["veto_accepted"] > 100


def test_checkpoint_resume_is_bit_exact(oracle, tmp_path):
    """save_checkpoint / load_checkpoint (the role of the reference's dumping handler + resume.py): a run resumed in a
    fresh engine commits exactly the events of the uninterrupted run -- Lennard-Jones with surplus particles, and water
    with root units and kept factor candidates."""
    pb, positions = _lj_batch(oracle, n_chains=12, n=100, cells=4, length=5.2, seed=21)
    with engine.Engine(pb, n_chains=12) as straight, engine.Engine(pb, n_chains=12) as first:
        for eng in (straight, first):
            eng.upload_positions(positions)
            eng.start(first_stream=3)
        straight.run(max_events=1500)
        straight.sync()
        first.run(max_events=700)
        first.save_checkpoint(tmp_path / "lj.npz")
    with engine.Engine(pb, n_chains=12) as resumed, engine.Engine(pb, n_chains=12) as reference:
        resumed.load_checkpoint(tmp_path / "lj.npz")
        resumed.run(max_events=800)
        resumed.sync()
        reference.upload_positions(positions)
        reference.start(first_stream=3)
        reference.run(max_events=1500)
        reference.sync()
        assert np.array_equal(resumed.download_positions(), reference.download_positions())
        assert np.array_equal(resumed.chain_states(), reference.chain_states())
        assert np.array_equal(resumed.cells()[0], reference.cells()[0])
    g = tu.load_trace("trace_water_dense")
    wb = tu.water_builder_of(g, ProgramBuilder)
    charges = g["charges"][None]

    def started():
        eng = engine.Engine(wb, n_chains=1)
        eng.upload_positions(g["positions0"][None], charges)
        eng.upload_roots(g["roots0"][None])
        eng.start(first_stream=int(g["seed"][1]))
        return eng

    with started() as straight, started() as first:
        straight.run(max_events=1200)
        straight.sync()
        first.run(max_events=500)
        first.save_checkpoint(tmp_path / "water.npz")
        other, _ = _lj_batch(oracle, n_chains=12, n=100, cells=4, length=5.2, seed=22)  # the same shapes, another seed
        with engine.Engine(other, n_chains=12) as foreign:
            with pytest.raises(ValueError):
                foreign.load_checkpoint(tmp_path / "lj.npz")
        with engine.Engine(wb, n_chains=1) as resumed:
            resumed.load_checkpoint(tmp_path / "water.npz")  # the charges of the start configuration are in the file
            resumed.run(max_events=700)
            resumed.sync()
            assert np.array_equal(resumed.download_positions(), straight.download_positions())
            assert np.array_equal(resumed.download_roots(), straight.download_roots())
            assert np.array_equal(resumed.chain_states(), straight.chain_states())


def test_surplus_overflow_is_reported(oracle):
    pb, positions = _lj_batch(oracle, n_chains=2, n=100, cells=4, length=5.2, seed=9)
    pb.program.max_surplus = 4  # 100 particles in 64 cells need at least 36 surplus slots
    with engine.Engine(pb, n_chains=2) as eng:
        eng.upload_positions(positions)
        with pytest.raises(engine.EcmcError) as error:
            eng.start()
        assert error.value.status == abi.ECMC_ERR_CAPACITY


def test_invalid_programs_are_rejected(oracle):
    pb, positions = _lj_batch(oracle, n_chains=1)
    pb.program.abi_version = 99
    with pytest.raises(engine.EcmcError) as error:
        engine.Engine(pb)
    assert error.value.status == abi.ECMC_ERR_INVALID
    pb.program.abi_version = abi.ECMC_ABI_VERSION
    with engine.Engine(pb) as eng:
        with pytest.raises(engine.EcmcError) as error:
            eng.run(max_events=10)  # not started
        assert error.value.status == abi.ECMC_ERR_STATE
        eng.upload_positions(positions)
        eng.start()
        with pytest.raises(engine.EcmcError) as error:
            eng.run()  # neither a time nor an event limit
        assert error.value.status == abi.ECMC_ERR_INVALID
