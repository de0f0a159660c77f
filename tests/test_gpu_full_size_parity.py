"""Event-by-event parity with the oracle AT the BASELINE.json sizes ("parity hole on the headline
config"): the exact programs the bench times -- C2 (4096 chains x 1024 Lennard-Jones particles, 12^3 cells), C3
(Coulomb atoms, N = 512: ~95 pair targets per event, several passes over the candidate list) and C5 (one chain of
65536 particles, 48^3 cells) -- with a subsample of the chains of the full launch followed by `OracleChain`s.

Every compared stretch starts from identical inputs: the oracle chain takes the state the device holds (positions,
lifting state incl. random-stream counters, cell occupancy and surplus list), then both run the same events. Discrete
fields bit-exact, times and positions 1e-12 (north_star's tolerance), cell occupancy and surplus list bit-exact after
every stretch. Between the compared stretches the device runs on alone (tens of thousands of events per chain), so the
later stretches see long surplus lists (more gathered targets than one pass of lanes holds)."""
import ctypes

import numpy as np
import pytest

import trace_util as tu
from jellyfysh_b200 import abi, engine, workloads

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def _seed_oracle_from_device(oracle, builder, positions, charges, state, occupants, surplus):
    chain = oracle.OracleChain(builder)
    chain.set_positions(positions, charges)
    chain.set_cells(occupants, surplus)
    chain.set_state(abi.EcmcChainState.from_buffer_copy(state.tobytes()))
    return chain


def _compare_stretch(oracle, builder, eng, sample, events, charges, length, tag):
    """The chains `sample` of the launch against oracle chains seeded from the device state; returns the largest number
    of gathered pair targets an event of the sampled chains can have had (nearby occupants + surplus)."""
    positions = eng.download_positions()
    states = eng.chain_states()
    occupants, surplus = eng.cells()
    chains = {c: _seed_oracle_from_device(oracle, builder, positions[c], None if charges is None else charges[c],
                                          states[c], occupants[c], surplus[c]) for c in sample}
    records, stats = eng.run_recorded(max_events=events, records_per_chain=events)
    assert stats["events"] == eng.n_chains * events and stats["capacity_errors"] == 0, (tag, stats)
    final = eng.download_positions()
    final_states = eng.chain_states()
    final_occupants, final_surplus = eng.cells()
    longest = 0
    for c, chain in chains.items():
        n, ref = chain.run(max_events=events, record=events)
        assert n == events, (tag, c, n)
        ours = records[c]
        differs = np.zeros(events, dtype=bool)
        for field in tu.DISCRETE_FIELDS:
            differs |= ours[field] != ref[field]
        if differs.any():
            k = int(np.nonzero(differs)[0][0])
            raise AssertionError(f"{tag} chain {c}: first difference at event {k}: ours {ours[k]} oracle {ref[k]}")
        assert tu.max_time_error(ours, ref) < RTOL, (tag, c, tu.max_time_error(ours, ref))
        assert np.max(np.abs(ours["active_pos"] - ref["active_pos"])) < RTOL * max(1.0, length), (tag, c)
        assert np.max(np.abs(final[c] - chain.positions())) < RTOL * max(1.0, length), (tag, c)
        oracle_occupants, oracle_surplus = chain.cells()
        assert np.array_equal(final_occupants[c], oracle_occupants), (tag, c)
        assert final_surplus[c].tolist() == oracle_surplus.tolist(), (tag, c)
        st = chain.state()
        assert (int(final_states[c]["active"]), int(final_states[c]["direction"]), int(final_states[c]["active_cell"]),
                int(final_states[c]["event_counter"]), int(final_states[c]["eoc_next_active"])) == \
               (st.active, st.direction, st.active_cell, st.event_counter, st.eoc_next_active), (tag, c)
        longest = max(longest, len(oracle_surplus))
    return longest


def test_c2_bench_program_event_parity(oracle):
    """workloads.lennard_jones(1024, 12) exactly as bench.py runs it (4096 chains, lattice start, streams = chain ids):
    chains 0, 511, 2048 and 4095 of the launch, 2000 events each from the start, again after 24 000 and after 60 000
    events per chain (the window the bench times ends at 44 000)."""
    n_chains, n, cells = 4096, 1024, 12
    builder, length = workloads.lennard_jones(n_particles=n, cells_per_side=cells)
    start = workloads.lattice_start(n_chains, n, cells, length)
    sample = (0, 511, 2048, 4095)
    with engine.Engine(builder, n_chains=n_chains) as eng:
        eng.upload_positions(start)
        eng.start(first_stream=0)
        _compare_stretch(oracle, builder, eng, sample, 2000, None, length, "C2 from the start")
        eng.run(max_events=22000)
        eng.sync()
        _compare_stretch(oracle, builder, eng, sample, 2000, None, length, "C2 after 24k events")
        eng.run(max_events=34000)
        eng.sync()
        _compare_stretch(oracle, builder, eng, sample, 2000, None, length, "C2 after 60k events")


def test_c2_long_run_event_parity(oracle):
    """The same program with room for 192 surplus particles, 1.5 x 10^5 events per chain on the device alone, then 2000
    events against the oracle: the reference never promotes a surplus particle into a freed cell
    (single_active_cell_occupancy.py:186-193), so by then the candidate list holds many more targets than the 27
    nearby cells."""
    n_chains, n, cells = 1024, 1024, 12
    builder, length = workloads.lennard_jones(n_particles=n, cells_per_side=cells, max_surplus=192)
    start = workloads.lattice_start(n_chains, n, cells, length)
    with engine.Engine(builder, n_chains=n_chains) as eng:
        eng.upload_positions(start)
        eng.start(first_stream=9000)
        eng.run(max_events=150000)
        eng.sync()
        longest = _compare_stretch(oracle, builder, eng, (0, 1, 500, 1023), 2000, None, length, "C2 after 1.5e5 events")
        assert longest > 30  # more gathered targets than one pass of 32 lanes holds


def test_c2_crowded_cells_event_parity(oracle):
    """The same program with two particles in every second cell: 512 surplus particles per chain from the first event
    on (maximum size of the surplus list), ~540 gathered targets per event = 17 passes over the candidate list."""
    n_chains, n, cells = 32, 1024, 12
    builder, length = workloads.lennard_jones(n_particles=n, cells_per_side=cells, max_surplus=640)
    start = workloads.lattice_start(n_chains, n, cells, length, jitter=0.02)
    side = length / cells
    start[:, 512:] = start[:, :512]
    start[:, :512, 0] -= 0.26 * side
    start[:, 512:, 0] += 0.26 * side
    # (not exactly in line: a pair at zero distance from the line of motion divides by zero in the reference's own
    # LennardJonesPotential.displacement, inverse_power_potential.py:143 via abstracts.py:517)
    start[:, 512:, 1] += 0.11 * side
    start[:, 512:, 2] -= 0.07 * side
    with engine.Engine(builder, n_chains=n_chains) as eng:
        eng.upload_positions(start)
        eng.start(first_stream=700)
        assert min(len(s) for s in eng.cells()[1]) >= 511
        for stretch in range(3):
            _compare_stretch(oracle, builder, eng, (0, 13, 31), 400, None, length, f"C2 crowded stretch {stretch}")


def test_c3_coulomb_atoms_512_event_parity(oracle):
    """C3 at its largest size: 512 like charges (merged-image Coulomb, inverse-power bound, cell veto), uniform start:
    ~95 pair targets per event. Every pair event lifts and the chain is chaotic, hence stretches of 250 events."""
    n_chains, n = 256, 512
    builder, length = workloads.coulomb_atoms(n_particles=n, points_per_side=4)
    start = workloads.uniform_start(n_chains, n, length)
    charges = np.ones((n_chains, n))
    with engine.Engine(builder, n_chains=n_chains) as eng:
        eng.upload_positions(start, charges)
        eng.start(first_stream=0)
        for stretch in range(4):
            _compare_stretch(oracle, builder, eng, (0, 100, 255), 250, charges, length, f"C3 N=512 stretch {stretch}")
            eng.run(max_events=500)
            eng.sync()


def test_c5_single_large_chain_event_parity(oracle):
    """C5: the one chain of 65536 particles in 48^3 cells, 4000 events from the start and 2000 more after 2 x 10^5."""
    n, cells = 65536, 48
    builder, length = workloads.lennard_jones(n_particles=n, cells_per_side=cells)
    start = workloads.lattice_start(1, n, cells, length)
    with engine.Engine(builder, n_chains=1) as eng:
        eng.upload_positions(start)
        eng.start(first_stream=0)
        _compare_stretch(oracle, builder, eng, (0,), 4000, None, length, "C5 from the start")
        eng.run(max_events=200000)
        eng.sync()
        _compare_stretch(oracle, builder, eng, (0,), 2000, None, length, "C5 after 2e5 events")
