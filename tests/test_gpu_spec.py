"""The batched (speculative) Lennard-Jones / cell-veto kernel, csrc/ecmc_spec.cuh, against the one-event-at-a-time
kernel and the oracle: it must commit the same events, event for event.

* event records of the two kernels on the same chains: every field, incl. the candidate counts; times and positions
  bit for bit (the same additions in the same order);
* candidate pruning on / off (ecmc_run): the same final state of every chain, bit for bit, and the same event mix;
* odd launch partitions (1, 2, 3, 5, ... events per launch): the batch is cut at the event limit;
* time limits between the events of a batch; 4 and 8 lanes per event;
* the oracle on the bench program with pruning: device chains that ran pruned for tens of thousands of events, then
  event by event against oracle chains seeded from them (tests/test_gpu_full_size_parity.py does the unpruned part)."""
import numpy as np
import pytest

from jellyfysh_b200 import engine, workloads
from jellyfysh_b200.engine import Engine

pytestmark = pytest.mark.gpu

BATCHED, PRUNE, LANES, BLOCKS = (Engine.OPTION_BATCHED_EVENTS, Engine.OPTION_PRUNE_CANDIDATES, Engine.OPTION_LANES_PER_EVENT,
                                 Engine.OPTION_CHAIN_BLOCKS)


def _program(n=216, cells=7, max_surplus=64, chain_time=2.5):
    builder, length = workloads.lennard_jones(n_particles=n, cells_per_side=cells, chain_time=chain_time,
                                              points_per_side=3, max_surplus=max_surplus)
    return builder, length


def _started(builder, positions, first_stream, **options):
    eng = engine.Engine(builder, n_chains=len(positions))
    for option, value in options.items():
        eng.set_option({"batched": BATCHED, "prune": PRUNE, "lanes": LANES, "blocks": BLOCKS}[option], value)
    eng.upload_positions(positions)
    eng.start(first_stream=first_stream)
    return eng


def _state_of(eng):
    occupants, surplus = eng.cells()
    return eng.download_positions(), eng.chain_states(), occupants, [s.tolist() for s in surplus]


def _assert_same_state(one, two, tag):
    a, b = _state_of(one), _state_of(two)
    assert np.array_equal(a[0], b[0]), tag
    assert np.array_equal(a[1], b[1]), tag
    assert np.array_equal(a[2], b[2]), tag
    assert a[3] == b[3], tag


@pytest.mark.parametrize("lanes,blocks", [(4, 0), (8, 0), (4, 1)])
def test_batched_kernel_commits_the_events_of_the_single_event_kernel(lanes, blocks):
    """blocks = 1: lj_chain_kernel (one CTA of four warps per chain, 32 events per batch; engines of at most 148 chains),
    blocks = 0: lj_spec_kernel (one warp per chain, 8 or 4 events per batch)."""
    n_chains, n, cells, events = 96, 216, 7, 3000
    builder, length = _program(n, cells)
    positions = workloads.lattice_start(n_chains, n, cells, length, jitter=0.15)
    with _started(builder, positions, 11, batched=0) as single, \
            _started(builder, positions, 11, lanes=lanes, blocks=blocks) as batched:
        assert ("lj_chain_kernel" if blocks else "lj_spec_kernel") in batched.kernel_name(record=True)
        ref, ref_stats = single.run_recorded(max_events=events, records_per_chain=events)
        rec, stats = batched.run_recorded(max_events=events, records_per_chain=events)
        assert stats == ref_stats
        assert stats["pair_events"] > 0 and stats["veto_accepted"] > 0 and stats["boundary_events"] > 0 and \
            stats["end_of_chain_events"] > 0
        for field in rec.dtype.names:
            same = rec[field] == ref[field]
            if not np.all(same):
                chain, event = [int(v[0]) for v in np.nonzero(~same.reshape(n_chains, events, -1).all(axis=2))]
                raise AssertionError(f"{field}: chain {chain} event {event}: {rec[chain][event]} vs {ref[chain][event]}")
        _assert_same_state(single, batched, "after the recorded launch")


@pytest.mark.parametrize("blocks", [0, 1])
def test_pruning_does_not_change_the_chains(blocks):
    n_chains, n, cells = 128, 216, 7
    builder, length = _program(n, cells)
    positions = workloads.lattice_start(n_chains, n, cells, length, jitter=0.15)
    with _started(builder, positions, 500, prune=0, blocks=blocks) as full, \
            _started(builder, positions, 500, prune=1, blocks=blocks) as pruned, \
            _started(builder, positions, 500, batched=0) as single:
        assert ("lj_chain_kernel" if blocks else "lj_spec_kernel") in pruned.kernel_name()
        totals = []
        for eng in (full, pruned, single):
            for events in (1, 2, 3, 5, 8, 13, 968, 4000):
                eng.run(max_events=events)
            totals.append(eng.sync())
        _assert_same_state(full, pruned, "pruned vs unpruned")
        _assert_same_state(full, single, "batched vs single-event kernel")
        assert totals[0] == totals[2]
        for key in totals[0]:
            if key != "candidates":
                assert totals[0][key] == totals[1][key], key
        assert totals[1]["candidates"] < totals[0]["candidates"]  # the pruned run evaluated fewer candidates


@pytest.mark.parametrize("blocks", [0, 1])
def test_time_limits_cut_batches_like_the_single_event_kernel(blocks):
    n_chains, n, cells = 64, 216, 7
    builder, length = _program(n, cells, chain_time=0.7)
    positions = workloads.lattice_start(n_chains, n, cells, length, jitter=0.15)
    with _started(builder, positions, 40, batched=0) as single, \
            _started(builder, positions, 40, prune=1, blocks=blocks) as batched:
        for eng in (single, batched):
            for k in range(1, 40):
                eng.run(until=(float(k // 8), (k % 8) / 8.0))  # sampling times every 1/8
                eng.run(max_events=3)                           # a kept candidate fires (or not) in a short launch
            eng.run(until=(6.0, 0.0625))
            eng.sync()
        _assert_same_state(single, batched, "after interleaved time limits")
        states = batched.chain_states()
        assert np.all(states["time_q"] == 6.0) and np.all(states["time_r"] == 0.0625)


def test_pruned_bench_program_against_the_oracle(oracle):
    """The C2 program of bench.py, pruned launches of 1024 events as the bench issues them; then 1500 events of sampled
    chains against oracle chains seeded from the device (recorded launches never prune, so this checks that the pruned
    launches left a state from which the chain continues exactly like the oracle's) -- and the whole trajectory against
    an engine that never pruned."""
    from test_gpu_full_size_parity import _compare_stretch
    n_chains, n, cells = 1024, 1024, 12
    builder, length = workloads.lennard_jones(n_particles=n, cells_per_side=cells)
    positions = workloads.lattice_start(n_chains, n, cells, length)
    with _started(builder, positions, 0, prune=1) as pruned, _started(builder, positions, 0, batched=0) as single:
        for eng in (pruned, single):
            for _ in range(30):
                eng.run(max_events=1024)
            eng.sync()
        _assert_same_state(pruned, single, "30 launches of 1024 events")
        _compare_stretch(oracle, builder, pruned, (0, 17, 1023), 1500, None, length, "C2 pruned, after 30k events")


def test_single_large_chain_in_a_block_against_the_oracle(oracle):
    """C5 through lj_chain_kernel (the default for one chain): 60 000 pruned events of the chain of 65536 particles, then
    2000 events against an oracle chain seeded from the device, and the whole trajectory against the one-warp kernel."""
    from test_gpu_full_size_parity import _compare_stretch
    n, cells = 65536, 48
    builder, length = workloads.lennard_jones(n_particles=n, cells_per_side=cells)
    positions = workloads.lattice_start(1, n, cells, length)
    with _started(builder, positions, 0) as blocked, _started(builder, positions, 0, blocks=0) as warped:
        assert "lj_chain_kernel" in blocked.kernel_name() and "lj_spec_kernel" in warped.kernel_name()
        for eng in (blocked, warped):
            for events in (7, 1000, 58993):
                eng.run(max_events=events)
            eng.sync()
        _assert_same_state(blocked, warped, "60 000 events of the single chain")
        _compare_stretch(oracle, builder, blocked, (0,), 2000, None, length, "C5 in a block, after 60k events")


# ---- the Coulomb model of the batched kernel (C3: inverse-power Coulomb bound, merged-image Coulomb, charges) ------------
_COULOMB_PROGRAMS = {}


def _coulomb_program(n=64, chain_time=0.78965):
    """(builder, length); the cell-veto tables are built once per program and module."""
    if (n, chain_time) not in _COULOMB_PROGRAMS:
        _COULOMB_PROGRAMS[(n, chain_time)] = workloads.coulomb_atoms(n_particles=n, chain_time=chain_time, points_per_side=4)
    return _COULOMB_PROGRAMS[(n, chain_time)]


def _coulomb_started(builder, positions, charges, first_stream, **options):
    eng = engine.Engine(builder, n_chains=len(positions))
    for option, value in options.items():
        eng.set_option({"batched": BATCHED, "prune": PRUNE, "lanes": LANES}[option], value)
    eng.upload_positions(positions, charges)
    eng.start(first_stream=first_stream)
    return eng


def _coulomb_chains(n_chains, n, length, mixed):
    positions = workloads.uniform_start(n_chains, n, length)
    charges = np.ones((n_chains, n))
    if mixed:  # both signs: attractive pairs, the lower-bound Walker tables for negative active charges
        charges[:, 1::2] = -1.0
    return positions, charges


@pytest.mark.parametrize("lanes,mixed", [(4, False), (8, False), (4, True)])
def test_coulomb_batched_kernel_commits_the_events_of_the_single_event_kernel(lanes, mixed):
    """lj_spec_kernel<coulomb> against event_kernel<IPCB, MIC, MIC> on the same chains: every field of every event record,
    times and positions bit for bit, the same counters."""
    n_chains, n, events = 48, 64, 2500
    builder, length = _coulomb_program(n)
    positions, charges = _coulomb_chains(n_chains, n, length, mixed)
    with _coulomb_started(builder, positions, charges, 11, batched=0) as single, \
            _coulomb_started(builder, positions, charges, 11, lanes=lanes) as batched:
        assert "lj_spec_kernel<coulomb" in batched.kernel_name(record=True)
        assert "event_kernel" in single.kernel_name(record=True)
        ref, ref_stats = single.run_recorded(max_events=events, records_per_chain=events)
        rec, stats = batched.run_recorded(max_events=events, records_per_chain=events)
        assert stats == ref_stats
        assert stats["pair_events"] > 0 and stats["veto_accepted"] > 0 and stats["boundary_events"] > 0 and \
            stats["end_of_chain_events"] > 0 and stats["veto_events"] > 10 * stats["pair_events"]
        for field in rec.dtype.names:
            same = rec[field] == ref[field]
            if not np.all(same):
                chain, event = [int(v[0]) for v in np.nonzero(~same.reshape(n_chains, events, -1).all(axis=2))]
                raise AssertionError(f"{field}: chain {chain} event {event}: {rec[chain][event]} vs {ref[chain][event]}")
        _assert_same_state(single, batched, "after the recorded launch")


@pytest.mark.parametrize("mixed", [False, True])
def test_coulomb_pruning_does_not_change_the_chains(mixed):
    n_chains, n = 96, 64
    builder, length = _coulomb_program(n)
    positions, charges = _coulomb_chains(n_chains, n, length, mixed)
    with _coulomb_started(builder, positions, charges, 500, prune=0) as full, \
            _coulomb_started(builder, positions, charges, 500, prune=1) as pruned, \
            _coulomb_started(builder, positions, charges, 500, batched=0) as single:
        assert "lj_spec_kernel<coulomb" in pruned.kernel_name() and "prune=1" in pruned.kernel_name()
        totals = []
        for eng in (full, pruned, single):
            for events in (1, 2, 3, 5, 8, 13, 968, 4000):
                eng.run(max_events=events)
            totals.append(eng.sync())
        _assert_same_state(full, pruned, "pruned vs unpruned")
        _assert_same_state(full, single, "batched vs single-event kernel")
        assert totals[0] == totals[2]
        for key in totals[0]:
            if key != "candidates":
                assert totals[0][key] == totals[1][key], key
        assert totals[1]["candidates"] < totals[0]["candidates"]


def test_coulomb_time_limits_cut_batches_like_the_single_event_kernel():
    import copy
    n_chains, n = 48, 64
    # the module's program with a shorter chain time (a copy of the EcmcProgram; its tables stay those of the cached builder)
    cached, length = _coulomb_program(n)
    builder = copy.copy(cached)
    builder.program = type(cached.program).from_buffer_copy(bytes(cached.program))
    builder.program.chain_time = 0.3
    positions, charges = _coulomb_chains(n_chains, n, length, True)
    with _coulomb_started(builder, positions, charges, 40, batched=0) as single, \
            _coulomb_started(builder, positions, charges, 40, prune=1) as batched:
        for eng in (single, batched):
            # time limits every 1 / 128 up to 0.3203: one end of chain (0.3) falls inside; the horizon of 3.06 this test
            # started with cost 220 s of the GPU suite for the same code paths
            for k in range(1, 40):
                eng.run(until=(0.0, k / 128.0))
                eng.run(max_events=3)
            eng.run(until=(0.0, 0.3203125))
            eng.sync()
        _assert_same_state(single, batched, "after interleaved time limits")
        states = batched.chain_states()
        assert np.all(states["time_q"] == 0.0) and np.all(states["time_r"] == 0.3203125)
        assert np.all(states["eoc_r"] > 0.3203125) and np.all(states["event_counter"] > 100)
