"""The host layer: compile a JeLLyFysh object graph, built by the reference's own factory from INI text, into an
EcmcProgram. Needs the installed reference copy under baseline/_ref (git-ignored, travels to the GPU box); skipped
without it. No GPU needed: the compiled program is run by the CPU oracle and compared with the reference trace."""
import configparser
import contextlib
import io
import os
import sys

import numpy as np
import pytest

import configs
import trace_util as tu

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "jellyfysh", "run.py")),
                                reason="baseline/_ref (installed reference) not present")


def build_reference_graph(ini_text, positions=None, composites=None):
    """The reference's run.py:182-183 on INI text; returns (mediator, setting module). positions / composites: the
    start configuration fed to the reference's random input handler (configs.patch_composite_start)."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import warnings
    warnings.filterwarnings("ignore")
    from jellyfysh.base import factory
    from jellyfysh.base.strings import to_camel_case
    import jellyfysh.setting as setting
    from jellyfysh.activator.tagger.factor_type_maps import FactorTypeMaps
    setting.reset()
    FactorTypeMaps._instance = None
    factory.used_sections.clear()
    config = configparser.ConfigParser()
    config.read_string(ini_text)
    factory.build_from_config(config, to_camel_case(config.get("Run", "setting")), "jellyfysh.setting")
    if positions is not None:
        iterator = iter([list(map(float, p)) for p in positions])
        setting.random_position = lambda: next(iterator)
    restore = configs.patch_composite_start(composites) if composites is not None else (lambda: None)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            mediator = factory.build_from_config(config, to_camel_case(config.get("Run", "mediator")),
                                                 "jellyfysh.mediator")
    finally:
        restore()
    return mediator, setting


@pytest.fixture
def lj_graph():
    g = tu.load_trace("trace_lj_small")
    ini = configs.lennard_jones_ini(int(g["meta_n"]), float(g["meta_system_length"]), int(g["meta_cells_per_side"][0]),
                                    chain_time=float(g["meta_chain_time"]), points_per_side=int(g["meta_estimator"][1]),
                                    estimator_prefactor=float(g["meta_estimator"][0]), end_of_run_time=50.0)
    mediator, setting = build_reference_graph(ini, g["positions0"])
    yield g, mediator
    setting.reset()


def test_compiled_program_matches_reference_tables(lj_graph):
    from jellyfysh_b200 import abi, compiler
    g, mediator = lj_graph
    compiled = compiler.compile_program(mediator._activator, mediator._state_handler.extract_global_state(),
                                        seed=int(g["seed"][0]))
    p = compiled.builder.program
    assert (p.dimension, p.n_particles, p.neighbor_layers, p.max_occupants) == (3, int(g["meta_n"]), 1, 1)
    assert [p.cells_per_side[d] for d in range(3)] == [int(c) for c in g["meta_cells_per_side"]]
    assert (p.system_length, p.beta, p.chain_time, p.speed) == (float(g["meta_system_length"]), 1.0,
                                                                float(g["meta_chain_time"]), 1.0)
    assert p.pair_handler == abi.PAIR_TWO_LEAF_UNIT and p.pair_potential.kind == abi.POT_LENNARD_JONES
    assert list(p.pair_potential.params)[:2] == list(g["meta_lj"])
    assert p.veto_enabled == 1 and p.veto_use_charge == 0 and compiled.charge_name is None
    ref = tu.reference_tables(g)
    ours = compiled.builder.tables
    assert np.array_equal(ours["bounds"], np.nan_to_num(ref["bounds"], nan=0.0))
    for kind in ("upper", "lower"):
        for d in range(3):
            for key in ("cell_a", "cell_b", "rate_a"):
                assert np.array_equal(ours[kind][d][key], ref[kind][d][key])
            assert ours[kind][d]["total_rate"] == ref[kind][d]["total_rate"]
    assert len(compiled.control_handlers) == 1  # end of run


def test_compiled_program_replays_reference_trace(oracle, lj_graph):
    """INI -> reference factory -> compiler -> oracle chain reproduces the recorded reference run bit for bit."""
    from jellyfysh_b200 import compiler
    g, mediator = lj_graph
    template = mediator._state_handler.extract_global_state()
    compiled = compiler.compile_program(mediator._activator, template, seed=int(g["seed"][0]))
    positions, charges, roots = compiler.positions_and_charges(template, compiled.charge_name)
    assert np.array_equal(positions, g["positions0"]) and charges is None and roots is None
    chain = oracle.OracleChain(compiled.builder)
    chain.set_positions(positions)
    chain.start(stream=int(g["seed"][1]))
    n, rec = chain.run(max_events=2000, record=2000)
    assert n == 2000 and tu.records_equal_discrete(rec, g["records"][:2000])
    assert np.array_equal(rec["time_r"], g["records"]["time_r"][:2000])


def test_coulomb_graph_compiles_with_charges():
    from jellyfysh_b200 import abi, compiler
    g = tu.load_trace("trace_coulomb_small")
    ini = configs.coulomb_atoms_ini(int(g["meta_n"]), [int(c) for c in g["meta_cells_per_side"]],
                                    points_per_side=int(g["meta_estimator"][1]), end_of_run_time=5.0)
    mediator, setting = build_reference_graph(ini, g["positions0"])
    try:
        compiled = compiler.compile_program(mediator._activator, mediator._state_handler.extract_global_state())
        p = compiled.builder.program
        assert p.pair_handler == abi.PAIR_TWO_LEAF_UNIT_BOUNDING
        assert p.pair_potential.kind == abi.POT_MERGED_IMAGE_COULOMB
        assert p.pair_bounding_potential.kind == abi.POT_INVERSE_POWER_COULOMB_BOUNDING
        assert list(p.pair_potential.params)[:4] == list(g["meta_mic"])
        assert p.pair_use_charge == 1 and p.veto_use_charge == 1 and compiled.charge_name == "electric_charge"
        assert [p.cells_per_side[d] for d in range(3)] == [int(c) for c in g["meta_cells_per_side"]]
    finally:
        setting.reset()


def test_power_bounded_graph_without_cells_compiles_and_replays(oracle):
    """The shipped coulomb_atoms/power_bounded.ini (no cell system, pair factors from the factor type map) sized for
    six atoms -> compiler -> oracle chain reproduces the reference trace of the same configuration bit for bit."""
    from jellyfysh_b200 import abi, compiler
    g = tu.load_trace("trace_coulomb_power_bounded")
    mediator, setting = build_reference_graph(configs.coulomb_power_bounded_ini(REF, n_atoms=int(g["meta_n"]),
                                                                                end_of_run_time=5.0), g["positions0"])
    try:
        state = mediator._state_handler.extract_global_state()
        compiled = compiler.compile_program(mediator._activator, state, seed=int(g["seed"][0]))
        p = compiled.builder.program
        assert p.no_cells == 1 and p.veto_enabled == 0 and p.max_surplus >= p.n_particles - 1
        assert [p.cells_per_side[d] for d in range(3)] == [1, 1, 1] and p.neighbor_layers == 0
        assert p.pair_handler == abi.PAIR_TWO_LEAF_UNIT_BOUNDING and p.pair_use_charge == 1
        assert p.pair_potential.kind == abi.POT_MERGED_IMAGE_COULOMB
        assert p.pair_bounding_potential.kind == abi.POT_INVERSE_POWER_COULOMB_BOUNDING
        assert compiled.charge_name == "electric_charge" and p.chain_time == float(g["meta_chain_time"])
        positions, charges, _ = compiler.positions_and_charges(state, compiled.charge_name)
        assert np.array_equal(positions, g["positions0"])
        chain = oracle.OracleChain(compiled.builder)
        chain.set_positions(positions, charges)
        chain.start(stream=int(g["seed"][1]))
        records = g["records"][:1500]
        n, rec = chain.run(max_events=len(records), record=len(records))
        assert n == len(records) and tu.records_equal_discrete(rec, records)
        assert np.array_equal(rec["time_q"], records["time_q"]) and np.array_equal(rec["time_r"], records["time_r"])
    finally:
        setting.reset()


@pytest.mark.parametrize("lifting", ["inside_first", "outside_first", "ratio"])
def test_dipole_factors_graph_without_cells_compiles_and_replays(oracle, lifting):
    """The shipped dipoles/dipole_factors_*.ini (composite objects without a cell system) sized for three dipoles ->
    compiler -> oracle chain reproduces the reference trace of the same configuration bit for bit."""
    from jellyfysh_b200 import abi, compiler
    g = tu.load_trace("trace_dipole_factors_" + lifting)
    n = int(g["meta_n"]) // 2
    ini = configs.shipped_without_sampling(
        REF, ("2018_JCP_149_064113", "dipoles", f"dipole_factors_{lifting}.ini"), end_of_run_time=5.0,
        replacements=[("number_of_root_nodes = 2", f"number_of_root_nodes = {n}"),
                      ("number_event_handlers = 1", f"number_event_handlers = {2 * (n - 1)}")])
    mediator, setting = build_reference_graph(ini, composites=(g["roots0"], g["positions0"].reshape(n, 2, 3)))
    try:
        state = mediator._state_handler.extract_global_state()
        compiled = compiler.compile_program(mediator._activator, state, seed=int(g["seed"][0]))
        p = compiled.builder.program
        assert p.no_cells == 1 and p.cell_level == 1 and p.nodes_per_root == 2 and p.veto_enabled == 0
        assert p.pair_handler == abi.PAIR_TWO_COMPOSITE_SUMMED_BOUNDING and p.pair_use_charge == 1
        assert p.composite_lifting == {"inside_first": abi.LIFTING_INSIDE_FIRST, "outside_first": abi.LIFTING_OUTSIDE_FIRST,
                                       "ratio": abi.LIFTING_RATIO}[lifting]
        assert p.n_bonds == 1 and p.n_inter_factors == 2
        assert sorted((p.inter_factors[i][0], p.inter_factors[i][1]) for i in range(2)) == [(0, 1), (1, 0)]
        positions, charges, roots = compiler.positions_and_charges(state, compiled.charge_name)
        assert np.array_equal(positions, g["positions0"]) and np.array_equal(roots, g["roots0"])
        chain = oracle.OracleChain(compiled.builder)
        chain.set_positions(positions, charges)
        chain.set_roots(roots)
        chain.start(stream=int(g["seed"][1]))
        records = g["records"][:1500]
        n_done, rec = chain.run(max_events=len(records), record=len(records))
        assert n_done == len(records) and tu.records_equal_discrete(rec, records)
        assert np.array_equal(rec["time_q"], records["time_q"]) and np.array_equal(rec["time_r"], records["time_r"])
    finally:
        setting.reset()


def test_dipole_motion_graph_compiles_and_replays(oracle):
    """The shipped dipoles/dipole_motion.ini (root-unit-active handlers + RootLeafUnitActiveSwitcher) sized for three
    dipoles -> compiler -> oracle chain reproduces the reference trace of the same configuration bit for bit."""
    from jellyfysh.base.exceptions import ConfigurationError
    from jellyfysh_b200 import abi, compiler
    g = tu.load_trace("trace_dipole_motion")
    n = int(g["meta_n"]) // 2
    ini = configs.shipped_without_sampling(
        REF, ("2018_JCP_149_064113", "dipoles", "dipole_motion.ini"), end_of_run_time=5.0,
        replacements=[("number_of_root_nodes = 2", f"number_of_root_nodes = {n}"),
                      ("number_event_handlers = 2", f"number_event_handlers = {2 * (n - 1)}"),
                      ("number_event_handlers = 1", f"number_event_handlers = {2 * (n - 1)}")])
    mediator, setting = build_reference_graph(ini, composites=(g["roots0"], g["positions0"].reshape(n, 2, 3)))
    try:
        state = mediator._state_handler.extract_global_state()
        compiled = compiler.compile_program(mediator._activator, state, seed=int(g["seed"][0]))
        p = compiled.builder.program
        assert p.no_cells == 1 and p.cell_level == 1 and p.nodes_per_root == 2 and p.veto_enabled == 0
        assert p.pair_handler == abi.PAIR_TWO_COMPOSITE_SUMMED_BOUNDING and p.pair_use_charge == 1
        assert p.composite_lifting == abi.LIFTING_INSIDE_FIRST and p.n_bonds == 1 and p.n_inter_factors == 2
        assert p.root_mode == 1 and (p.switch_chain_length[0], p.switch_chain_length[1]) == (0.69, 0.7)
        positions, charges, roots = compiler.positions_and_charges(state, compiled.charge_name)
        chain = oracle.OracleChain(compiled.builder)
        chain.set_positions(positions, charges)
        chain.set_roots(roots)
        chain.start(stream=int(g["seed"][1]))
        records = g["records"][:2000]
        n_done, rec = chain.run(max_events=len(records), record=len(records))
        assert n_done == len(records) and tu.records_equal_discrete(rec, records)
        assert np.array_equal(rec["mode"], records["reserved"]) and (rec["kind"] == abi.EVENT_SWITCH).sum() > 50
        assert np.array_equal(rec["time_q"], records["time_q"]) and np.array_equal(rec["time_r"], records["time_r"])
        # a root-unit-active handler with another potential than the leaf-unit-active one is refused
        root_handler = [h for h in mediator._activator.get_event_handlers()
                        if "RootUnitActiveTwoLeafUnitEventHandler" in {c.__name__ for c in type(h).__mro__}][0]
        root_handler._potential = type(root_handler._potential)(prefactor=2.0e-6, power=6)
        with pytest.raises(ConfigurationError, match="root-unit-active"):
            compiler.compile_program(mediator._activator, state, seed=1)
    finally:
        setting.reset()


def test_leaf_cell_water_graph_compiles_and_replays(oracle):
    """The shipped water/coulomb_power_bounded_lj_cell_bounded.ini (oxygen-only cell system, piecewise-constant-bound and
    cell-bounding Lennard-Jones handlers) sized for twelve molecules -> compiler -> oracle chain reproduces the reference
    trace of the same configuration bit for bit."""
    from jellyfysh.base.exceptions import ConfigurationError
    from jellyfysh_b200 import abi, compiler
    g = tu.load_trace("trace_water_lj_cell_bounded")
    n = int(g["meta_n"]) // 3
    ini = configs.shipped_without_sampling(
        REF, ("2018_JCP_149_064113", "water", "coulomb_power_bounded_lj_cell_bounded.ini"), end_of_run_time=5.0,
        replacements=[("number_of_root_nodes = 2", f"number_of_root_nodes = {n}"),
                      ("number_event_handlers = 1", f"number_event_handlers = {n - 1}")])
    mediator, setting = build_reference_graph(ini, composites=(g["roots0"], g["positions0"].reshape(n, 3, 3)))
    try:
        state = mediator._state_handler.extract_global_state()
        compiled = compiler.compile_program(mediator._activator, state, seed=int(g["seed"][0]))
        p = compiled.builder.program
        assert p.no_cells == 0 and p.cell_level == 1 and p.nodes_per_root == 3 and p.cell_child == 2
        assert [p.cells_per_side[d] for d in range(3)] == [6, 6, 6] and p.neighbor_layers == 1 and p.max_occupants == 1
        assert p.pair_handler == abi.PAIR_TWO_COMPOSITE_SUMMED_BOUNDING and p.veto_enabled == abi.FAR_CELL_BOUNDING
        assert p.n_inter_factors == 1 and (p.inter_factors[0][0], p.inter_factors[0][1]) == (1, 1)
        assert (p.inter_bound_offset, p.inter_bound_max_displacement) == (10.0, 0.24353253124)
        assert p.boundary_keeps_factors == 1 and p.bending_enabled == 1 and p.n_bonds == 2
        assert np.array_equal(compiled.builder.tables["bounds"][..., 0], np.nan_to_num(g["bounds"][..., 0]))
        positions, charges, roots = compiler.positions_and_charges(state, compiled.charge_name)
        chain = oracle.OracleChain(compiled.builder)
        chain.set_positions(positions, charges)
        chain.set_roots(roots)
        chain.start(stream=int(g["seed"][1]))
        records = g["records"][:2000]
        n_done, rec = chain.run(max_events=len(records), record=len(records))
        assert n_done == len(records) and tu.records_equal_discrete(rec, records)
        assert np.array_equal(rec["time_q"], records["time_q"]) and np.array_equal(rec["time_r"], records["time_r"])
        # a cell system that stores two kinds of leaves is refused
        occupancy = mediator._activator._internal_states[0]
        occupancy._is_relevant_unit = lambda unit: unit.charge["electric_charge"] > 0
        with pytest.raises(ConfigurationError, match="charge filter"):
            compiler.compile_program(mediator._activator, state, seed=1)
    finally:
        setting.reset()


def test_atom_factors_graph_compiles_and_replays(oracle):
    """The shipped dipoles/atom_factors.ini (Coulomb as bounded leaf-to-leaf factors between the dipoles, no cell system)
    sized for three dipoles -> compiler -> oracle chain reproduces the reference trace bit for bit."""
    from jellyfysh_b200 import abi, compiler
    g = tu.load_trace("trace_dipole_atom_factors")
    n = int(g["meta_n"]) // 2
    ini = configs.shipped_without_sampling(
        REF, ("2018_JCP_149_064113", "dipoles", "atom_factors.ini"), end_of_run_time=5.0,
        replacements=[("number_of_root_nodes = 2", f"number_of_root_nodes = {n}"),
                      ("number_event_handlers = 2", f"number_event_handlers = {2 * (n - 1)}"),
                      ("number_event_handlers = 1", f"number_event_handlers = {2 * (n - 1)}")])
    mediator, setting = build_reference_graph(ini, composites=(g["roots0"], g["positions0"].reshape(n, 2, 3)))
    try:
        state = mediator._state_handler.extract_global_state()
        compiled = compiler.compile_program(mediator._activator, state, seed=int(g["seed"][0]))
        p = compiled.builder.program
        assert p.no_cells == 1 and p.cell_level == 1 and p.nodes_per_root == 2
        assert p.pair_handler == abi.PAIR_TWO_LEAF_UNIT_BOUNDING and p.pair_use_charge == 1
        assert p.n_bonds == 1 and p.n_inter_factors == 2
        positions, charges, roots = compiler.positions_and_charges(state, compiled.charge_name)
        chain = oracle.OracleChain(compiled.builder)
        chain.set_positions(positions, charges)
        chain.set_roots(roots)
        chain.start(stream=int(g["seed"][1]))
        records = g["records"][:1500]
        n_done, rec = chain.run(max_events=len(records), record=len(records))
        assert n_done == len(records) and tu.records_equal_discrete(rec, records)
        assert np.array_equal(rec["time_q"], records["time_q"]) and np.array_equal(rec["time_r"], records["time_r"])
    finally:
        setting.reset()


def test_dipole_cell_bounded_graph_compiles(oracle):
    """The shipped dipoles/cell_bounded.ini (TwoCompositeObjectCellBoundingPotentialEventHandler for the far field) sized
    for four dipoles -> compiler -> a program the oracle accepts and runs. (The bounds come from the reference's Monte
    Carlo estimator, which draws from the unseeded global random module here: the replay against the reference trace, with
    the bounds of the recorded run, is tests/test_oracle_traces.py::test_composite_cell_bounding_chain_replay_bit_exact.)"""
    from jellyfysh_b200 import abi, compiler
    g = tu.load_trace("trace_dipole_cell_bounded")
    n = int(g["meta_n"]) // 2
    ini = configs.shipped_without_sampling(
        REF, ("2018_JCP_149_064113", "dipoles", "cell_bounded.ini"), end_of_run_time=5.0,
        replacements=[("number_of_root_nodes = 2", f"number_of_root_nodes = {n}"),
                      ("number_event_handlers = 1", f"number_event_handlers = {2 * (n - 1)}"),
                      ("number_trials = 1000", "number_trials = 100")])
    mediator, setting = build_reference_graph(ini, composites=(g["roots0"], g["positions0"].reshape(n, 2, 3)))
    try:
        state = mediator._state_handler.extract_global_state()
        compiled = compiler.compile_program(mediator._activator, state, seed=int(g["seed"][0]))
        p = compiled.builder.program
        assert p.veto_enabled == abi.FAR_CELL_BOUNDING and p.cell_level == 1 and p.no_cells == 0
        assert p.pair_handler == abi.PAIR_TWO_COMPOSITE_SUMMED_BOUNDING and p.composite_lifting == abi.LIFTING_INSIDE_FIRST
        assert [p.cells_per_side[d] for d in range(3)] == [3, 5, 7] and p.boundary_keeps_factors == 1
        assert compiled.builder.tables["bounds"].shape == (105, 3, 2)
        positions, charges, roots = compiler.positions_and_charges(state, compiled.charge_name)
        chain = oracle.OracleChain(compiled.builder)
        chain.set_positions(positions, charges)
        chain.set_roots(roots)
        chain.start(stream=int(g["seed"][1]))
        n_done, rec = chain.run(max_events=500, record=500)
        assert n_done == 500 and (rec["kind"] == 5).sum() > 100 and chain.stats()["capacity_errors"] == 0
    finally:
        setting.reset()


def test_single_hard_disk_dipole_graph_compiles(oracle):
    """single_hard_disk_dipole.ini: one tethered pair of disks, no sphere factors, the velocity rotated by 23 degrees at
    every end of chain (general velocities) -> compiler -> a disk program; the oracle runs it (tether events and ends of
    chain only) and the velocity stays a unit vector that is not along an axis."""
    from jellyfysh_b200 import abi, compiler
    import math
    ini = configs.shipped_without_sampling(REF, ("hard_disk_dipoles", "single_hard_disk_dipole.ini"), end_of_run_time=5.0)
    mediator, setting = build_reference_graph(ini)
    try:
        template = mediator._state_handler.extract_global_state()
        compiled = compiler.compile_program(mediator._activator, template, seed=3)
        p = compiled.builder.program
        assert (p.dimension, p.n_particles, p.nodes_per_root, p.no_cells, p.eoc_sequential) == (2, 2, 2, 1, 1)
        assert p.n_bonds == 1 and p.bond_potential.kind == abi.POT_HARD_DIPOLE and p.n_inter_factors == 0
        assert p.pair_handler == abi.PAIR_NONE and p.chain_time == 0.5
        assert (p.eoc_cos, p.eoc_sin) == (math.cos(23.0 * math.pi / 180.0), math.sin(23.0 * math.pi / 180.0))
        positions, charges, roots = compiler.positions_and_charges(template, compiled.charge_name)
        chain = oracle.OracleChain(compiled.builder)
        chain.set_positions(positions)
        chain.set_roots(roots)
        chain.start(stream=1)
        n, rec = chain.run(max_events=400, record=400)
        stats = chain.stats()
        assert n == 400 and stats["bond_events"] > 100 and stats["end_of_chain_events"] > 50
        assert stats["bond_events"] + stats["end_of_chain_events"] == 400
        velocity = chain.state().velocity[:]
        assert abs(math.hypot(*velocity) - 1.0) < 1e-13 and all(abs(v) > 1e-3 for v in velocity)
    finally:
        setting.reset()


def test_sequential_direction_graph_compiles_and_replays(oracle):
    """The shipped hard_disk_dipoles.ini (81 dipoles from the shipped PDB file, no cell system, sequential-direction end
    of chain) -> compiler -> oracle chain reproduces the reference's recorded run bit for bit."""
    from jellyfysh_b200 import abi, compiler
    import jellyfysh_b200
    jellyfysh_b200.install()  # MDAnalysis stand-in for the PdbInputHandler where MDAnalysis is not installed
    g = tu.load_trace("trace_hard_disk_dipoles_sequential")
    mediator, setting = build_reference_graph(configs.hard_disk_dipoles_ini(REF, end_of_run_time=50.0,
                                                                            chain_time=float(g["meta_chain_time"])))
    try:
        template = mediator._state_handler.extract_global_state()
        compiled = compiler.compile_program(mediator._activator, template, seed=int(g["seed"][0]))
        p = compiled.builder.program
        assert (p.dimension, p.n_particles, p.nodes_per_root, p.no_cells, p.eoc_sequential) == (2, 162, 2, 1, 1)
        assert p.n_bonds == 1 and p.bond_potential.kind == abi.POT_HARD_DIPOLE
        assert p.n_inter_factors == 4 and p.inter_potential.kind == abi.POT_HARD_SPHERE
        assert sorted((p.inter_factors[i][0], p.inter_factors[i][1]) for i in range(4)) == [(0, 0), (0, 1), (1, 0), (1, 1)]
        reference = tu.sequential_dipole_builder_of(g, oracle.ProgramBuilder).program
        assert (p.eoc_cos, p.eoc_sin, p.chain_time) == (reference.eoc_cos, reference.eoc_sin, reference.chain_time)
        positions, charges, roots = compiler.positions_and_charges(template, compiled.charge_name)
        assert np.array_equal(positions, g["positions0"]) and np.array_equal(roots, g["roots0"])
        chain = oracle.OracleChain(compiled.builder)
        chain.set_positions(positions)
        chain.set_roots(roots)
        chain.start(stream=int(g["seed"][1]))
        n, rec = chain.run(max_events=3000, record=3000)
        assert n == 3000 and tu.records_equal_discrete(rec, g["records"][:3000])
        assert np.array_equal(rec["time_r"], g["records"]["time_r"][:3000])
    finally:
        setting.reset()


def test_hard_disk_dipole_graph_compiles_and_replays(oracle):
    """C1: the shipped hard_disk_dipoles_cells.ini (composite point objects, leaf-level cells, unbounded occupancy,
    hard-sphere pairs + hard-dipole tether from the factor type map) -> compiler -> oracle chain reproduces the
    reference's recorded run bit for bit."""
    from jellyfysh_b200 import abi, compiler
    g = tu.load_trace("trace_hard_disk_dipoles")
    n_roots = len(g["roots0"])
    composites = (g["roots0"], g["positions0"].reshape(n_roots, 2, -1))
    mediator, setting = build_reference_graph(configs.hard_disk_dipoles_cells_ini(REF, end_of_run_time=50.0),
                                              composites=composites)
    try:
        template = mediator._state_handler.extract_global_state()
        compiled = compiler.compile_program(mediator._activator, template, seed=int(g["seed"][0]), occupant_capacity=6)
        p = compiled.builder.program
        assert (p.dimension, p.n_particles, p.nodes_per_root, p.n_bonds) == (2, 162, 2, 1)
        assert (p.bonds[0][0], p.bonds[0][1]) == (0, 1) and p.bond_potential.kind == abi.POT_HARD_DIPOLE
        assert p.pair_handler == abi.PAIR_TWO_LEAF_UNIT and p.pair_potential.kind == abi.POT_HARD_SPHERE
        assert p.veto_enabled == 0 and p.max_occupants == 6 and p.max_surplus == 0 and p.initial_active == 0
        # the lengths recovered from the stored squares reproduce them exactly
        assert 4.0 * p.pair_potential.params[0] * p.pair_potential.params[0] == 4.0 * 0.476190476190476 * 0.476190476190476
        assert p.bond_potential.params[0] * p.bond_potential.params[0] == 0.952380952380952 * 0.952380952380952
        positions, charges, roots = compiler.positions_and_charges(template, compiled.charge_name)
        assert np.array_equal(positions, g["positions0"]) and np.array_equal(roots, g["roots0"]) and charges is None
        chain = oracle.OracleChain(compiled.builder)
        chain.set_positions(positions)
        chain.set_roots(roots)
        chain.start(stream=int(g["seed"][1]))
        n, rec = chain.run(max_events=3000, record=3000)
        assert n == 3000 and tu.records_equal_discrete(rec, g["records"][:3000])
        assert np.array_equal(rec["time_r"], g["records"]["time_r"][:3000])
    finally:
        setting.reset()


def test_water_graph_compiles_and_replays(oracle):
    """C4: the shipped water/coulomb_cell_veto_lj_inverted.ini -> compiler -> oracle chain reproduces the reference's
    recorded run bit for bit (the cell-veto tables are those the reference built, from the fixture: the dipole Monte
    Carlo estimator draws random numbers)."""
    from jellyfysh_b200 import abi, compiler
    g = tu.load_trace("trace_water")
    n_roots = len(g["roots0"])
    composites = (g["roots0"], g["positions0"].reshape(n_roots, 3, 3))
    mediator, setting = build_reference_graph(configs.water_ini(REF, n_molecules=n_roots, number_trials=2,
                                                                end_of_run_time=50.0), composites=composites)
    try:
        template = mediator._state_handler.extract_global_state()
        compiled = compiler.compile_program(mediator._activator, template, seed=int(g["seed"][0]))
        p = compiled.builder.program
        assert (p.dimension, p.n_particles, p.nodes_per_root, p.cell_level, p.neighbor_layers) == (3, 96, 3, 1, 2)
        assert p.pair_handler == abi.PAIR_TWO_COMPOSITE_SUMMED_BOUNDING and p.composite_lifting == abi.LIFTING_INSIDE_FIRST
        assert p.pair_potential.kind == abi.POT_MERGED_IMAGE_COULOMB and p.pair_potential.params[0] == 332.0
        assert p.pair_bounding_potential.kind == abi.POT_INVERSE_POWER_COULOMB_BOUNDING
        assert p.veto_enabled == abi.FAR_CELL_VETO and p.veto_use_charge == 1 and compiled.charge_name == "electric_charge"
        assert p.n_bonds == 2 and sorted((p.bonds[i][0], p.bonds[i][1]) for i in range(2)) == [(0, 1), (1, 2)]
        assert p.bond_potential.kind == abi.POT_DISPLACED_EVEN_POWER
        assert p.n_inter_factors == 1 and (p.inter_factors[0][0], p.inter_factors[0][1]) == (1, 1)
        assert p.inter_potential.kind == abi.POT_LENNARD_JONES
        assert p.bending_enabled == 1 and p.bending_lifting == abi.LIFTING_RATIO
        assert list(p.bending_children) == [0, 1, 2] and list(p.bending_separations) == [1, 0, 1, 2]
        assert (p.bending_offset, p.bending_max_displacement) == (10.0, 0.112321434)
        assert p.boundary_keeps_factors == 1 and p.initial_active == 1
        positions, charges, roots = compiler.positions_and_charges(template, compiled.charge_name)
        assert np.array_equal(positions, g["positions0"]) and np.array_equal(roots, g["roots0"])
        assert np.array_equal(charges, g["charges"])
        # the fixture's tables (1000 -> 200 trials there, 2 here) replace the freshly estimated ones
        compiled.builder.set_veto(compiled.builder.program.veto_potential, tu.reference_tables(g), use_charge=True,
                                  target_charge=1.0)
        chain = oracle.OracleChain(compiled.builder)
        chain.set_positions(positions, charges)
        chain.set_roots(roots)
        chain.start(stream=int(g["seed"][1]))
        n, rec = chain.run(max_events=2000, record=2000)
        assert n == 2000 and tu.records_equal_discrete(rec, g["records"][:2000])
        assert np.array_equal(rec["time_r"], g["records"]["time_r"][:2000])
    finally:
        setting.reset()


def test_unsupported_graphs_are_rejected(lj_graph):
    """A graph the device cannot run faithfully is refused, never approximated."""
    from jellyfysh.base.exceptions import ConfigurationError
    from jellyfysh_b200 import compiler
    g, mediator = lj_graph
    occupancy = mediator._activator._internal_states[0]
    occupancy._cell_level = 2
    with pytest.raises(ConfigurationError):
        compiler.compile_program(mediator._activator, mediator._state_handler.extract_global_state())
    occupancy._cell_level = 1
    compiler.compile_program(mediator._activator, mediator._state_handler.extract_global_state())
    # a cell occupancy that stores only units with a non-zero charge (single_active_cell_occupancy.py:92): the device
    # would bin the neutral units too, so the configuration is refused
    occupancy._is_relevant_unit = lambda unit: unit.charge["electric_charge"] != 0
    with pytest.raises(ConfigurationError, match="charge filter"):
        compiler.compile_program(mediator._activator, mediator._state_handler.extract_global_state())


def test_shipped_pdb_start_configuration_through_the_reference_input_handler():
    """config_files/hard_disk_dipoles/hard_disk_dipoles_cells.ini UNCHANGED (with its `pdb_input_handler`): where
    MDAnalysis is not installed, jellyfysh_b200.install() provides the .pdb reader the reference's PdbInputHandler needs
    (jellyfysh_b200/shims/MDAnalysis). The start configuration the reference builds from it -- coordinates rounded through
    float32 as MDAnalysis stores them, pdb_input_handler.py:165, barycentres over shortest separations :172-186 -- is the
    one the C1 trace fixture starts from, and the graph compiles to the C1 device program."""
    import jellyfysh_b200
    from jellyfysh_b200 import compiler
    if REF not in sys.path:
        sys.path.insert(0, REF)
    jellyfysh_b200.install()
    import MDAnalysis
    text = configs.shipped_ini(REF, "hard_disk_dipoles", "hard_disk_dipoles_cells.ini")
    text = text.replace("filename = config_files/", "filename = " + os.path.join(REF, "jellyfysh", "config_files") + "/")
    text = text.replace("output/hard_disk_dipoles/", "/tmp/jf_b200_pdb_test_")
    assert "input_handler = pdb_input_handler" in text
    mediator, setting = build_reference_graph(text)
    try:
        state = mediator._state_handler.extract_global_state()
        roots = np.array([node.value.position for node in state])
        leaves = np.array([[child.value.position for child in node.children] for node in state])
        compiled = compiler.compile_program(mediator._activator, state, seed=1)
    finally:
        setting.reset()
    expected_roots, expected_leaves = configs.read_pdb_dipoles(REF)
    assert np.array_equal(leaves, expected_leaves) and np.array_equal(roots, expected_roots)
    g = tu.load_trace("trace_hard_disk_dipoles")
    assert np.array_equal(leaves.reshape(-1, 2), g["positions0"]) and np.array_equal(roots, g["roots0"])
    assert compiled.nodes_per_root == 2 and compiled.n_particles == 162
    if "shim" in getattr(MDAnalysis, "__version__", ""):
        with pytest.raises(NotImplementedError):
            MDAnalysis.Writer("x.pdb", 3)
