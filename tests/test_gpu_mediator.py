"""The drop-in: an INI file of the reference grammar with `[Run] mediator = cuda_batched_mediator` runs through the
reference's own factory, input handler, state handler and output handlers, with the event loop on the GPU.
Needs the installed reference copy (baseline/_ref) next to a CUDA device."""
import os

import numpy as np
import pytest

import configs
import trace_util as tu
from test_compiler import REF, build_reference_graph

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(os.path.join(REF, "jellyfysh", "run.py")),
                                 reason="baseline/_ref (installed reference) not present")]


def _device_ini(g, tmp_path, chains, end_of_run_time, sampling_interval):
    ini = configs.lennard_jones_ini(int(g["meta_n"]), float(g["meta_system_length"]), int(g["meta_cells_per_side"][0]),
                                    chain_time=float(g["meta_chain_time"]), points_per_side=int(g["meta_estimator"][1]),
                                    estimator_prefactor=float(g["meta_estimator"][0]), end_of_run_time=end_of_run_time,
                                    sampling_interval=sampling_interval)
    ini = ini.replace("mediator = single_process_mediator", "mediator = cuda_batched_mediator")
    ini = ini.replace("[SingleProcessMediator]", "[CudaBatchedMediator]\nnumber_of_chains = %d\nseed = %d\n"
                                                 "first_random_stream = %d" % (chains, int(g["seed"][0]), int(g["seed"][1])))
    return ini.replace("/tmp/jf_b200_golden_separation.dat", str(tmp_path / "separation.dat"))


def test_ini_runs_on_the_device(oracle, tmp_path):
    import sys
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import jellyfysh_b200
    from jellyfysh_b200 import engine
    from jellyfysh_b200.program import ProgramBuilder
    jellyfysh_b200.install()
    from jellyfysh.base.exceptions import EndOfRun

    g = tu.load_trace("trace_lj_small")
    end, interval = 2.5, 0.4
    mediator, setting = build_reference_graph(_device_ini(g, tmp_path, 1, end, interval), g["positions0"])
    try:
        assert type(mediator).__mro__[1].__name__ == "Mediator" or "CudaBatchedMediator" in type(mediator).__name__
        with pytest.raises(EndOfRun):
            mediator.run()
        mediator.post_run()
        final = np.array([node.value.position for node in mediator._state_handler.extract_global_state()])
        stats = mediator.statistics
    finally:
        setting.reset()
    assert stats["events"] > 1000 and stats["capacity_errors"] == 0
    # the same program driven directly, interrupted at the same control times
    times = [oracle.time_from_float(interval * k) for k in range(1, int(end / interval) + 1)] + [oracle.time_from_float(end)]
    with engine.Engine(tu.builder_of(g, ProgramBuilder), n_chains=1) as eng:
        eng.upload_positions(g["positions0"][None])
        eng.start(first_stream=int(g["seed"][1]))
        events = 0
        for t in times:
            eng.run(until=t)
            events += eng.sync()["events"]
        direct = eng.download_positions()[0]
    assert events == stats["events"]
    assert np.array_equal(final, direct)
    # and the oracle agrees with both
    chain = oracle.OracleChain(tu.builder_of(g, oracle.ProgramBuilder))
    chain.set_positions(g["positions0"])
    chain.start(stream=int(g["seed"][1]))
    n_oracle = sum(chain.run(until=t)[0] for t in times)
    assert n_oracle == events
    assert np.max(np.abs(chain.positions() - final)) < 1e-12 * float(g["meta_system_length"])
    # the reference's own output handler wrote the samples of every sampling event
    lines = [line for line in (tmp_path / "separation.dat").read_text().strip().splitlines() if not line.startswith("#")]
    n = int(g["meta_n"])
    assert len(lines) == int(end / interval) * n * (n - 1) // 2


def test_many_chains_from_the_input_handler(tmp_path):
    """number_of_chains > 1: further start configurations come from the reference's input handler."""
    import sys
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import jellyfysh_b200
    jellyfysh_b200.install()
    from jellyfysh.base.exceptions import EndOfRun
    g = tu.load_trace("trace_lj_small")
    rng = np.random.default_rng(3)
    side = 4
    grid = np.stack(np.meshgrid(*[np.arange(side)] * 3, indexing="ij"), axis=-1).reshape(-1, 3)[:int(g["meta_n"])]
    length = float(g["meta_system_length"])
    chains = 6
    positions = np.concatenate([(grid + 0.5) * (length / side) + rng.uniform(-0.1, 0.1, size=grid.shape)
                                for _ in range(chains)])
    mediator, setting = build_reference_graph(_device_ini(g, tmp_path, chains, 1.0, None), positions)
    try:
        with pytest.raises(EndOfRun):
            mediator.run()
        stats = mediator.statistics
        states = mediator.engine.chain_states()
    finally:
        setting.reset()
    assert stats["events"] > chains * 100
    assert np.all(states["time_q"] == 1.0) and np.all(states["time_r"] == 0.0)
    assert len(set(states["event_counter"].tolist())) > 1  # the chains are different


def test_several_engines_and_device_observables(tmp_path):
    """`devices = 0, 0, 0` (three engines, here on one device; chain blocks 3 + 3 + 2) and `device_observables`: the chains
    are those of a single engine, chain for chain, and the separation histogram accumulated on the device is the histogram
    of the lines the reference's SeparationOutputHandler prints for the same run."""
    import sys
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import jellyfysh_b200
    jellyfysh_b200.install()
    from jellyfysh.base.exceptions import EndOfRun
    g = tu.load_trace("trace_lj_small")
    n, length, chains, end, interval, bins = int(g["meta_n"]), float(g["meta_system_length"]), 8, 1.7, 0.4, 50
    rng = np.random.default_rng(3)
    grid = np.stack(np.meshgrid(*[np.arange(4)] * 3, indexing="ij"), axis=-1).reshape(-1, 3)[:n]
    positions = np.concatenate([(grid + 0.5) * (length / 4) + rng.uniform(-0.1, 0.1, size=grid.shape) for _ in range(chains)])
    results = {}
    for tag, extra in (("one", ""), ("three", "\ndevices = 0, 0, 0\ndevice_observables = true\nhistogram_bins = %d" % bins)):
        folder = tmp_path / tag
        folder.mkdir()
        ini = _device_ini(g, folder, chains, end, interval).replace("first_random_stream", "first_random_stream")
        ini = ini.replace("[CudaBatchedMediator]", "[CudaBatchedMediator]" + extra)
        mediator, setting = build_reference_graph(ini, positions)
        try:
            with pytest.raises(EndOfRun):
                mediator.run()
            mediator.post_run()
            results[tag] = (mediator.chain_states(), np.concatenate([e.download_positions() for e in mediator.engines]),
                            mediator.statistics, dict(mediator.observables), len(mediator.engines))
        finally:
            setting.reset()
    one, three = results["one"], results["three"]
    assert one[4] == 1 and three[4] == 3
    assert np.array_equal(one[0], three[0]) and np.array_equal(one[1], three[1])
    assert {k: v for k, v in one[2].items()} == {k: v for k, v in three[2].items()}
    # the reference's output handler printed every separation of every sample of the first run
    lines = np.loadtxt(tmp_path / "one" / "separation.dat", comments="#")
    samples = int(end / interval)
    assert len(lines) == samples * chains * n * (n - 1) // 2
    observable = three[3]["separation_output_handler"]
    assert observable["samples"] == samples and len(observable["counts"]) == bins
    expected, _ = np.histogram(lines, bins=observable["edges"])
    assert int(observable["counts"].sum()) == len(lines)
    # (a separation within rounding distance of a bin edge may fall on the other side of it)
    assert np.abs(observable["counts"].astype(np.int64) - expected).sum() <= 4
    with np.load(str(tmp_path / "three" / "separation.dat") + ".histogram.npz") as written:
        assert np.array_equal(written["counts"], observable["counts"]) and int(written["samples"]) == samples


def test_cuda_state_handler_reads_single_chains(tmp_path):
    """`state_handler = cuda_state_handler`: the reference's state-handler contract (state_handler.py:63-165) over the
    chains of the device engines -- select_chain(c) downloads chain c alone (ecmc_download_chain) and the four contract
    methods then describe it: positions of all units, velocity and time stamp of the active unit only."""
    import sys
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import jellyfysh_b200
    jellyfysh_b200.install()
    from jellyfysh.base.exceptions import EndOfRun
    g = tu.load_trace("trace_lj_small")
    n, length, chains = int(g["meta_n"]), float(g["meta_system_length"]), 5
    rng = np.random.default_rng(8)
    grid = np.stack(np.meshgrid(*[np.arange(4)] * 3, indexing="ij"), axis=-1).reshape(-1, 3)[:n]
    positions = np.concatenate([(grid + 0.5) * (length / 4) + rng.uniform(-0.1, 0.1, size=grid.shape) for _ in range(chains)])
    ini = _device_ini(g, tmp_path, chains, 1.3, None).replace("state_handler = tree_state_handler",
                                                              "state_handler = cuda_state_handler")
    ini = ini.replace("[TreeStateHandler]", "[CudaStateHandler]").replace("[CudaBatchedMediator]",
                                                                          "[CudaBatchedMediator]\ndevices = 0, 0")
    assert "cuda_state_handler" in ini and "[CudaStateHandler]" in ini
    mediator, setting = build_reference_graph(ini, positions)
    try:
        with pytest.raises(EndOfRun):
            mediator.run()
        handler = mediator._state_handler
        assert type(handler).__mro__[0].__name__.startswith("CudaStateHandler") and handler.number_of_chains == chains
        everything = np.concatenate([e.download_positions() for e in mediator.engines])
        states = mediator.chain_states()
        for chain in (3, 0, 4):
            handler.select_chain(chain)
            assert handler.selected_chain == chain
            state = handler.extract_global_state()
            assert np.array_equal(np.array([node.value.position for node in state]), everything[chain])
            moving = [node.value for node in state if node.value.velocity is not None]
            assert len(moving) == 1 and moving[0].identifier == (int(states[chain]["active"]),)
            assert moving[0].velocity[int(states[chain]["direction"])] == 1.0
            assert (moving[0].time_stamp.quotient, moving[0].time_stamp.remainder) == \
                (float(states[chain]["time_q"]), float(states[chain]["time_r"]))
            active = handler.extract_active_global_state()
            assert len(active) == 1 and active[0].value.identifier == moving[0].identifier
        with pytest.raises(IndexError):
            handler.select_chain(chains)
    finally:
        setting.reset()


@pytest.mark.parametrize("config,output", [("cell_veto.ini", "SamplesOfSeparation_CellVeto.dat"),
                                           ("cell_bounded.ini", "SamplesOfSeparation_CellBounded.dat"),
                                           ("power_bounded.ini", "SamplesOfSeparation_PowerBounded.dat")])
def test_shipped_coulomb_atoms_config_matches_reference_statistics(tmp_path, config, output):
    """The statistical check of the reference (README.md:169-189): the shipped coulomb_atoms/cell_veto.ini (far field
    through the cell-veto handler), cell_bounded.ini (through the cell-bounding potential handlers) and
    power_bounded.ini (no cell system: the pair factor of the factor type map with its bounding potential), run
    unchanged except for the mediator line, the output file and the run length, must reproduce the cumulative
    histogram of the pair separation that the reference ships (ReferenceDataCoulombAtoms.dat, reversible Monte Carlo;
    fixture tests/golden/reference_cdfs.npz). 2048 chains in parallel give ~10^5 samples in a few seconds."""
    import sys
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import kat_replay as kr
    import jellyfysh_b200
    jellyfysh_b200.install()
    from jellyfysh.base.exceptions import EndOfRun
    path = os.path.join(REF, "jellyfysh", "config_files", "2018_JCP_149_064113", "coulomb_atoms", config)
    ini = open(path).read().replace("filename = config_files/",
                                    "filename = " + os.path.join(REF, "jellyfysh", "config_files") + "/")
    chains = 2048
    ini = ini.replace("mediator = single_process_mediator", "mediator = cuda_batched_mediator")
    ini = ini.replace("[SingleProcessMediator]", "[CudaBatchedMediator]\nnumber_of_chains = %d\nseed = 11" % chains)
    ini = ini.replace("end_of_run_time = 100000", "end_of_run_time = 40")
    ini = ini.replace("output/2018_JCP_149_064113/coulomb_atoms/" + output, str(tmp_path / "separation.dat"))
    assert "cuda_batched_mediator" in ini and str(tmp_path) in ini
    mediator, setting = build_reference_graph(ini)
    try:
        with pytest.raises(EndOfRun):
            mediator.run()
        mediator.post_run()
        stats = mediator.statistics
    finally:
        setting.reset()
    assert stats["bound_violations"] == 0 and stats["capacity_errors"] == 0
    samples = np.loadtxt(tmp_path / "separation.dat", comments="#")
    per_chain = int(40 / 0.56789)
    assert len(samples) == chains * per_chain
    samples = samples.reshape(per_chain, chains)[10:].ravel()  # drop the first samples of every chain (random start)
    g = kr.load_npz("reference_cdfs")
    x, cdf = g["coulomb_atoms_x"], g["coulomb_atoms_cdf"]
    edges = x + 0.5 * (x[1] - x[0])  # the fixture tabulates the cumulative histogram at bin centres
    ours = np.searchsorted(np.sort(samples), edges, side="right") / len(samples)
    distance = np.max(np.abs(ours - cdf))
    # Kolmogorov-Smirnov: 1.95 / sqrt(n) at the 0.1 % level for n independent samples; consecutive samples of a chain are
    # correlated, so n is taken as the number of chains times a conservative 10 independent samples each
    assert distance < 1.95 / np.sqrt(chains * 10) + 1.0e-3, distance
    assert 0.2 < np.median(samples) < 0.8


def test_coulomb_atoms_at_scale_with_device_observables_and_estimators(tmp_path):
    """The shipped coulomb_atoms/cell_veto.ini with 4096 chains on two engines, its cell-veto bounds estimated on the device
    (device_estimators) and its SeparationOutputHandler samples accumulated on the device (device_observables): the
    cumulative histogram still follows ReferenceDataCoulombAtoms.dat, and sampling costs a small part of the run."""
    import sys
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import kat_replay as kr
    import jellyfysh_b200
    jellyfysh_b200.install()
    from jellyfysh.base.exceptions import EndOfRun
    path = os.path.join(REF, "jellyfysh", "config_files", "2018_JCP_149_064113", "coulomb_atoms", "cell_veto.ini")
    ini = open(path).read().replace("filename = config_files/",
                                    "filename = " + os.path.join(REF, "jellyfysh", "config_files") + "/")
    chains, bins = 4096, 4000
    ini = ini.replace("mediator = single_process_mediator", "mediator = cuda_batched_mediator")
    ini = ini.replace("[SingleProcessMediator]", "[CudaBatchedMediator]\nnumber_of_chains = %d\nseed = 12\ndevices = 0, 0\n"
                      "device_observables = true\ndevice_estimators = true\nhistogram_bins = %d\n"
                      "equilibration_samples = 10" % (chains, bins))
    ini = ini.replace("end_of_run_time = 100000", "end_of_run_time = 40")
    ini = ini.replace("output/2018_JCP_149_064113/coulomb_atoms/SamplesOfSeparation_CellVeto.dat", str(tmp_path / "separation.dat"))
    mediator, setting = build_reference_graph(ini)
    try:
        with pytest.raises(EndOfRun):
            mediator.run()
        mediator.post_run()
        stats, timings = mediator.statistics, mediator.timings
        observable = mediator.observables["separation_output_handler"]
    finally:
        setting.reset()
    assert stats["bound_violations"] == 0 and stats["capacity_errors"] == 0
    per_chain = int(40 / 0.56789) - 10
    assert observable["samples"] == per_chain and int(observable["counts"].sum()) == chains * per_chain  # two atoms: one pair
    g = kr.load_npz("reference_cdfs")
    x, cdf = g["coulomb_atoms_x"], g["coulomb_atoms_cdf"]
    edges = x + 0.5 * (x[1] - x[0])
    cumulative = np.concatenate([[0.0], np.cumsum(observable["counts"])]) / observable["counts"].sum()
    ours = np.interp(edges, observable["edges"], cumulative)
    distance = np.max(np.abs(ours - cdf))
    print("KS distance", distance, "timings", timings)
    assert distance < 1.95 / np.sqrt(chains * 10) + 1.0e-3 + 1.0 / bins, distance
    assert timings["output_seconds"] < 0.25 * timings["advance_seconds"], timings


def _dipole_ini(tmp_path, chains, end_of_run_time, sampling):
    ini = configs.hard_disk_dipoles_cells_ini(REF, end_of_run_time=end_of_run_time, sampling=sampling,
                                              output=str(tmp_path / "polarization.dat"))
    ini = ini.replace("mediator = single_process_mediator", "mediator = cuda_batched_mediator")
    ini = ini.replace("[SingleProcessMediator]", "[CudaBatchedMediator]\nnumber_of_chains = %d\nseed = 17\n"
                                                 "first_random_stream = 9\noccupant_capacity = 6" % chains)
    return ini


def test_shipped_hard_disk_dipoles_config_runs_on_the_device(oracle, tmp_path):
    """C1: config_files/hard_disk_dipoles/hard_disk_dipoles_cells.ini through CudaBatchedMediator (only the mediator
    line and the input handler differ, see configs.hard_disk_dipoles_cells_ini), from the shipped start configuration.
    The final state handed back to the reference's tree state handler (leaf and root units, velocities with the
    root's weight, time stamps) is the one the oracle -- bit-exact with the reference on this trace -- reaches."""
    import sys
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from jellyfysh.base.exceptions import EndOfRun
    import jellyfysh_b200
    jellyfysh_b200.install()
    g = tu.load_trace("trace_hard_disk_dipoles")
    n_roots = len(g["roots0"])
    composites = (g["roots0"], g["positions0"].reshape(n_roots, 2, -1))
    end = 30.0
    mediator, setting = build_reference_graph(_dipole_ini(tmp_path, 1, end, sampling=True), composites=composites)
    try:
        with pytest.raises(EndOfRun):
            mediator.run()
        mediator.post_run()
        state = mediator._state_handler.extract_global_state()
        roots = np.array([node.value.position for node in state])
        leaves = np.array([child.value.position for node in state for child in node.children])
        active = [(node.value.identifier, node.value.velocity) for node in state if node.value.velocity is not None]
        active += [(child.value.identifier, child.value.velocity) for node in state for child in node.children
                   if child.value.velocity is not None]
        stats = mediator.statistics
    finally:
        setting.reset()
    assert stats["events"] > 400 and stats["bond_events"] > 20 and stats["capacity_errors"] == 0
    chain = oracle.OracleChain(tu.dipole_builder_of(g, oracle.ProgramBuilder))
    chain.set_positions(g["positions0"])
    chain.set_roots(g["roots0"])
    chain.start(stream=int(g["seed"][1]))
    total = 0
    for k in list(range(1, int(end / 10.01) + 1)) + [None]:
        until = oracle.time_from_float(end if k is None else 10.01 * k)
        n, _ = chain.run(until=until)
        total += n
    assert total == stats["events"]
    assert np.max(np.abs(leaves - chain.positions())) < 1e-12 * 12.836
    assert np.max(np.abs(roots - chain.roots())) < 1e-12 * 12.836
    st = chain.state()
    assert len(active) == 2  # the active disk and its dipole
    assert active[0][0] == (st.active // 2,) and active[1][0] == (st.active // 2, st.active % 2)
    assert active[1][1][st.direction] == 1.0 and active[0][1][st.direction] == 0.5
    samples = np.loadtxt(tmp_path / "polarization.dat", comments="#")
    assert samples.shape == (int(end / 10.01) + 1, 2)  # first_event_time_zero = True


def test_hard_disk_dipoles_polarization_statistics(tmp_path):
    """Statistical check of C1: the polarization of 81 hard-disk dipoles sampled by the reference's own
    PolarizationOutputHandler from 192 device chains follows the cumulative histograms the reference ships
    (ReferenceDataP{x,y}_81Dipoles_NewtonianECMC.dat, fixture tests/golden/reference_cdfs.npz)."""
    import sys
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import kat_replay as kr
    from jellyfysh.base.exceptions import EndOfRun
    import jellyfysh_b200
    jellyfysh_b200.install()
    g = tu.load_trace("trace_hard_disk_dipoles")
    n_roots = len(g["roots0"])
    composites = (g["roots0"], g["positions0"].reshape(n_roots, 2, -1))
    # the polarization is a slow collective variable (relaxation time ~3 x 10^4 time units measured here: the distance
    # to the reference histogram falls from 0.46 over 1.5 x 10^3 time units to 0.10 over 6 x 10^4): sample every 3003
    # time units (the shipped file: 10.01) over 3 x 10^5 time units per chain, about 6 x 10^6 events each
    chains, end, interval = 256, 300300.0, 3003.0
    ini = _dipole_ini(tmp_path, chains, end, sampling=True).replace("sampling_interval = 10.01",
                                                                    "sampling_interval = %r" % interval)
    assert "sampling_interval = 3003.0" in ini
    mediator, setting = build_reference_graph(ini, composites=composites)
    try:
        with pytest.raises(EndOfRun):
            mediator.run()
        mediator.post_run()
        stats = mediator.statistics
    finally:
        setting.reset()
    assert stats["capacity_errors"] == 0
    samples = np.loadtxt(tmp_path / "polarization.dat", comments="#")
    per_chain = int(end / interval) + 1
    assert samples.shape == (chains * per_chain, 2)
    samples = samples.reshape(per_chain, chains, 2)[per_chain // 5:].reshape(-1, 2)  # all chains share the start
    ref = kr.load_npz("reference_cdfs")
    for axis, key in enumerate(("dipoles_px", "dipoles_py")):
        x, cdf = ref[key + "_x"], ref[key + "_cdf"]
        edges = x + 0.5 * (x[1] - x[0])
        ours = np.searchsorted(np.sort(samples[:, axis]), edges, side="right") / len(samples)
        distance = np.max(np.abs(ours - cdf))
        print(key, "KS distance", distance, "samples", len(samples))
        # Kolmogorov-Smirnov at the 0.1 % level, counting a conservative 10 independent samples per chain
        assert distance < 1.95 / np.sqrt(chains * 10) + 5.0e-3, (key, distance)


def _sequential_dipole_ini(tmp_path, chains, end_of_run_time, sampling):
    ini = configs.hard_disk_dipoles_ini(REF, end_of_run_time=end_of_run_time, sampling=sampling,
                                        output=str(tmp_path / "polarization.dat"))
    ini = ini.replace("mediator = single_process_mediator", "mediator = cuda_batched_mediator")
    ini = ini.replace("[SingleProcessMediator]", "[CudaBatchedMediator]\nnumber_of_chains = %d\nseed = 17\n"
                                                 "first_random_stream = 9" % chains)
    return ini


def test_shipped_hard_disk_dipoles_without_cells_runs_on_the_device(oracle, tmp_path):
    """config_files/hard_disk_dipoles/hard_disk_dipoles.ini (no cell system, general velocities: the sequential-direction
    end-of-chain handler) through CudaBatchedMediator with only the mediator line changed, from the shipped PDB start
    configuration. The final state handed back to the reference's tree state handler -- leaf and root positions, the
    general velocities of the active leaf and of its root unit -- is the one the oracle reaches (bit-exact with the
    reference on this configuration, tests/test_oracle_traces.py)."""
    import sys
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from jellyfysh.base.exceptions import EndOfRun
    import jellyfysh_b200
    jellyfysh_b200.install()
    g = tu.load_trace("trace_hard_disk_dipoles_sequential")
    end = 25.0
    mediator, setting = build_reference_graph(_sequential_dipole_ini(tmp_path, 1, end, sampling=True))
    try:
        assert "disk_kernel" in mediator.engine.kernel_name()
        with pytest.raises(EndOfRun):
            mediator.run()
        mediator.post_run()
        state = mediator._state_handler.extract_global_state()
        roots = np.array([node.value.position for node in state])
        leaves = np.array([child.value.position for node in state for child in node.children])
        active = [(node.value.identifier, node.value.velocity) for node in state if node.value.velocity is not None]
        active += [(child.value.identifier, child.value.velocity) for node in state for child in node.children
                   if child.value.velocity is not None]
        stats = mediator.statistics
        builder = mediator._compiled.builder
    finally:
        setting.reset()
    assert stats["events"] > 100 and stats["bond_events"] > 20 and stats["factor_pair_events"] > 20
    assert stats["end_of_chain_events"] == int(end / 6.0)
    chain = oracle.OracleChain(builder)
    chain.set_positions(g["positions0"])
    chain.set_roots(g["roots0"])
    chain.start(stream=9)
    total = 0
    for k in list(range(1, int(end / 10.01) + 1)) + [None]:
        until = oracle.time_from_float(end if k is None else 10.01 * k)
        n, _ = chain.run(until=until)
        total += n
    assert total == stats["events"]
    assert np.max(np.abs(leaves - chain.positions())) < 1e-12 * 12.836
    assert np.max(np.abs(roots - chain.roots())) < 1e-12 * 12.836
    st = chain.state()
    assert len(active) == 2  # the active disk and its dipole
    assert active[0][0] == (st.active // 2,) and active[1][0] == (st.active // 2, st.active % 2)
    assert np.max(np.abs(np.array(active[1][1]) - np.array(st.velocity[:]))) < 1e-12
    assert np.max(np.abs(np.array(active[0][1]) - np.array(st.root_velocity[:]))) < 1e-12
    assert all(abs(v) > 1e-3 for v in active[1][1])  # rotated four times by 20 degrees: not along an axis
    samples = np.loadtxt(tmp_path / "polarization.dat", comments="#")
    assert samples.shape == (int(end / 10.01) + 1, 2)  # first_event_time_zero = True


def test_hard_disk_dipoles_without_cells_polarization_statistics(tmp_path):
    """Statistical check of the shipped hard_disk_dipoles.ini on the device: the polarization of the 81 dipoles
    sampled by the reference's own PolarizationOutputHandler from 256 device chains with general velocities follows the
    cumulative histograms the reference ships for this system (ReferenceDataP{x,y}_81Dipoles_NewtonianECMC.dat)."""
    import sys
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import kat_replay as kr
    from jellyfysh.base.exceptions import EndOfRun
    import jellyfysh_b200
    jellyfysh_b200.install()
    chains, end, interval = 256, 300300.0, 3003.0  # see test_hard_disk_dipoles_polarization_statistics
    ini = _sequential_dipole_ini(tmp_path, chains, end, sampling=True).replace("sampling_interval = 10.01",
                                                                               "sampling_interval = %r" % interval)
    assert "sampling_interval = 3003.0" in ini
    mediator, setting = build_reference_graph(ini)
    try:
        with pytest.raises(EndOfRun):
            mediator.run()
        mediator.post_run()
        stats = mediator.statistics
    finally:
        setting.reset()
    # (the last end of chain falls on the end-of-run time itself and is not run)
    assert stats["capacity_errors"] == 0 and stats["end_of_chain_events"] == chains * (int(end / 6.0) - 1)
    samples = np.loadtxt(tmp_path / "polarization.dat", comments="#")
    per_chain = int(end / interval) + 1
    assert samples.shape == (chains * per_chain, 2)
    samples = samples.reshape(per_chain, chains, 2)[per_chain // 5:].reshape(-1, 2)  # all chains share the start
    ref = kr.load_npz("reference_cdfs")
    for axis, key in enumerate(("dipoles_px", "dipoles_py")):
        x, cdf = ref[key + "_x"], ref[key + "_cdf"]
        edges = x + 0.5 * (x[1] - x[0])
        ours = np.searchsorted(np.sort(samples[:, axis]), edges, side="right") / len(samples)
        distance = np.max(np.abs(ours - cdf))
        print(key, "KS distance", distance, "samples", len(samples), "events", stats["events"])
        assert distance < 1.95 / np.sqrt(chains * 10) + 5.0e-3, (key, distance)


def test_shipped_single_hard_disk_dipole_matches_the_analytic_distributions(tmp_path):
    """config_files/hard_disk_dipoles/single_hard_disk_dipole.ini on the device (one tethered pair of disks, the velocity
    rotated by 23 degrees at every end of chain): the polarization sampled by the reference's PolarizationOutputHandler from
    512 chains -- each from its own draw of the reference's random input handler -- follows the analytic distributions the
    reference's plot script compares with (output/hard_disk_dipoles/plot_histogram_single_hard_disk_dipole.py:61-77,
    :104-106): extension rho with cumulative (rho^2 - min^2) / (max^2 - min^2), angle uniform on (-pi, pi]."""
    import sys
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from jellyfysh.base.exceptions import EndOfRun
    import jellyfysh_b200
    jellyfysh_b200.install()
    chains, interval, per_chain = 512, 2.215132, 120
    end = interval * (per_chain - 0.5)
    ini = configs.shipped_ini(REF, "hard_disk_dipoles", "single_hard_disk_dipole.ini")
    ini = ini.replace("filename = config_files/", "filename = " + os.path.join(REF, "jellyfysh", "config_files") + "/")
    ini = ini.replace("end_of_run_time = 3322698", "end_of_run_time = %r" % end)
    ini = ini.replace("output/hard_disk_dipoles/Polarization_SingleHardDiskDipole.dat", str(tmp_path / "polarization.dat"))
    ini = ini.replace("mediator = single_process_mediator", "mediator = cuda_batched_mediator")
    ini = ini.replace("[SingleProcessMediator]", "[CudaBatchedMediator]\nnumber_of_chains = %d\nseed = 5" % chains)
    assert repr(end) in ini and "polarization.dat" in ini
    mediator, setting = build_reference_graph(ini)
    try:
        assert "disk_kernel" in mediator.engine.kernel_name()
        with pytest.raises(EndOfRun):
            mediator.run()
        mediator.post_run()
        stats = mediator.statistics
    finally:
        setting.reset()
    assert stats["bond_events"] > 0 and stats["factor_pair_events"] == 0
    assert stats["end_of_chain_events"] == chains * int(end / 0.5)
    samples = np.loadtxt(tmp_path / "polarization.dat", comments="#")
    assert samples.shape == (chains * per_chain, 2)
    samples = samples.reshape(per_chain, chains, 2)[per_chain // 6:].reshape(-1, 2)
    rho, theta = np.hypot(samples[:, 0], samples[:, 1]), np.arctan2(samples[:, 1], samples[:, 0])
    lo, hi = 0.6666666666666666, 1.333333333333333
    assert rho.min() >= lo - 1e-12 and rho.max() <= hi + 1e-12
    grid = np.linspace(lo, hi, 1001)
    ours = np.searchsorted(np.sort(rho), grid, side="right") / len(rho)
    d_rho = np.max(np.abs(ours - (grid ** 2 - lo ** 2) / (hi ** 2 - lo ** 2)))
    grid = np.linspace(-np.pi, np.pi, 1001)
    ours = np.searchsorted(np.sort(theta), grid, side="right") / len(theta)
    d_theta = np.max(np.abs(ours - (grid + np.pi) / (2.0 * np.pi)))
    print("KS distances: extension", d_rho, "angle", d_theta, "samples", len(rho), "events", stats["events"])
    bound = 1.95 / np.sqrt(chains * 10) + 5.0e-3  # conservative: 10 independent samples per chain
    assert d_rho < bound and d_theta < bound, (d_rho, d_theta, bound)


@pytest.mark.parametrize("config", ["coulomb_cell_veto_lj_inverted.ini", "coulomb_power_bounded_lj_inverted.ini",
                                    "coulomb_power_bounded_lj_cell_bounded.ini"])
def test_shipped_water_config_matches_reference_statistics(tmp_path, config):
    """C4 statistical check: the shipped water/coulomb_cell_veto_lj_inverted.ini (two SPC/Fw molecules;
    composite-object Coulomb handlers on root-level cells) and water/coulomb_power_bounded_lj_inverted.ini (no cell
    system, Coulomb as nine bounded leaf-to-leaf factors between the molecules) and
    water/coulomb_power_bounded_lj_cell_bounded.ini (composite-object Coulomb factors from the factor type map, the
    Lennard-Jones factor of the oxygens through an oxygen-only cell system), unchanged except for the mediator line,
    the run length, the sampling interval and the output file, run as many independent device chains, reproduce the
    cumulative histogram of the oxygen-oxygen separation the reference ships (ReferenceOOSeparation.dat; fixture
    tests/golden/reference_cdfs.npz)."""
    import sys
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import kat_replay as kr
    from jellyfysh.base.exceptions import EndOfRun
    import jellyfysh_b200
    jellyfysh_b200.install()
    cell_veto = config == "coulomb_cell_veto_lj_inverted.ini"
    # Two molecules that start apart have to find each other. The pairwise factors of the second file move the pair
    # together more slowly than the composite-object handlers with their lifting over six leaves: measured on B200, the
    # distance to the reference histogram after simulated times 500 / 2000 / 8000 is 0.72 / 0.24 / 0.003.
    leaf_cells = config == "coulomb_power_bounded_lj_cell_bounded.ini"
    chains, end, interval = 1024, 2000.0 if cell_veto or leaf_cells else 8000.0, 20.0
    if cell_veto:
        ini = configs.water_ini(REF, n_molecules=2, end_of_run_time=end, sampling_interval=interval,
                                output=str(tmp_path / "oo_separation.dat"))
        assert "number_trials = 1000" in ini
    else:
        ini = configs.shipped_ini(REF, "2018_JCP_149_064113", "water", config)
        ini = ini.replace("filename = config_files/", "filename = " + os.path.join(REF, "jellyfysh", "config_files") + "/")
        ini = ini.replace("end_of_run_time = 500000", "end_of_run_time = %r" % end)
        ini = ini.replace("sampling_interval = 2.6789", "sampling_interval = %r" % interval)
        ini = ini.replace("output/2018_JCP_149_064113/water/SamplesOfOOSeparation_CoulombPowerBounded_" +
                          ("LJCellBounded.dat" if leaf_cells else "LJInverted.dat"), str(tmp_path / "oo_separation.dat"))
        assert str(tmp_path) in ini and "sampling_interval = 20.0" in ini
    ini = ini.replace("mediator = single_process_mediator", "mediator = cuda_batched_mediator")
    ini = ini.replace("[SingleProcessMediator]", "[CudaBatchedMediator]\nnumber_of_chains = %d\nseed = 29" % chains)
    assert "number_of_root_nodes = 2" in ini
    # Every chain starts from its own pair of well separated molecules: the random input handler of the shipped file
    # can put a hydrogen next to the other oxygen, which has no repulsive core for it -- such a chain collapses (event
    # rate -> infinity) in the reference as well.
    starts = [configs.water_start(2, 10.0, seed=500 + c) for c in range(chains)]
    composites = (np.concatenate([r for r, _ in starts]), np.concatenate([l for _, l in starts]))
    mediator, setting = build_reference_graph(ini, composites=composites)
    try:
        with pytest.raises(EndOfRun):
            mediator.run()
        mediator.post_run()
        stats = mediator.statistics
    finally:
        setting.reset()
    assert stats["capacity_errors"] == 0 and stats["bond_events"] > 0 and (stats["veto_events"] > 0) == cell_veto
    assert (stats["boundary_events"] > 0) == (cell_veto or leaf_cells)
    samples = np.loadtxt(tmp_path / "oo_separation.dat", comments="#")
    per_chain = int(end / interval)
    assert len(samples) == chains * per_chain
    samples = samples.reshape(per_chain, chains)[per_chain // 2:].ravel()  # random starts: let the pairs find each other
    ref = kr.load_npz("reference_cdfs")
    x, cdf = ref["water_oo_x"], ref["water_oo_cdf"]
    edges = x + 0.5 * (x[1] - x[0])
    ours = np.searchsorted(np.sort(samples), edges, side="right") / len(samples)
    distance = np.max(np.abs(ours - cdf))
    print("water O-O KS distance", distance, "samples", len(samples), "median", np.median(samples), stats)
    assert distance < 1.95 / np.sqrt(chains * 5) + 5.0e-3, distance


def _dipole_pair_start(seed, length=1.0):
    """Two dipoles (charges +1, -1 at separation ~0.1) with centres at least 0.3 apart, random orientations."""
    rng = np.random.default_rng(seed)
    while True:
        centres = rng.uniform(0.0, length, size=(2, 3))
        d = np.mod(centres[1] - centres[0] + length / 2, length) - length / 2
        if np.linalg.norm(d) > 0.3:
            break
    roots = np.empty((2, 3))
    leaves = np.empty((2, 2, 3))
    for k in range(2):
        axis = rng.normal(size=3)
        axis *= 0.05 / np.linalg.norm(axis)
        roots[k] = centres[k] % length
        leaves[k, 0] = (centres[k] + axis) % length
        leaves[k, 1] = (centres[k] - axis) % length
    return roots, leaves


@pytest.mark.parametrize("config,output", [
    ("cell_veto.ini", "SamplesOfSeparation_CellVeto.dat"),
    ("dipole_factors_inside_first.ini", "SamplesOfSeparation_DipoleFactors_InsideFirst.dat"),
    ("dipole_factors_outside_first.ini", "SamplesOfSeparation_DipoleFactors_OutsideFirst.dat"),
    ("dipole_factors_ratio.ini", "SamplesOfSeparation_DipoleFactors_Ratio.dat"),
    ("atom_factors.ini", "SamplesOfSeparation_AtomFactors.dat"),
    ("cell_bounded.ini", "SamplesOfSeparation_CellBounded.dat"),
    ("dipole_motion.ini", "SamplesOfSeparation_DipoleMotion.dat")])
def test_shipped_dipole_config_matches_reference_statistics(tmp_path, config, output):
    """dipoles/cell_veto.ini of 2018_JCP_149_064113 (two dipoles: composite-object Coulomb handlers with cell veto on
    anisotropic 3 x 5 x 7 root-level cells, harmonic bond, 1/r^6 repulsion between the opposite charges of different
    dipoles as a factor between objects) and the three dipole_factors_*.ini (no cell system: the composite-object
    Coulomb factor of the factor type map with inside-first, outside-first and ratio lifting) and atom_factors.ini (the
    Coulomb interaction as four bounded leaf-to-leaf factors between the dipoles) and dipole_motion.ini (the independent
    active unit alternates between a leaf unit and the root unit of a dipole), unchanged except for the
    mediator line, run length, sampling interval and output file: the separations between like and unlike charges of
    different dipoles follow the cumulative histograms the reference ships (ReferenceDataDipoles_13.dat / _14.dat)."""
    import sys
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import kat_replay as kr
    from jellyfysh.base.exceptions import EndOfRun
    import jellyfysh_b200
    jellyfysh_b200.install()
    chains, end, interval = 1024, 400.0, 4.0
    ini = configs.shipped_ini(REF, "2018_JCP_149_064113", "dipoles", config)
    ini = ini.replace("filename = config_files/", "filename = " + os.path.join(REF, "jellyfysh", "config_files") + "/")
    ini = ini.replace("mediator = single_process_mediator", "mediator = cuda_batched_mediator")
    ini = ini.replace("[SingleProcessMediator]", "[CudaBatchedMediator]\nnumber_of_chains = %d\nseed = 31" % chains)
    ini = ini.replace("end_of_run_time = 500000", "end_of_run_time = %r" % end)
    ini = ini.replace("sampling_interval = 0.56789", "sampling_interval = %r" % interval)
    ini = ini.replace("output/2018_JCP_149_064113/dipoles/" + output, str(tmp_path / "separation.dat"))
    assert "cuda_batched_mediator" in ini and str(tmp_path) in ini and "end_of_run_time = 400.0" in ini
    starts = [_dipole_pair_start(900 + c) for c in range(chains)]
    composites = (np.concatenate([r for r, _ in starts]), np.concatenate([l for _, l in starts]))
    mediator, setting = build_reference_graph(ini, composites=composites)
    try:
        with pytest.raises(EndOfRun):
            mediator.run()
        mediator.post_run()
        stats = mediator.statistics
    finally:
        setting.reset()
    assert stats["capacity_errors"] == 0 and stats["factor_pair_events"] > 0 and stats["bond_events"] > 0
    ref = kr.load_npz("reference_cdfs")
    per_chain = int(end / interval)
    for name, pairs in (("13", 2), ("14", 2)):
        samples = np.loadtxt(tmp_path / ("separation_%s.dat" % name), comments="#")
        assert len(samples) == chains * per_chain * pairs
        samples = samples.reshape(per_chain, chains * pairs)[per_chain // 2:].ravel()
        x, cdf = ref["dipoles_%s_x" % name], ref["dipoles_%s_cdf" % name]
        edges = x + 0.5 * (x[1] - x[0])
        ours = np.searchsorted(np.sort(samples), edges, side="right") / len(samples)
        distance = np.max(np.abs(ours - cdf))
        print("dipoles", name, "KS distance", distance, "samples", len(samples), stats)
        assert distance < 1.95 / np.sqrt(chains * 5) + 5.0e-3, (name, distance)


def test_shipped_single_molecule_config_matches_reference_statistics(tmp_path):
    """water/single_molecule.ini of 2018_JCP_149_064113 (one SPC/Fw molecule, no cell system: the two harmonic bonds and
    the bending factor with its piecewise constant bound and ratio lifting), unchanged except for the mediator line, run
    length and output file: bond lengths and bond angle, sampled by the reference's own BondLengthAndAngleOutputHandler,
    follow the cumulative histograms the reference ships (ReferenceLengthSingleMolecule.dat / ReferenceAngle...)."""
    import sys
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import kat_replay as kr
    from jellyfysh.base.exceptions import EndOfRun
    import jellyfysh_b200
    jellyfysh_b200.install()
    chains, end = 1024, 300.0
    ini = configs.shipped_ini(REF, "2018_JCP_149_064113", "water", "single_molecule.ini")
    ini = ini.replace("filename = config_files/", "filename = " + os.path.join(REF, "jellyfysh", "config_files") + "/")
    ini = ini.replace("mediator = single_process_mediator", "mediator = cuda_batched_mediator")
    ini = ini.replace("[SingleProcessMediator]", "[CudaBatchedMediator]\nnumber_of_chains = %d\nseed = 37" % chains)
    ini = ini.replace("end_of_run_time = 500000", "end_of_run_time = %r" % end)
    ini = ini.replace("output/2018_JCP_149_064113/water/SamplesOfBonds_SingleMolecule.dat", str(tmp_path / "bonds.dat"))
    assert "cuda_batched_mediator" in ini and str(tmp_path) in ini and "end_of_run_time = 300.0" in ini
    starts = [configs.water_start(1, 10.0, seed=500 + c) for c in range(chains)]
    composites = (np.concatenate([r for r, _ in starts]), np.concatenate([l for _, l in starts]))
    mediator, setting = build_reference_graph(ini, composites=composites)
    try:
        with pytest.raises(EndOfRun):
            mediator.run()
        mediator.post_run()
        stats = mediator.statistics
    finally:
        setting.reset()
    assert stats["capacity_errors"] == 0 and stats["bond_events"] > 0 and stats["pair_events"] == 0
    ref = kr.load_npz("reference_cdfs")
    per_chain = int(end / 2.6789)
    for name, suffix, per_sample in (("water_length", "Length", 2), ("water_angle", "Angle", 1)):
        samples = np.loadtxt(tmp_path / ("bonds_%s.dat" % suffix), comments="#")
        assert len(samples) == chains * per_chain * per_sample
        samples = samples.reshape(per_chain, chains * per_sample)[per_chain // 4:].ravel()
        x, cdf = ref[name + "_x"], ref[name + "_cdf"]
        edges = x + 0.5 * (x[1] - x[0])
        ours = np.searchsorted(np.sort(samples), edges, side="right") / len(samples)
        distance = np.max(np.abs(ours - cdf))
        print("single molecule", name, "KS distance", distance, "samples", len(samples), stats)
        assert distance < 1.95 / np.sqrt(chains * 5) + 5.0e-3, (name, distance)


def test_shipped_dump_config_dumps_and_resumes(tmp_path):
    """coulomb_atoms/power_bounded_dump.ini: its FixedIntervalDumpingEventHandler writes the device checkpoint next to the
    DumpingOutputHandler's file name, and a second mediator built from the same file with `resume_file` continues from
    the last dump to the end of the run exactly like the uninterrupted run (the role of resume.py)."""
    import sys
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from jellyfysh.base.exceptions import EndOfRun
    import jellyfysh_b200
    jellyfysh_b200.install()
    chains = 64
    ini = configs.shipped_ini(REF, "2018_JCP_149_064113", "coulomb_atoms", "power_bounded_dump.ini")
    ini = ini.replace("filename = config_files/", "filename = " + os.path.join(REF, "jellyfysh", "config_files") + "/")
    ini = ini.replace("mediator = single_process_mediator", "mediator = cuda_batched_mediator")
    ini = ini.replace("end_of_run_time = 2000", "end_of_run_time = 30")
    ini = ini.replace("dumping_interval = 1100", "dumping_interval = 11")
    ini = ini.replace("filename = dump.dat", "filename = " + str(tmp_path / "dump.dat"))
    ini = ini.replace("output/2018_JCP_149_064113/coulomb_atoms/SamplesOfSeparation_PowerBoundedDump.dat",
                      str(tmp_path / "separation.dat"))
    assert str(tmp_path / "dump.dat") in ini and "dumping_interval = 11" in ini and "end_of_run_time = 30" in ini

    def run(extra):
        text = ini.replace("[SingleProcessMediator]", "[CudaBatchedMediator]\nnumber_of_chains = %d\nseed = 41%s" % (chains, extra))
        mediator, setting = build_reference_graph(text)
        try:
            with pytest.raises(EndOfRun):
                mediator.run()
            mediator.post_run()
            return mediator.engine.download_positions(), mediator.engine.chain_states(), mediator.statistics
        finally:
            setting.reset()

    full_positions, full_states, full_stats = run("")
    # DumpingOutputHandler tags its file name with the interpreter (dump_cpython_3_12_3.dat); the device dump sits next to it
    import glob
    dump_path, = glob.glob(str(tmp_path / "dump*.dat.npz"))
    dump = np.load(dump_path)
    assert np.all(dump["chain_states"]["time_q"] == 22.0) and np.all(dump["chain_states"]["time_r"] == 0.0)
    dumped_events = int(dump["chain_states"]["event_counter"].sum())
    resumed_positions, resumed_states, resumed_stats = run("\nresume_file = " + dump_path)
    assert np.array_equal(full_positions, resumed_positions)
    for field in ("active", "direction", "time_q", "time_r", "event_counter", "eoc_q", "eoc_r"):
        assert np.array_equal(full_states[field], resumed_states[field]), field
    assert resumed_stats["events"] == full_stats["events"] - dumped_events > 0


def test_device_estimators_reproduce_the_reference_bounds():
    """jellyfysh_b200.estimators.accelerate on the reference's four estimator classes (inner_point_estimator.py:139-163,
    boundary_point_estimator.py:108-174, dipole_monte_carlo_estimator.py:100-155 -- same draws of Python's `random` in
    the same order --, dipole_inner_point_estimator.py:100-163): the bounds of the device-backed `derivative_bound`
    against the reference's own point-by-point loop for cells of the water cell system, merged-image Coulomb potential."""
    import random
    import sys
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from jellyfysh.base.exceptions import EndOfRun  # noqa: F401 - the package must be importable
    from jellyfysh.setting import hypercubic_setting
    import jellyfysh.setting as setting
    from jellyfysh.estimator.inner_point_estimator import InnerPointEstimator
    from jellyfysh.estimator.boundary_point_estimator import BoundaryPointEstimator
    from jellyfysh.estimator.dipole_monte_carlo_estimator import DipoleMonteCarloEstimator
    from jellyfysh.estimator.dipole_inner_point_estimator import DipoleInnerPointEstimator
    from jellyfysh.potential.merged_image_coulomb_potential import MergedImageCoulombPotential
    from jellyfysh.potential.lennard_jones_potential import LennardJonesPotential
    from jellyfysh_b200 import estimators
    setting.reset()
    hypercubic_setting.HypercubicSetting(beta=1.679, dimension=3, system_length=10.0)
    setting.set_number_of_root_nodes(2)
    setting.set_number_of_nodes_per_root_node(3)
    setting.set_number_of_node_levels(2)
    try:
        coulomb = MergedImageCoulombPotential(prefactor=332.0)
        cases = [
            ("inner / Coulomb", lambda: InnerPointEstimator(potential=coulomb, prefactor=1.3, points_per_side=5)),
            ("inner / Lennard-Jones", lambda: InnerPointEstimator(
                potential=LennardJonesPotential(prefactor=0.6217012, characteristic_length=3.165492), prefactor=1.5,
                points_per_side=4, empirical_bound=60.0)),
            ("boundary / Coulomb", lambda: BoundaryPointEstimator(potential=coulomb, prefactor=1.2, points_per_side=6)),
            ("dipole Monte Carlo", lambda: DipoleMonteCarloEstimator(
                potential=coulomb, dipole_separation=1.3, dipole_charge=0.82, prefactor=1.1, number_trials=300)),
            ("dipole inner point", lambda: DipoleInnerPointEstimator(
                potential=coulomb, dipole_separation=1.3, dipole_charge=0.82, prefactor=1.2, points_per_side=4)),
        ]
        side = 10.0 / 6
        regions = [([3 * side - side, -side, -side], [3 * side + side, side, side]),          # three cells away along x
                   ([-4 * side, 2 * side, -3 * side], [-2 * side, 4 * side, -side]),           # a corner cell (wraps)
                   ([2 * side, 2 * side, 2 * side], [4 * side, 4 * side, 4 * side])]
        for name, make in cases:
            for lower, upper in regions:
                for direction in range(3):
                    for both in (True, False):
                        reference = make()
                        random.seed(77)
                        expected = reference.derivative_bound(list(lower), list(upper), direction, both)
                        ours = make()
                        assert estimators.accelerate(ours), name
                        random.seed(77)
                        got = ours.derivative_bound(list(lower), list(upper), direction, both)
                        assert len(got) == len(expected) == (2 if both else 1), name
                        # (the dipole inner point estimator normalises a gradient of nearly cancelling derivatives: the
                        # device's 1e-13 shows up as 1e-8 in the orientation of its dipoles)
                        tolerance = 1e-6 if name == "dipole inner point" else 1e-10
                        for a, b in zip(got, expected):
                            assert abs(a - b) <= tolerance * max(1.0, abs(b)), (name, lower, direction, got, expected)
    finally:
        setting.reset()
