"""CudaBatchedMediator -- the reference's mediator contract over the B200 event-chain engine.

Drop-in for jellyfysh/mediator/single_process_mediator.py:57-156: same constructor arguments (input-output handler,
state handler, scheduler, activator -- all built by the reference's factory from an unchanged INI file), same `run()`
/ `post_run()` protocol (run.py:183-200), raises base.exceptions.EndOfRun like mediator.py:375. Selected with

    [Run]
    mediator = cuda_batched_mediator

    [CudaBatchedMediator]
    state_handler = tree_state_handler
    scheduler = heap_scheduler
    activator = tag_activator
    input_output_handler = input_output_handler
    number_of_chains = 4096        ; independent Markov chains advanced at once (default 1)
    devices = 0, 1, 2, 3           ; CUDA devices; the chains are split into contiguous blocks, one engine per device
    seed = 0
    device_observables = true      ; histogram-type output handlers are accumulated on the device (see below)

(see INTEGRATION.md for how the module becomes `jellyfysh.mediator.cuda_batched_mediator`).

What moves to the device: every iteration of the reference loop whose winner is an interaction, cell-veto,
cell-boundary or end-of-chain handler. What stays on the host: sampling and end-of-run handlers, which the reference
calls at their own event times -- here between `ecmc_run(until = their time)` launches, with the chain states downloaded
into the reference's state handler so that the reference's output handlers see exactly what they expect.
The scheduler object is kept for the host control events only (the device argmin replaces it for interaction events).
Chain 0 starts from the state the input handler produced; chains c > 0 from further reads of the same input handler.

Sampling at scale (`device_observables = true`): the reference's output handlers print every sample of every chain from a
Python loop (SeparationOutputHandler: N (N - 1) / 2 lines per chain and sample). With thousands of chains the observables
are accumulated where the configurations live instead -- ecmc_separation_histogram[_subset] (SeparationOutputHandler,
OxygenOxygenSeparationOutputHandler), ecmc_polarization (PolarizationOutputHandler: one vector per chain and sample, written
in the reference's text format), ecmc_bond_histograms (BondLengthAndAngleOutputHandler) -- and the histograms are written
next to the output handler's file (`<filename>.histogram.npz`: edges, counts) at the end of the run; `observables` holds
them. Output handlers of other types keep the chain-by-chain path.
"""
import logging
from typing import Sequence

import numpy as np

from jellyfysh.activator import Activator
from jellyfysh.base.exceptions import EndOfRun
from jellyfysh.base.time import Time
from jellyfysh.input_output_handler import InputOutputHandler
from jellyfysh.mediator.mediator import Mediator
from jellyfysh.scheduler import Scheduler
from jellyfysh.state_handler import StateHandler

from jellyfysh_b200 import compiler, engine, sharding


class CudaBatchedMediator(Mediator):
    """Mediator that advances `number_of_chains` independent chains on one or several CUDA devices."""

    def __init__(self, input_output_handler: InputOutputHandler, state_handler: StateHandler, scheduler: Scheduler,
                 activator: Activator, number_of_chains: int = 1, device: int = 0, seed: int = 0,
                 first_random_stream: int = 0, maximum_surplus: int = 0, occupant_capacity: int = 8,
                 events_per_launch: int = 4000000, resume_file: str = "", devices: Sequence[int] = (),
                 device_observables: bool = False, histogram_bins: int = 1000, device_estimators: bool = False,
                 equilibration_samples: int = 0) -> None:
        """
        Parameters follow SingleProcessMediator (single_process_mediator.py:57-72); in addition:

        number_of_chains : independent Markov chains, all devices together.
        device : CUDA device index (one device).
        devices : CUDA device indices; the chains are split into contiguous blocks (sharding.split_chains), one engine
            per entry (an index may repeat: several engines on one device), chain c reads random stream
            first_random_stream + c whatever the number of devices, so the chains do not depend on the split.
        device_observables, histogram_bins : accumulate the samples of histogram-type output handlers on the devices.
        equilibration_samples : device observables skip the first samples of every output handler (start configuration).
        device_estimators : the estimators of cell-veto handlers / cell-bounding potentials evaluate their points on the
            device (jellyfysh_b200/estimators.py) when the activator is initialised, instead of point by point in Python.
        seed, first_random_stream : chain c reads the counter-based random stream (seed, first_random_stream + c).
        maximum_surplus : capacity of the per-chain surplus list (0: one slot per particle).
        occupant_capacity : occupants per cell kept on the device when the reference's cell occupancy is unbounded.
        events_per_launch : upper limit of events per chain and kernel launch; a chain that needs more to reach the
            next control time is continued by further launches, and one that stops advancing in time (a collapsing
            configuration, e.g. overlapping molecules with unbounded attraction) raises an error instead of hanging.
        resume_file : a dump written by a DumpingEventHandler of an earlier run of the same configuration (the file name
            of its DumpingOutputHandler plus ".npz"): the chains continue from it, event for event as the uninterrupted
            run would (the role of jellyfysh/resume.py, whose dill of the reference's mediator cannot hold device state).
        """
        self._logger = logging.getLogger(__name__)
        if number_of_chains < 1:
            raise compiler._configuration_error("number_of_chains must be at least 1")
        state_handler.initialize(input_output_handler.read())
        if device_estimators:
            from jellyfysh_b200 import estimators
            first_device = int(devices[0]) if len(devices) else int(device)
            self._logger.info("%d estimator(s) evaluate their points on device %d",
                              estimators.accelerate_activator(activator, first_device), first_device)
        super().__init__(input_output_handler, state_handler, scheduler, activator)
        template = state_handler.extract_global_state()
        self._compiled = compiler.compile_program(activator, template, seed=seed,
                                                  max_surplus=maximum_surplus if maximum_surplus > 0 else None,
                                                  occupant_capacity=occupant_capacity)
        self._number_of_chains = number_of_chains
        positions, charges, roots = compiler.positions_and_charges(template, self._compiled.charge_name)
        all_positions = np.empty((number_of_chains,) + positions.shape)
        all_charges = None if charges is None else np.empty((number_of_chains,) + charges.shape)
        all_roots = None if roots is None else np.empty((number_of_chains,) + roots.shape)
        for chain in range(number_of_chains):
            if chain > 0:
                positions, charges, roots = compiler.positions_and_charges(input_output_handler.read(),
                                                                           self._compiled.charge_name)
            all_positions[chain] = positions
            if charges is not None:
                all_charges[chain] = charges
            if roots is not None:
                all_roots[chain] = roots
        device_list = [int(d) for d in devices] if len(devices) else [int(device)]
        if number_of_chains < len(device_list):
            raise compiler._configuration_error("more devices than chains")
        self._shards = sharding.split_chains(number_of_chains, len(device_list))
        self._engines = []
        for (first, count), index in zip(self._shards, device_list):
            eng = engine.Engine(self._compiled.builder, n_chains=count, device=index)
            eng.upload_positions(all_positions[first:first + count],
                                 None if all_charges is None else all_charges[first:first + count])
            if all_roots is not None:
                eng.upload_roots(all_roots[first:first + count])
            eng.start(first_stream=first_random_stream + first)
            self._engines.append(eng)
        self._engine = self._engines[0]
        if hasattr(state_handler, "bind"):  # cuda_state_handler: reads single chains straight from the engines
            program = self._compiled.builder.program
            state_handler.bind(self._engines, self._shards, program.speed, program.dimension, self._compiled.nodes_per_root,
                               bool(program.eoc_sequential))
        self._device_observables = bool(device_observables)
        self._histogram_bins = int(histogram_bins)
        self._equilibration_samples = int(equilibration_samples)
        self._skipped = {}
        self._timings = {"advance_seconds": 0.0, "output_seconds": 0.0}
        self._observables = {}
        self._template_charges = {}
        self._statistics = {}
        self._control_times = {}
        self._events_per_launch = max(int(events_per_launch), 1)
        if resume_file:
            self._resume(resume_file, all_charges)

    # ---- state hand-over to the reference's state handler ------------------------------------------------------
    def _load_chain_into_state_handler(self, chain, positions, states, roots=None):
        """Write one chain's device state through the public state-handler contract
        (state_handler.py:63-165): positions of all units, velocity / time stamp of the active leaf unit and, for
        composite point objects, of its root unit (velocity * weight, event_handler/abstracts/abstracts.py:165-190)."""
        if hasattr(self._state_handler, "load"):
            self._state_handler.load(positions[chain], None if roots is None else roots[chain], states[chain])
            return
        from jellyfysh_b200.state_handler.cuda_state_handler import fill_tree
        program = self._compiled.builder.program
        cnodes = self._state_handler.extract_global_state()
        fill_tree(cnodes, positions[chain], None if roots is None else roots[chain], states[chain], program.speed,
                  program.dimension, self._compiled.nodes_per_root, bool(program.eoc_sequential))
        self._state_handler.insert_into_global_state(cnodes)

    def _download(self):
        positions = np.concatenate([eng.download_positions() for eng in self._engines])
        roots = (np.concatenate([eng.download_roots() for eng in self._engines])
                 if self._compiled.nodes_per_root > 1 else None)
        return positions, roots, self.chain_states()

    def chain_states(self):
        """Lifting state (EcmcChainState) of all chains, in chain order over the devices."""
        return np.concatenate([eng.chain_states() for eng in self._engines])

    # ---- observables accumulated on the devices ---------------------------------------------------------------
    def _sample_on_device(self, name):
        """One sampling event of the output handler `name`, if its observable has a device accumulator. The handlers
        and their observables: separation_output_handler.py:75-97, oxygen_oxygen_separation_output_handler.py:78-97,
        polarization_output_handler.py:76-101, bond_length_and_angle_output_handler.py:77-103."""
        output = self._input_output_handler._output_handlers_dictionary[name]
        kinds = {cls.__name__ for cls in type(output).__mro__}
        known = {"SeparationOutputHandler", "OxygenOxygenSeparationOutputHandler", "PolarizationOutputHandler",
                 "BondLengthAndAngleOutputHandler"}
        if kinds & known and self._skipped.get(name, 0) < self._equilibration_samples:
            self._skipped[name] = self._skipped.get(name, 0) + 1
            return True
        program = self._compiled.builder.program
        length, dimension, bins = float(program.system_length), int(program.dimension), self._histogram_bins
        npr = self._compiled.nodes_per_root
        if "SeparationOutputHandler" in kinds or "OxygenOxygenSeparationOutputHandler" in kinds:
            oxygen = "OxygenOxygenSeparationOutputHandler" in kinds
            if oxygen and npr != 3:
                return False
            entry = self._observables.setdefault(name, {
                "kind": "oxygen_oxygen_separation" if oxygen else "separation", "filename": output._output_filename,
                "edges": np.linspace(0.0, length * dimension ** 0.5 / 2.0, bins + 1),
                "counts": np.zeros(bins, dtype=np.uint64), "samples": 0})
            for eng in self._engines:
                eng.separation_histogram(bins, entry["edges"][0], entry["edges"][-1], out=entry["counts"],
                                         first=1 if oxygen else 0, stride=npr if oxygen else 1)
            entry["samples"] += 1
            return True
        if "PolarizationOutputHandler" in kinds and npr > 1:
            if name not in self._template_charges:
                template = self._state_handler.extract_global_state()
                self._template_charges[name] = np.array([child.value.charge[output._charge] for cnode in template
                                                         for child in cnode.children], dtype=np.float64)
            vectors = np.concatenate([eng.polarization(self._template_charges[name]) for eng in self._engines])
            for vector in vectors:  # one line per chain and sample, the reference's format
                print("\t".join(map(str, vector.tolist())), file=output._file)
            entry = self._observables.setdefault(name, {"kind": "polarization", "filename": output._output_filename,
                                                        "samples": 0})
            entry["samples"] += 1
            return True
        if "BondLengthAndAngleOutputHandler" in kinds and npr == 3 and dimension == 3:
            entry = self._observables.setdefault(name, {
                "kind": "bond_length_and_angle", "filename": output._output_filename,
                "length_edges": np.linspace(0.0, length / 2.0, bins + 1), "angle_edges": np.linspace(0.0, np.pi, bins + 1),
                "length_counts": np.zeros(bins, dtype=np.uint64), "angle_counts": np.zeros(bins, dtype=np.uint64),
                "samples": 0})
            for eng in self._engines:
                eng.bond_histograms(bins, (0.0, length / 2.0), (0.0, np.pi),
                                    out=(entry["length_counts"], entry["angle_counts"]))
            entry["samples"] += 1
            return True
        return False

    def _write_observables(self):
        for entry in self._observables.values():
            arrays = {key: value for key, value in entry.items() if isinstance(value, np.ndarray)}
            if arrays:
                np.savez(entry["filename"] + ".histogram.npz", samples=np.asarray(entry["samples"]), **arrays)

    @property
    def observables(self):
        """Observables accumulated on the devices so far: {output handler name: {kind, edges, counts, samples, ...}}."""
        return self._observables

    def _write_output(self, handler):
        if handler.output_handler is None:
            return
        if self._device_observables and self._sample_on_device(handler.output_handler):
            return
        positions, roots, states = self._download()
        for chain in range(self._number_of_chains):
            self._load_chain_into_state_handler(chain, positions, states, roots)
            self._input_output_handler.write(handler.output_handler, self._state_handler.extract_global_state())

    # ---- dumping / resuming (DumpingEventHandler + DumpingOutputHandler, dumping_output_handler.py:70-90; resume.py) ----
    def _dump(self, handler):
        """The device state of all chains plus the schedule of the control handlers, next to the file name the
        configuration gives its DumpingOutputHandler (that handler itself pickles the reference's mediator, which a device
        handle does not survive)."""
        output = self._input_output_handler._output_handlers_dictionary[handler.output_handler]
        path = output._output_filename + ".npz"
        controls = self._compiled.control_handlers
        times = np.array([[self._control_times[h].quotient, self._control_times[h].remainder] for h in controls])
        # the dumping handler itself has not been rescheduled yet: its next time follows from its own event time
        for index, eng in enumerate(self._engines):
            eng.save_checkpoint(path if index == 0 else path[:-4] + ".device%d.npz" % index)
        with np.load(path) as data:
            arrays = dict(data)
        arrays["control_times"] = times
        arrays["control_names"] = np.array([type(h).__name__ for h in controls])
        arrays["dumping_handler"] = np.array(controls.index(handler))
        np.savez_compressed(path, **arrays)
        print("Writing dump into file {0}".format(path))

    def _resume(self, path, charges):
        controls = self._compiled.control_handlers
        for index, (eng, (first, count)) in enumerate(zip(self._engines, self._shards)):
            eng.load_checkpoint(path if index == 0 else path[:-4] + ".device%d.npz" % index,
                                None if charges is None else charges[first:first + count])
        with np.load(path) as data:
            names, times, dumping = data["control_names"].tolist(), data["control_times"], int(data["dumping_handler"])
        if names != [type(h).__name__ for h in controls]:
            raise compiler._configuration_error("the dump {0} belongs to a configuration with other control handlers"
                                                .format(path))
        for index, handler in enumerate(controls):
            # fixed-interval handlers continue from their own last event time (fixed_interval_*_event_handler.py)
            handler._event_time = Time(float(times[index][0]), float(times[index][1]))
            if index == dumping:
                self._control_times[handler] = handler.send_event_time()  # the dump was written at this handler's event
            else:
                self._control_times[handler] = handler._event_time

    # ---- the loop ------------------------------------------------------------------------------------------------
    def run(self) -> None:
        """Advance all chains from control event to control event until the end-of-run handler fires."""
        controls = self._compiled.control_handlers
        for handler in controls:
            if handler not in self._control_times:
                self._control_times[handler] = handler.send_event_time()
        import time as _time
        while True:
            handler = min(controls, key=lambda h: self._control_times[h])
            event_time = self._control_times[handler]
            t0 = _time.perf_counter()
            self._advance_to(event_time)
            self._timings["advance_seconds"] += _time.perf_counter() - t0
            t0 = _time.perf_counter()
            try:
                self._handle_control_event(handler)
            finally:
                self._timings["output_seconds"] += _time.perf_counter() - t0

    def _handle_control_event(self, handler):
        """What the reference's mediate methods do for sampling, dumping and end-of-run handlers (mediator.py:377-385)."""
        self._event_handler_with_shortest_event_time = handler
        names = {cls.__name__ for cls in type(handler).__mro__}
        if "EndOfRunEventHandler" in names:
            self._write_output(handler)
            self._write_observables()
            positions, roots, states = self._download()
            self._load_chain_into_state_handler(0, positions, states, roots)
            raise EndOfRun
        if "DumpingEventHandler" in names:
            self._dump(handler)
        else:
            self._write_output(handler)
        self._control_times[handler] = handler.send_event_time()

    def _advance_to(self, event_time):
        """ecmc_run(until = event_time), in launches of at most events_per_launch events per chain."""
        until = (event_time.quotient, event_time.remainder)
        stalled = 0
        while True:
            before = self.chain_states()
            for eng in self._engines:  # asynchronous: the devices run side by side
                eng.run(until=until, max_events=self._events_per_launch)
            for eng in self._engines:
                for key, value in eng.sync().items():
                    self._statistics[key] = self._statistics.get(key, 0) + value
            after = self.chain_states()
            behind = (after["time_q"] != until[0]) | (after["time_r"] != until[1])
            if not behind.any():
                return
            advanced = (after["time_q"] - before["time_q"]) + (after["time_r"] - before["time_r"])
            stalled = stalled + 1 if float(np.min(advanced[behind])) < 1.0e-9 else 0
            if stalled >= 3:
                chain = int(np.nonzero(behind)[0][0])
                raise RuntimeError("chain {0} does not advance in time any more ({1} events per launch): "
                                   "collapsing configuration?".format(chain, self._events_per_launch))

    @property
    def timings(self):
        """Host seconds spent advancing the chains (device launches) and in control events (sampling, dumps)."""
        return dict(self._timings)

    @property
    def statistics(self):
        """Event counters of all chains so far (EcmcStats of include/ecmc.h)."""
        return dict(self._statistics)

    @property
    def engine(self):
        """The engine of the first device."""
        return self._engine

    @property
    def engines(self):
        return list(self._engines)

    def update_logging(self) -> None:
        self._logger = logging.getLogger(__name__)
