"""bench.py -- ECMC events/sec of the batched event-chain kernels on N B200 GPUs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default workload C2 (BASELINE.json configs[1], the configuration the metric is quoted on): 3D Lennard-Jones, N = 1024
particles per chain, 4096 independent chains per GPU (weak scaling: every rank owns its own chains and random streams, no
data-path collective), cells 12^3, nearby cells + surplus by exact LJ inversion, all other cells by cell veto. The other
BASELINE configurations are `--workload c1` (shipped hard-disk dipoles), `c3` (Coulomb atoms, `--particles` 64..512),
`c4` (SPC/Fw water, 32 molecules, the oxygen-oxygen histogram of every step all-reduced over the ranks inside the timed
region) and `c5` (one Lennard-Jones chain of 65536 particles; value = 1 / latency). A "step" advances every chain by
`events_per_chain_per_step` events (one ecmc_run launch; C2: 4096 = four sweeps of the 1024 particles, `--events` changes
it). One JSON line is printed by rank 0:

  value      events/s over all ranks, chain state resident in HBM, CUDA events on the launching stream, max over ranks
  e2e        the same steps from HOST buffers: every step takes the configuration from a pinned host buffer (the whole
             configuration crosses the link), bins it into the cells, runs the events and returns the configuration to
             the host -- through ecmc_submit_from_host_sparse (only the particles that moved are written back, by the
             device; "full_copy": ecmc_submit_from_host, "synchronous": the blocking ecmc_run_from_host). The chains
             restart from the start configuration with the same random streams and CONTINUE from step to step
             (ECMC_OPTION_CONTINUE_HOST_STEPS), so a leg runs the events of the device-timed region: --warmup untimed
             steps, then --e2e-steps (default --steps) timed ones; host clock between barriers. Composite objects (c1,
             c4): upload / start / run / download calls, a start of run per step
  roofline   the event kernel against the measured HBM peak: SURVEY.md 8(d)'s algorithmic bytes per event x the events of
             one launch / the launch duration measured in THIS run (CUDA events around every launch);
             "traffic" only when a committed ncu capture of the SAME kernel exists (source named, else null)
  fp64       the same launch against the DFMA rate measured in this run (operation count of the reference formulae)
  single_chain  N = 1 and C2 only: ns/event of ONE Lennard-Jones chain of 65536 particles (C5, the second half of the
             metric), measured after the timed region on its own engine handle
  cpu_baseline  the unmodified reference (baseline/_ref, CPython) on the host cores over the same event window per chain

`--impl reference` times only the CPU reference arm, on rank 0: one chain per core, W + K steps of the same number of
events per chain as the device arm's steps (the same window of every chain's history), the K last ones timed.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ecmc_events_per_sec"
UNIT = "events/s"


def parse_args():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=20)
    parser.add_argument("--warmup", type=int, default=3)
    parser.add_argument("--impl", default="ours", choices=["ours", "reference"])
    parser.add_argument("--workload", default="c2", choices=["c1", "c1n", "c2", "c3", "c4", "c5"])
    parser.add_argument("--chains", type=int, default=None, help="independent chains per GPU (default: the workload's)")
    parser.add_argument("--particles", type=int, default=None, help="C2 / C3: particles per chain; C4: molecules")
    parser.add_argument("--cells", type=int, default=None, help="C2: cells per side")
    parser.add_argument("--events", type=int, default=None, help="events per chain and step")
    parser.add_argument("--e2e-steps", type=int, default=None,
                        help="steps of each end-to-end leg (default: --steps, the same number the device-timed region runs)")
    parser.add_argument("--cpu-seconds", type=float, default=12.0, help="wall budget of the CPU baseline sample")
    parser.add_argument("--no-cpu-baseline", action="store_true")
    parser.add_argument("--ref-seconds", type=float, default=None,
                        help="--impl reference: steps are wall-clock segments of this length instead of event-bounded ones")
    parser.add_argument("--ref-max-seconds", type=float, default=150.0, help="--impl reference: bound of the whole run")
    parser.add_argument("--no-single-chain", action="store_true", help="skip the C5 single-chain latency leg (N = 1, C2)")
    return parser.parse_args()


# ---------------------------------------------------------------------------------------------------------
# workloads: program, start configuration, algorithmic bytes / operations, the reference's INI
# ---------------------------------------------------------------------------------------------------------
class Workload:
    """One BASELINE configuration. Subclasses fill: name, text, chains, events, dimension, nearby (cells read per
    event), charged, composite (root units), and build the device program / the reference job."""
    composite = False
    charged = False

    def __init__(self, args):
        self.args = args

    def config(self, world):
        out = {"workload": self.text, "chains_per_gpu": self.chains, "events_per_chain_per_step": self.events,
               "particles_per_chain": self.particles, "parallelism": "chains sharded over %d GPU(s)" % world}
        return out

    # device side ------------------------------------------------------------------------------------
    def engine(self, device, first_chain):
        """(started Engine, dict of pinned-able host arrays: positions [, charges, roots])"""
        raise NotImplementedError

    def bytes_per_event(self, stats):
        """B_event = 8D (active position) + 4K (occupancy slots of the nearby cells) + n_cand (8D [+ 8 if
        charged]) (target positions) + 4 + 8D (cell-veto target slot + position) + 8D + 16 (write active position + time)
        + 8 (two occupancy updates); n_cand = the pair targets gathered per event, measured in this run. Tables are not
        charged. Also returned: the same count with the 32-byte padded records the device actually moves."""
        d, k = self.dimension, self.nearby
        n_cand = stats["pair_targets"] / max(stats["events"], 1)
        per_target = 8 * d + (8 if self.charged else 0)
        algorithmic = 8 * d + 4 * k + n_cand * per_target + (4 + 8 * d) + (8 * d + 16) + 8
        layout = 32 + 4 * k + n_cand * 32 + (4 + 32) + (32 + 16) + 8
        return algorithmic, layout, n_cand

    def flops_per_event(self, n_cand):
        raise NotImplementedError

    # reference side ---------------------------------------------------------------------------------
    def reference_job(self, cores):
        """(ini text, [positions or None] per process, composites or None) for baseline/reference_runner.py"""
        raise NotImplementedError


class LennardJones(Workload):
    name, dimension, nearby = "C2", 3, 27

    def __init__(self, args):
        super().__init__(args)
        self.particles = args.particles or 1024
        self.cells = args.cells or 12
        self.chains = args.chains or 4096
        # A step = 4096 events per chain = four sweeps of a chain's 1024 particles between two visits of the host. (Rounds 1
        # and 2 measured steps of 1024 events, `--events 1024`: 2.4 ms of kernel per 100.7 MB upload, which eight ranks
        # sharing one host cannot feed -- 5.4e9 events/s end to end on eight GPUs against 1.33e10 device-resident.)
        self.events = args.events or 4096
        self.length = float((self.particles / 0.5) ** (1.0 / 3.0))
        self.text = ("%s: 3D Lennard-Jones (prefactor 4, sigma 1), density 0.5, N = %d, cells %d^3 nl=1 one occupant per cell, "
                     "exact inversion for nearby cells + surplus, cell-veto far field, chain_time 10, beta 1, jittered "
                     "lattice start" % (self.name, self.particles, self.cells))

    def config(self, world):
        out = super().config(world)
        state_mb = (self.chains * self.particles * 32 + self.chains * self.cells ** 3 * 4) / 1e6
        out["l2"] = "chain state (particles + cell occupancy) of one GPU = %.0f MB %s 126 MB L2" % (
            state_mb, ">" if state_mb > 126 else "<")
        return out

    def builder(self, device):
        from jellyfysh_b200 import workloads
        return workloads.lennard_jones(n_particles=self.particles, cells_per_side=self.cells, device=device)[0]

    def engine(self, device, first_chain):
        from jellyfysh_b200 import engine, workloads
        positions = workloads.lattice_start(self.chains, self.particles, self.cells, self.length, first_chain=first_chain)
        eng = engine.Engine(self.builder(device), n_chains=self.chains, device=device)
        eng.upload_positions(positions)
        eng.start(first_stream=first_chain)
        return eng, {"positions": positions}

    def flops_per_event(self, n_cand):
        """Operation count of the reference formulae: ~70 fp64 operations per Lennard-Jones pair candidate
        (separation 9, energy 2 x 8, inversion 8, root displacement 7, expovariate, time), ~30 veto draw, ~6 boundary,
        ~25 argmin / commit."""
        return 70.0 * n_cand + 30.0 + 6.0 + 25.0

    def reference_job(self, cores):
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import configs
        from jellyfysh_b200 import workloads
        # one handler instance per possible surplus particle: the reference raises when the list outgrows its pool
        ini = configs.lennard_jones_ini(self.particles, self.length, self.cells, chain_time=10.0, surplus_handlers=400)
        return ini, list(workloads.lattice_start(cores, self.particles, self.cells, self.length)), None


class SingleChain(LennardJones):
    name = "C5"

    def __init__(self, args):
        args.particles = args.particles or 65536
        args.cells = args.cells or 48
        args.chains = args.chains or 1
        args.events = args.events or 20000
        super().__init__(args)


class CoulombAtoms(Workload):
    name, dimension, nearby, charged = "C3", 3, 27, True

    def __init__(self, args):
        super().__init__(args)
        self.particles = args.particles or 64
        self.chains = args.chains or {64: 4096, 128: 4096, 256: 2048, 512: 1024}.get(self.particles, 1024)
        self.events = args.events or 1000
        self.length = 1.0
        import numpy as np
        self.cells = int(np.ceil((2 * self.particles) ** (1.0 / 3.0)))
        self.text = ("C3: Coulomb atoms (coulomb_atoms/cell_veto.ini shape), N = %d charges +1, L = 1, beta 2, merged-image "
                     "Coulomb (alpha 3.45, cutoffs 6 / 2) bounded by the inverse-power Coulomb bound in the nearby cells, "
                     "cell veto (inner-point estimator, 10 points per side) elsewhere, cells %d^3, uniform random start"
                     % (self.particles, self.cells))

    def engine(self, device, first_chain):
        import numpy as np
        from jellyfysh_b200 import engine, workloads
        builder, _ = workloads.coulomb_atoms(n_particles=self.particles, device=device)
        positions = workloads.uniform_start(self.chains, self.particles, self.length, first_chain=first_chain)
        charges = np.ones((self.chains, self.particles))
        eng = engine.Engine(builder, n_chains=self.chains, device=device)
        eng.upload_positions(positions, charges)
        eng.start(first_stream=first_chain)
        return eng, {"positions": positions, "charges": charges}

    def flops_per_event(self, n_cand):
        """~45 operations per bounded pair candidate (inverse-power Coulomb bound), one merged-image Coulomb derivative
        (2140 operations) per confirmed pair / cell-veto event -- about one per event --, veto, boundary."""
        return 45.0 * n_cand + 2140.0 + 30.0 + 6.0 + 25.0

    def reference_job(self, cores):
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import configs
        # (uniform random start: up to a few dozen particles share a cell at N = 512; one surplus handler per particle is
        # always enough)
        ini = configs.coulomb_atoms_ini(self.particles, [self.cells] * 3, points_per_side=10,
                                        surplus_handlers=max(16, self.particles // 2))
        return ini, [configs.uniform_start(self.particles, 1.0, seed=1000 + k) for k in range(cores)], None


class HardDiskDipoles(Workload):
    name, dimension, nearby, composite = "C1", 2, 9, True

    def __init__(self, args):
        super().__init__(args)
        self.particles = 162
        self.chains = args.chains or 4096
        self.events = args.events or 4000
        self.text = ("C1: shipped hard_disk_dipoles/hard_disk_dipoles_cells.ini: 81 hard-disk dipoles (162 disks, tethered "
                     "pairs), 2D, L = 12.836, 13^2 leaf-level cells, shipped start configuration, every chain its own "
                     "random stream")

    def engine(self, device, first_chain):
        import numpy as np
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import trace_util as tu
        from jellyfysh_b200 import engine
        from jellyfysh_b200.program import ProgramBuilder
        g = tu.load_trace("trace_hard_disk_dipoles")
        builder = tu.dipole_builder_of(g, ProgramBuilder)
        positions = np.tile(g["positions0"], (self.chains, 1, 1))
        roots = np.tile(g["roots0"], (self.chains, 1, 1))
        eng = engine.Engine(builder, n_chains=self.chains, device=device)
        eng.upload_positions(positions)
        eng.upload_roots(roots)
        eng.start(first_stream=first_chain)
        return eng, {"positions": positions, "roots": roots}

    def flops_per_event(self, n_cand):
        """~25 operations per hard-disk pair candidate (separation, discriminant, root), tether lanes, argmin / commit."""
        return 25.0 * n_cand + 2 * 40.0 + 25.0

    def reference_job(self, cores):
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        import configs
        import reference_runner
        roots, leaves = configs.read_pdb_dipoles(reference_runner.REF_ROOT)
        return configs.hard_disk_dipoles_cells_ini(reference_runner.REF_ROOT), [None] * cores, [(roots, leaves)] * cores


class HardDiskDipolesNoCells(HardDiskDipoles):
    """The no-cell sibling of C1: every other disk is a candidate of every event, general velocities."""
    name, nearby = "C1n", 0

    def __init__(self, args):
        super().__init__(args)
        self.events = args.events or 1000
        self.text = ("C1n: shipped hard_disk_dipoles/hard_disk_dipoles.ini: the same 81 hard-disk dipoles without a cell "
                     "system (160 hard-disk candidates + the tether per event) and with general velocities (the "
                     "sequential-direction end of chain rotates the velocity by 20 degrees), shipped start configuration, "
                     "every chain its own random stream")

    def engine(self, device, first_chain):
        import numpy as np
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import trace_util as tu
        from jellyfysh_b200 import engine
        from jellyfysh_b200.program import ProgramBuilder
        g = tu.load_trace("trace_hard_disk_dipoles_sequential")
        builder = tu.sequential_dipole_builder_of(g, ProgramBuilder)
        builder.program.chain_time = 6.0  # the shipped value (the committed trace was recorded with shorter chains)
        positions = np.tile(g["positions0"], (self.chains, 1, 1))
        roots = np.tile(g["roots0"], (self.chains, 1, 1))
        eng = engine.Engine(builder, n_chains=self.chains, device=device)
        eng.upload_positions(positions)
        eng.upload_roots(roots)
        eng.start(first_stream=first_chain)
        return eng, {"positions": positions, "roots": roots}

    def reference_job(self, cores):
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        import configs
        import reference_runner
        # the shipped file reads its start configuration through PdbInputHandler: where MDAnalysis is not installed the
        # (spawned) reference processes find jellyfysh_b200's .pdb reader under that name
        try:
            import MDAnalysis  # noqa: F401
        except ImportError:
            shims = os.path.join(ROOT, "jellyfysh_b200", "shims")
            os.environ["PYTHONPATH"] = shims + os.pathsep + os.environ.get("PYTHONPATH", "")
            if shims not in sys.path:
                sys.path.append(shims)  # multiprocessing's spawn hands the parent's sys.path to the children
        return configs.hard_disk_dipoles_ini(reference_runner.REF_ROOT), [None] * cores, None


class Water(Workload):
    name, dimension, nearby, charged, composite = "C4", 3, 125, True, True

    def __init__(self, args):
        super().__init__(args)
        self.molecules = args.particles or 32
        self.particles = 3 * self.molecules
        # one wave of the kernel's CTAs: sixteen chains (one per warp) on every one of the 148 SMs; with at most eight
        # chains per SM (--chains 1024, the size of rounds 1 and 2) the kernel runs CTAs of eight warps
        self.chains = args.chains or 2368
        self.events = args.events or 500
        self.text = ("C4: SPC/Fw water (water/coulomb_cell_veto_lj_inverted.ini), %d molecules, L = 10, beta 1.679, "
                     "merged-image Coulomb between molecules (inverse-power bound nearby, cell veto elsewhere), Lennard-Jones "
                     "between oxygens, harmonic bonds + bending, inside-first lifting, cells 6^3 nl=2; replica sweep: the "
                     "oxygen-oxygen separation histogram (1000 bins) of every step is all-reduced over the ranks inside "
                     "the timed region" % self.molecules)

    def engine(self, device, first_chain):
        import numpy as np
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import configs
        import trace_util as tu
        from jellyfysh_b200 import engine
        from jellyfysh_b200.program import ProgramBuilder
        g = dict(tu.load_trace("trace_water"))
        g["meta_n"] = np.asarray(self.particles)
        builder = tu.water_builder_of(g, ProgramBuilder)
        roots = np.empty((self.chains, self.molecules, 3))
        leaves = np.empty((self.chains, self.particles, 3))
        for c in range(self.chains):
            r, l = configs.water_start(self.molecules, 10.0, seed=first_chain + c)
            roots[c], leaves[c] = r, l.reshape(-1, 3)
        charges = np.tile([0.41, -0.82, 0.41], (self.chains, self.molecules))
        eng = engine.Engine(builder, n_chains=self.chains, device=device)
        eng.upload_positions(leaves, charges)
        eng.upload_roots(roots)
        eng.start(first_stream=first_chain)
        return eng, {"positions": leaves, "charges": charges, "roots": roots}

    def bytes_per_event(self, stats):
        """As Workload.bytes_per_event with composite targets: a gathered target is a molecule, three charged leaves."""
        n_cand = stats["pair_targets"] / max(stats["events"], 1)
        algorithmic = 24 + 4 * self.nearby + n_cand * 3 * 32 + (4 + 3 * 32) + (24 + 24 + 16) + 8
        return algorithmic, algorithmic, n_cand

    def flops_per_event(self, n_cand):
        """Three bounded leaf-leaf candidates per gathered molecule (~45 operations each), ~10 intramolecular / LJ factors,
        and the out-state of a composite pair: up to nine merged-image Coulomb derivatives (2140 each; ~3 per event)."""
        return 3 * 45.0 * n_cand + 10 * 60.0 + 3 * 2140.0 + 60.0

    def reference_job(self, cores):
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        import configs
        import reference_runner
        starts = [configs.water_start(self.molecules, 10.0, seed=k) for k in range(cores)]
        ini = configs.water_ini(reference_runner.REF_ROOT, n_molecules=self.molecules, number_trials=200)
        return ini, [None] * cores, [(r, l.reshape(self.molecules, 3, 3)) for r, l in starts]


WORKLOADS = {"c1": HardDiskDipoles, "c1n": HardDiskDipolesNoCells, "c2": LennardJones, "c3": CoulombAtoms, "c4": Water, "c5": SingleChain}


# ---------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.samples = []  # (arrival time, fields)
        self.process = None
        self.device_index = device_index
        self.window = [None, None]

    def open_window(self):
        self.window[0] = time.perf_counter()

    def close_window(self):
        self.window[1] = time.perf_counter()

    def _poll_nvml(self):
        """NVML directly (what nvidia-smi reads), every 5 ms: the timed region of the default run lasts a fraction of a
        second, in which a `nvidia-smi -lms` loop delivers only a handful of lines."""
        import pynvml
        names = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))
        while not self._stop.is_set():
            try:
                clock = pynvml.nvmlDeviceGetClockInfo(self._nvml_handle, pynvml.NVML_CLOCK_SM)
                reasons = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self._nvml_handle)
            except pynvml.NVMLError:
                break
            fields = [str(clock), str(self._nvml_max)] + ["Active" if reasons & bit else "Not Active" for bit, _ in names]
            self.samples.append((time.perf_counter(), fields))
            self._stop.wait(0.005)

    def __enter__(self):
        self._stop = threading.Event()
        self.source = "nvidia-smi"
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml_handle = pynvml.nvmlDeviceGetHandleByIndex(self.device_index)
            self._nvml_max = pynvml.nvmlDeviceGetMaxClockInfo(self._nvml_handle, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self._nvml_handle)  # raises where it is not supported
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            self.source = "nvml"
            return self
        except Exception:  # noqa: BLE001 - fall back to the nvidia-smi loop
            pass
        try:
            self.process = subprocess.Popen(["nvidia-smi", "-i", str(self.device_index), "--query-gpu=" + self.FIELDS,
                                             "--format=csv,noheader,nounits", "-lms", "20"],
                                            stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.process = None
        return self

    def _read(self):
        for line in self.process.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6 and parts[0].isdigit():
                self.samples.append((time.perf_counter(), parts))

    def __exit__(self, *exc):
        self._stop.set()
        if self.process is not None:
            self.process.terminate()
            try:
                self.process.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.process.kill()

    def summary(self):
        begin, end = self.window
        inside = [fields for stamp, fields in self.samples if begin is not None and begin <= stamp <= (end or stamp)]
        # nvidia-smi reports with a delay: if the timed region was shorter than its period, take the nearest samples
        samples = inside or [fields for _, fields in self.samples[-3:]]
        if not samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        clocks = sorted(int(s[0]) for s in samples)
        reasons = set()
        for s in samples:
            for name, value in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if value.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": clocks[len(clocks) // 2], "sm_max_mhz": int(samples[0][1]), "reasons": sorted(reasons),
                "samples": len(samples), "samples_inside_timed_region": len(inside), "source": self.source}


# ---------------------------------------------------------------------------------------------------------
# CPU baselines
# ---------------------------------------------------------------------------------------------------------
def interpreter():
    return "CPython %d.%d" % sys.version_info[:2]


def reference_by_events(workload, warmup, steps, max_seconds):
    """The unmodified reference on all host cores, one chain per core: `warmup` + `steps` segments of the workload's
    events-per-step each -- the same window of every chain's history as the device arm's steps -- within max_seconds of
    wall clock (fewer timed segments if the budget ends first). None when baseline/_ref is not installed."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import reference_runner
    if not reference_runner.available():
        return None
    cores = os.cpu_count() or 1
    ini, positions, composites = workload.reference_job(cores)
    rate, processes, events, init_seconds, completed = reference_runner.run_event_segments(
        ini, positions, workload.events, warmup + steps, warmup, max_seconds, composites=composites)
    window = "events %d..%d of every chain" % (warmup * workload.events, (warmup + completed) * workload.events)
    return {"value": rate, "unit": UNIT, "cores": processes, "kind": "reference", "timed_steps": completed,
            "event_window_per_chain": [warmup * workload.events, (warmup + completed) * workload.events],
            "sample": "unmodified JeLLyFysh 1.1 (baseline/_ref, %s) single_process_mediator, one chain of the same "
                      "workload per core, %d warm-up + %d timed steps of %d events per chain (%s): %d events; init %.1f s "
                      "per process not counted" % (interpreter(), warmup, completed, workload.events, window, events,
                                                   init_seconds)}


def reference_by_seconds(workload, warmup, steps, seconds):
    """The same with steps that are wall-clock segments (`--ref-seconds`)."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import reference_runner
    if not reference_runner.available():
        return None
    cores = os.cpu_count() or 1
    ini, positions, composites = workload.reference_job(cores)
    rates, processes, events, init_seconds = reference_runner.run_segments(
        ini, positions, warmup_seconds=0.0, budget_seconds=seconds, segments=warmup + steps, composites=composites)
    value = sum(rates[warmup:]) / steps
    return {"value": value, "unit": UNIT, "cores": processes, "kind": "reference",
            "sample": "unmodified JeLLyFysh 1.1 (baseline/_ref, %s) single_process_mediator, one chain of the same workload "
                      "per core, %d warm-up + %d timed segments of %.1f s: %d events; init %.1f s per process not counted"
                      % (interpreter(), warmup, steps, seconds, events, init_seconds)}


def _port_worker(job):
    particles, cells, chain, warm_events, events = job
    from jellyfysh_b200 import workloads
    from oracle import oracle
    builder, length = _PORT_STATE
    positions = workloads.lattice_start(1, particles, cells, length, first_chain=chain)[0]
    chain_object = oracle.OracleChain(builder)
    chain_object.set_positions(positions)
    chain_object.start(stream=chain)
    chain_object.run(max_events=warm_events)
    t0 = time.perf_counter()
    n, _ = chain_object.run(max_events=events)
    return n, time.perf_counter() - t0


_PORT_STATE = None


def port_sample(workload, warmup, steps):
    """No installed reference on this box: the oracle's C port of the same algorithm (Lennard-Jones workloads only) on
    all host cores over the same event window per chain (fork: the program's tables are inherited)."""
    import multiprocessing
    global _PORT_STATE
    if not isinstance(workload, LennardJones):
        raise RuntimeError("baseline/_ref is not installed and the C port only covers the Lennard-Jones workloads")
    from jellyfysh_b200 import workloads
    _PORT_STATE = workloads.lennard_jones(n_particles=workload.particles, cells_per_side=workload.cells, veto=True)
    cores = os.cpu_count() or 1
    context = multiprocessing.get_context("fork")
    with context.Pool(cores) as pool:
        results = pool.map(_port_worker, [(workload.particles, workload.cells, c, warmup * workload.events,
                                           steps * workload.events) for c in range(cores)])
    rate = sum(n / dt for n, dt in results)
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
            "event_window_per_chain": [warmup * workload.events, (warmup + steps) * workload.events],
            "sample": "oracle/ecmc_oracle.c (plain-C restatement of the reference algorithm), one chain of the same "
                      "workload per core, %d events each after %d warm-up events" % (steps * workload.events,
                                                                                     warmup * workload.events)}


def run_reference_arm(args, rank, world):
    """`--impl reference`: the CPU reference on the host cores, rank 0 only."""
    if rank != 0:
        return
    workload = WORKLOADS[args.workload](args)
    t0 = time.perf_counter()
    if args.ref_seconds:
        sample = reference_by_seconds(workload, args.warmup, args.steps, args.ref_seconds)
    else:
        sample = reference_by_events(workload, args.warmup, args.steps, args.ref_max_seconds)
    if sample is None:
        sample = port_sample(workload, args.warmup, args.steps)
    wall = time.perf_counter() - t0
    value = sample["value"]
    timed_steps = sample.get("timed_steps", args.steps)
    # one step = events_per_chain_per_step events of every one of the `cores` chains
    ms_per_step = 1e3 * args.ref_seconds if args.ref_seconds else 1e3 * workload.events * sample["cores"] / value
    config = workload.config(world)
    config["chains_timed"] = sample["cores"]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": timed_steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": sample, "wall_seconds": wall,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as handle:
            return json.load(handle), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def dfma_peak(device):
    library = os.path.join(ROOT, "tools", "libfp64_peak.so")
    if not os.path.exists(library):
        return None
    lib = ctypes.CDLL(library)
    lib.measure_dfma_tflops.restype = ctypes.c_double
    lib.measure_dfma_tflops.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    sms = ctypes.c_int(0)
    value = lib.measure_dfma_tflops(device, ctypes.byref(sms))
    return value if value > 0 else None


def committed_capture(kernel_name):
    """The ncu summary under profiles/ whose "kernel_name" is the kernel this run launched (ecmc_kernel_name), if any: an
    annotation from a committed capture of the same kernel, NOT a measurement of this run."""
    folder = os.path.join(ROOT, "profiles")
    best = None
    for name in sorted(os.listdir(folder)) if os.path.isdir(folder) else []:
        if not name.endswith("ncu_summary.json"):
            continue
        try:
            with open(os.path.join(folder, name)) as handle:
                summary = json.load(handle)
        except (OSError, ValueError):
            continue
        if summary.get("kernel_name") == kernel_name:
            best = (name, summary)  # the last one in name order = the latest round
    return best


def single_chain_latency(device):
    """The second half of BASELINE.json's metric, "ns/event single chain" (SURVEY.md 8d, C5): ONE Lennard-Jones chain of
    65536 particles in 48^3 cells, same potential, density and far field as C2. One warp on the whole GPU: the number is
    the latency of the dependent instruction chain of an event, not a throughput. Device time of the event kernel
    (CUDA events on the engine's stream) after one warm-up launch."""
    from jellyfysh_b200 import engine, workloads
    n, cells, events = 65536, 48, 100000
    builder, length = workloads.lennard_jones(n_particles=n, cells_per_side=cells, device=device)
    start = workloads.lattice_start(1, n, cells, length)
    with engine.Engine(builder, n_chains=1, device=device) as eng:
        eng.upload_positions(start)
        eng.start(first_stream=0)
        eng.run(max_events=events)
        eng.sync()
        before = eng.kernel_seconds
        for _ in range(2):
            eng.run(max_events=events)
        stats = eng.sync()
        seconds = eng.kernel_seconds - before
        kernel = eng.kernel_name()
    return {"workload": "C5: single 3D Lennard-Jones chain, N = 65536, cells 48^3 nl=1, cell-veto far field",
            "ns_per_event": 1e9 * seconds / stats["events"], "events": stats["events"], "unit": "ns/event",
            "event_window": [events, 3 * events], "kernel": kernel}


def run_ours(args, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist
    from jellyfysh_b200 import engine, sharding

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    workload = WORKLOADS[args.workload](args)
    first_chain, _ = sharding.chain_shard(rank, world, workload.chains)
    eng, host = workload.engine(local_rank, first_chain)
    stream = torch.cuda.ExternalStream(eng.cuda_stream, device=device)
    kernel_name = eng.kernel_name()
    water = isinstance(workload, Water)
    oo_histogram = np.zeros(1000, dtype=np.uint64)
    oo_total = torch.zeros(1000, dtype=torch.int64, device=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        eng.run(max_events=workload.events)
        if water:
            # the replica sweep's estimator: oxygen-oxygen separations of this rank's chains on the device,
            # then ONE all-reduce of the 1000-bin histogram over the ranks, every step
            oo_histogram[:] = 0
            eng.separation_histogram(1000, 2.0, 7.0, out=oo_histogram, first=1, stride=3)
            reduced = sharding.reduce_histogram(oo_histogram.astype(np.int64), device=device)
            oo_total.add_(torch.as_tensor(np.asarray(reduced), device=device))

    with ClockSampler(local_rank) as clocks:  # the sampler needs ~0.1 s to start: it runs through the warm-up
        for _ in range(args.warmup):
            step()
        warm_stats = eng.sync()
        launches_before = eng.kernel_launches
        kernel_seconds_before = eng.kernel_seconds
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        clocks.open_window()
        wall0 = time.perf_counter()
        start.record(stream)
        for _ in range(args.steps):
            step()
        stop.record(stream)
        stop.synchronize()
        barrier()
        wall_seconds = time.perf_counter() - wall0
        clocks.close_window()
        time.sleep(0.05)
    stats = eng.sync()
    # C4: the step holds host-synchronous pieces (histogram download, all-reduce), so its time is the host clock between
    # the two barriers; otherwise the CUDA events on the launching stream
    elapsed_ms = 1e3 * wall_seconds if water else start.elapsed_time(stop)
    launches = eng.kernel_launches - launches_before
    kernel_seconds = eng.kernel_seconds - kernel_seconds_before
    _, surplus = eng.cells() if not workload.composite else (None, [])
    mean_surplus = float(np.mean([len(s) for s in surplus])) if surplus else None

    # ---- end to end from host buffers ----------------------------------------------------------------------------
    pinned = {name: [torch.from_numpy(np.ascontiguousarray(array)).pin_memory() for _ in range(2)]
              for name, array in host.items()}
    for name, array in host.items():
        pinned[name][0].numpy()[...] = array
    h2d = sum(array.nbytes for array in host.values())
    d2h = host["positions"].nbytes + (host["roots"].nbytes if "roots" in host else 0) + 96
    total_chains = world * workload.chains

    pipelined = not workload.composite
    charges = pinned["charges"][0].numpy() if "charges" in pinned else None

    def e2e_step(k):
        """Composite objects: configuration in from the pinned buffer step k - 1 wrote, out into the other one; a start of
        run with fresh random streams per step (upload, start, run, download calls)."""
        source, target = k % 2, (k + 1) % 2
        eng.upload_positions(pinned["positions"][source].numpy(), charges)
        eng.upload_roots(pinned["roots"][source].numpy())
        eng.start(first_stream=first_chain + (k + 1) * total_chains)
        eng.run(max_events=workload.events)
        step_stats = eng.sync()
        pinned["positions"][target].numpy()[...] = eng.download_positions()
        pinned["roots"][target].numpy()[...] = eng.download_roots()
        return step_stats

    def link_rate(to_device, together=False):
        """GB/s of one pinned-host <-> device copy of the positions buffer alone (CUDA events): what bounds an e2e step.
        together: every repetition starts at a barrier of all ranks and the MEAN is returned -- what the box's host
        memory and PCIe complex sustain per rank when all N GPUs copy at once; otherwise the best of three."""
        host_buffer = pinned["positions"][0]
        device_buffer = torch.empty_like(host_buffer, device=device)
        begin, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        rates = []
        for _ in range(4 if together else 3):
            if together:
                barrier()
            begin.record()
            for _ in range(3 if together else 1):  # (several copies back to back: the ranks overlap for most of them)
                if to_device:
                    device_buffer.copy_(host_buffer, non_blocking=True)
                else:
                    host_buffer.copy_(device_buffer, non_blocking=True)
            end.record()
            end.synchronize()
            rates.append((3 if together else 1) * host_buffer.numel() * 8 / (begin.elapsed_time(end) * 1e-3) * 1e-9)
        return sum(rates[1:]) / len(rates[1:]) if together else max(rates)

    h2d_rate = link_rate(True)
    d2h_rate = link_rate(False)
    shared_h2d_rate = link_rate(True, together=True)
    barrier()
    # (the probe copies buffer 0 to the device and back: its content is unchanged)
    e2e_launches_before = eng.kernel_launches

    def host_leg(mode):
        """One end-to-end leg over the SAME events as the device-timed region: the chains restart from the start
        configuration with the same random streams, and every step takes the configuration from the pinned host buffer the
        step before returned it in, continuing the chains (ECMC_OPTION_CONTINUE_HOST_STEPS: the lifting state stays on the
        device, the cell occupancy is rebuilt from the uploaded configuration). --warmup untimed steps, then --e2e-steps
        timed ones. mode: "blocking" (ecmc_run_from_host per step), "full" (ecmc_submit_from_host x steps + ecmc_wait, the
        whole configuration copied back), "sparse" (ecmc_submit_from_host_sparse in place on one buffer)."""
        buffers = [engine.pinned_array(host["positions"].shape)] if mode == "sparse" else \
                  [pinned["positions"][0].numpy(), pinned["positions"][1].numpy()]
        buffers[0][...] = host["positions"]
        eng.set_option(eng.OPTION_CONTINUE_HOST_STEPS, 1)
        eng.upload_positions(host["positions"], charges)
        eng.start(first_stream=first_chain)
        n = len(buffers)

        def submit(k):
            if mode == "blocking":
                return eng.run_from_host(buffers[k % n], charges, max_events=workload.events, out=buffers[(k + 1) % n])[1]
            eng.submit_from_host(buffers[k % n], charges, max_events=workload.events, out=buffers[(k + 1) % n],
                                 sparse=mode == "sparse")
            return None
        for k in range(args.warmup):
            submit(k)
        if mode != "blocking":
            eng.wait()
        written_before = eng.host_bytes_written
        launches_before = eng.kernel_launches
        barrier()
        t0 = time.perf_counter()
        events, targets = 0, 0
        for k in range(args.warmup, args.warmup + args.e2e_steps):
            step_stats = submit(k)
            if step_stats is not None:
                events += step_stats["events"]
                targets += step_stats["pair_targets"]
        if mode != "blocking":
            wait_stats = eng.wait()
            events, targets = wait_stats["events"], wait_stats["pair_targets"]
        barrier()
        seconds = time.perf_counter() - t0
        eng.set_option(eng.OPTION_CONTINUE_HOST_STEPS, 0)
        return {"seconds": seconds, "events": events, "targets": targets, "launches": eng.kernel_launches - launches_before,
                "bytes_written": (eng.host_bytes_written - written_before) / args.e2e_steps}

    sparse_seconds, sparse_events, sparse_targets, sparse_bytes = 0.0, 0, 0, 0
    fused = pipelined and "lj_spec_kernel" in kernel_name
    if pipelined:
        blocking_leg = host_leg("blocking")
        full_leg = host_leg("full")
        sparse_leg = host_leg("sparse")
        sync_e2e_seconds, sync_e2e_events = blocking_leg["seconds"], blocking_leg["events"]
        e2e_seconds, e2e_events, e2e_targets = full_leg["seconds"], full_leg["events"], full_leg["targets"]
        sparse_seconds, sparse_events, sparse_targets = sparse_leg["seconds"], sparse_leg["events"], sparse_leg["targets"]
        sparse_bytes = sparse_leg["bytes_written"]
        # staged steps, per chain slice: pack, start, events, unpack (ecmc_kernel_launches counts the event kernels); a
        # sparse step of a Lennard-Jones / cell-veto program is one launch per chain slice, of any other program four
        e2e_launches = 4 * (blocking_leg["launches"] + full_leg["launches"]) + (1 if fused else 4) * sparse_leg["launches"]
    else:
        e2e_warmup = 2
        for k in range(e2e_warmup):  # first touch of the pinned buffers
            e2e_step(k)
        barrier()
        t0 = time.perf_counter()
        e2e_events, e2e_targets = 0, 0
        for k in range(e2e_warmup, e2e_warmup + args.e2e_steps):
            step_stats = e2e_step(k)
            e2e_events += step_stats["events"]
            e2e_targets += step_stats["pair_targets"]
        barrier()
        sync_e2e_seconds = time.perf_counter() - t0
        sync_e2e_events = e2e_events
        e2e_seconds = sync_e2e_seconds
        e2e_launches = 4 * (eng.kernel_launches - e2e_launches_before)

    # ---- an observable reduced over ranks: pair-separation histogram of all chains, one NCCL all-reduce
    if water:
        histogram = oo_total.cpu().numpy()
        observable = {"kind": "oxygen-oxygen separation histogram on [2, 7] (ecmc_separation_histogram_subset), accumulated "
                              "over the warm-up and timed steps, all-reduced over the ranks every step inside the timed region",
                      "bins": 1000, "samples": int(histogram.sum())}
    else:
        box = float(eng._builder.program.system_length)
        local_histogram = eng.separation_histogram(256, 0.0, box * workload.dimension ** 0.5 / 2.0)
        histogram = sharding.reduce_histogram(local_histogram.astype(np.int64), device=device)
        observable = {"kind": "pair-separation histogram of all chains (ecmc_separation_histogram), summed over ranks by one "
                              "all-reduce after the timed region", "bins": 256, "pairs_counted": int(histogram.sum()),
                      "expected_pairs": world * workload.chains * workload.particles * (workload.particles - 1) // 2}

    # ---- reduce over ranks (NCCL): event counters are summed, times are the slowest rank's
    all_stats = sharding.reduce_counters(stats, device=device)
    total_e2e_events, total_sync_events, total_sparse_events = [int(v) for v in sharding.reduce_histogram(
        [e2e_events, sync_e2e_events, sparse_events], device=device)]
    max_ms, max_e2e_seconds, max_sync_seconds, max_sparse_seconds = sharding.reduce_max(
        [elapsed_ms, e2e_seconds, sync_e2e_seconds, sparse_seconds], device=device).tolist()
    total_events = all_stats["events"]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = total_events / (max_ms * 1e-3)
    e2e_value = total_e2e_events / max_e2e_seconds
    peaks, peak_source = measured_peaks()
    bytes_per_event, layout_bytes, n_cand = workload.bytes_per_event(stats)
    # per launch, this rank: algorithmic bytes of one launch / average launch duration (per-launch CUDA events)
    events_per_launch = stats["events"] / launches
    launch_seconds = kernel_seconds / launches
    achieved_gbs = events_per_launch * bytes_per_event / launch_seconds * 1e-9
    capture = committed_capture(kernel_name)
    traffic, traffic_source = None, None
    if capture is not None:
        traffic = capture[1].get("dram_bytes_per_launch")
        traffic_source = ("profiles/%s: ncu --set full capture of the same kernel (%s), %d events per launch -- an annotation "
                          "from a committed capture, not measured in this run" %
                          (capture[0], kernel_name, int(capture[1].get("events_per_launch", 0))))
    roofline = {"bound": "hbm", "achieved": achieved_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved_gbs / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_source,
                "peak_source": peak_source, "kernel": kernel_name,
                "algorithmic_bytes_per_event": bytes_per_event, "layout_bytes_per_event": layout_bytes,
                "bytes_formula": "SURVEY.md 8(d): 8D + 4K + n_cand (8D [+8 charged]) + (4 + 8D) + (8D + 16) + 8, n_cand = "
                                 "pair targets gathered per event in the timed region; layout = the same with the 32-byte "
                                 "records the device moves",
                "events_per_launch": events_per_launch, "launch_ms": 1e3 * launch_seconds,
                "pair_targets_per_event": n_cand,
                "note": "the chain state the events touch is L2-resident (DRAM traffic is a few % of the algorithmic bytes) and "
                        "the kernel is bound by instruction issue and the latency of each chain's dependent instruction "
                        "sequence, not by HBM: see fp64 and profiles/"}
    dfma = dfma_peak(local_rank)
    flops_per_event = workload.flops_per_event(n_cand)
    achieved_tflops = events_per_launch * flops_per_event / launch_seconds * 1e-12
    fp64 = {"achieved_algorithmic_tflops": achieved_tflops, "peak_dfma_tflops": dfma,
            "frac_algorithmic": None if not dfma else achieved_tflops / dfma,
            "algorithmic_flops_per_event": flops_per_event,
            "peak_source": "tools/fp64_peak.cu measured in this run (2 flop per DFMA)"}
    if capture is not None:
        fp64["committed_capture"] = {"source": "profiles/" + capture[0], "kernel": kernel_name,
                                     "fp64_pipe_utilisation_pct": capture[1].get("fp64_pipe_pct"),
                                     "issue_slot_utilisation_pct": capture[1].get("issue_active_pct"),
                                     "warp_instructions_per_event": capture[1].get("warp_instructions_per_event"),
                                     "note": "ncu figures of a committed capture of the same kernel, not of this run"}
    config = workload.config(world)
    config["event_window_per_chain"] = {"before_timed_region": args.warmup * workload.events,
                                        "timed_region": [args.warmup * workload.events,
                                                         (args.warmup + args.steps) * workload.events],
                                        "pair_targets_per_event_warmup": warm_stats["pair_targets"] / max(warm_stats["events"], 1),
                                        "pair_targets_per_event_timed": n_cand,
                                        "mean_surplus_particles_at_end": mean_surplus}
    full_copy = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                 "steps": args.e2e_steps, "ms_per_step": 1e3 * max_e2e_seconds / args.e2e_steps,
                 "host_gb_per_s_per_rank": (h2d + d2h) * args.e2e_steps / max_e2e_seconds * 1e-9,
                 "pair_targets_per_event": e2e_targets / max(e2e_events, 1),
                 "call": ("ecmc_upload_positions + ecmc_upload_roots + ecmc_start + ecmc_run + ecmc_sync + "
                          "ecmc_download_positions + ecmc_download_roots per step" if workload.composite else
                          "ecmc_submit_from_host x steps + ecmc_wait: per step pinned host configuration -> H2D -> cell "
                          "binning -> events -> D2H of the whole configuration on the streams of the chain slices; the host "
                          "does not block between steps, the device orders them slice by slice") +
                         ("; step k + 1 starts from the configuration step k returned, with fresh random streams"
                          if workload.composite else
                          "; the chains restart from the start configuration and run the same events as the device-timed "
                          "region (--warmup untimed steps first): step k + 1 uploads the configuration step k returned and "
                          "continues the chains, their lifting state stays on the device (ECMC_OPTION_CONTINUE_HOST_STEPS)")}
    e2e = dict(full_copy)
    if pipelined and max_sparse_seconds > 0.0:
        # the headline: the same steps, the configuration returned through the sparse write-back
        e2e = {"value": total_sparse_events / max_sparse_seconds, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(sparse_bytes) + 96, "steps": args.e2e_steps,
               "ms_per_step": 1e3 * max_sparse_seconds / args.e2e_steps,
               "host_gb_per_s_per_rank": (h2d + sparse_bytes + 96) * args.e2e_steps / max_sparse_seconds * 1e-9,
               "pair_targets_per_event": sparse_targets / max(sparse_events, 1),
               "call": "ecmc_submit_from_host_sparse x steps + ecmc_wait, in place on one pinned buffer: per step the whole "
                       "host configuration crosses the link to the device" +
                       (" -- read by the event kernel itself, which bins it into the cells, runs the events and writes every "
                        "position it changes through to the host buffer (one launch per chain slice)" if fused else
                        " (H2D copy) -> cell binning -> events -> the device writes the coordinates of the particles that "
                        "moved straight into the host buffer") +
                       " (d2h_bytes_per_step, measured: ecmc_host_bytes_written); the buffer then is the complete "
                       "configuration again and is what step k + 1 reads; the chains restart from the start configuration "
                       "and run the same events as the device-timed region (--warmup untimed steps first), continuing from "
                       "step to step (ECMC_OPTION_CONTINUE_HOST_STEPS: lifting state on the device, cell occupancy rebuilt "
                       "from every uploaded configuration); the host does not block between steps",
               "full_copy": full_copy}
    e2e["link_gb_per_s"] = {"h2d": h2d_rate, "d2h": d2h_rate, "h2d_all_ranks_at_once_per_rank": shared_h2d_rate,
                            "note": "one pinned copy of the positions buffer alone on rank 0, best of 3: h2d_bytes / h2d rate "
                                    "is the floor of a step before its last slice can start"}
    e2e["synchronous"] = {"value": total_sync_events / max_sync_seconds, "ms_per_step": 1e3 * max_sync_seconds / args.e2e_steps,
                          "call": "the blocking form (ecmc_run_from_host / the upload-start-run-download calls), full copy "
                                  "both ways, one step at a time"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config,
            "clocks": clocks.summary(),
            "e2e": e2e,
            "gpu_launches": int(launches + e2e_launches),
            "roofline": roofline, "fp64": fp64, "observable": observable,
            "event_mix": {k: all_stats[k] for k in ("pair_events", "veto_events", "veto_accepted", "boundary_events",
                                                    "end_of_chain_events", "bound_violations")}}
    if world == 1 and isinstance(workload, SingleChain):
        line["single_chain"] = {"workload": workload.text, "ns_per_event": 1e9 * kernel_seconds / stats["events"],
                                "events": stats["events"], "unit": "ns/event", "kernel": kernel_name}
    elif world == 1 and args.workload == "c2" and not args.no_single_chain:
        line["single_chain"] = single_chain_latency(local_rank)
    eng.close()
    if world == 1 and not args.no_cpu_baseline:
        # the same window of every chain's history, as far as the wall budget allows
        # (C5: building the reference's cell-veto tables for 48^3 cells in Python takes longer than any bounded sample;
        # the C port of its algorithm stands in and says so)
        baseline = None
        if not isinstance(workload, SingleChain):
            try:
                baseline = reference_by_events(workload, args.warmup, args.steps, args.cpu_seconds)
            except RuntimeError as error:
                if "warm-up" not in str(error):
                    raise
                # the reference cannot even finish the warm-up steps within the budget (C1n: 161 handler calls per event
                # in Python): time it over wall-clock segments from its start instead, and say so
                baseline = reference_by_seconds(workload, 1, 2, args.cpu_seconds / 3.0)
                baseline["sample"] += " (the event window of the device arm does not fit the wall budget of this leg)"
        if baseline is None:
            baseline = port_sample(workload, args.warmup, min(args.steps, 4))
        line["cpu_baseline"] = baseline
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_RESULT_FD = None


def emit(line):
    """The ONE JSON line of the contract goes to the real stdout; everything else a library prints there (NCCL's
    version banner, for instance) was diverted to stderr by main()."""
    text = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(text.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, text)


def main():
    global _RESULT_FD
    args = parse_args()
    if args.e2e_steps is None:
        args.e2e_steps = args.steps
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)  # file-descriptor level: also catches what native libraries print
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
