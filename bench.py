"""bench.py -- ECMC events/sec of the batched Lennard-Jones workload C2 on N B200 GPUs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): 3D Lennard-Jones, N = 1024 particles per chain, 4096 independent chains per
GPU (weak scaling: every rank owns its own 4096 chains and random streams, no data-path collective), cells 12^3,
nearby cells + surplus by exact LJ inversion, all other cells by cell veto. A "step" advances every chain by
`--events` events (one ecmc_run launch). One JSON line is printed by rank 0:

  value      events/s over all ranks, chain state resident in HBM, CUDA events on the launching stream, max over ranks
  e2e        the same step through the host-buffer entry point ecmc_run_from_host: pinned host positions in ->
             H2D -> cell binning -> events -> D2H positions out, host clock around synchronous calls
  roofline   the event kernel against the measured HBM peak (algorithmic bytes per event) -- plus
             "fp64": the same kernel against the measured DFMA rate, which is the pipe that actually binds it
  single_chain  N = 1 only: ns/event of ONE Lennard-Jones chain of 65536 particles (C5, the second half of the metric),
             measured after the timed region on its own engine handle
  cpu_baseline  the unmodified reference (baseline/_ref, CPython) on the host cores, bounded sample; else the C port

`--impl reference` times only the CPU reference arm, on rank 0.
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
    parser.add_argument("--steps", type=int, default=40)
    parser.add_argument("--warmup", type=int, default=3)
    parser.add_argument("--impl", default="ours", choices=["ours", "reference"])
    parser.add_argument("--chains", type=int, default=4096, help="independent chains per GPU")
    parser.add_argument("--particles", type=int, default=1024)
    parser.add_argument("--cells", type=int, default=12)
    parser.add_argument("--events", type=int, default=1024, help="events per chain and step")
    parser.add_argument("--e2e-steps", type=int, default=8)
    parser.add_argument("--cpu-seconds", type=float, default=8.0, help="wall budget of the CPU baseline sample")
    parser.add_argument("--no-cpu-baseline", action="store_true")
    parser.add_argument("--ref-seconds", type=float, default=None,
                        help="--impl reference: wall-clock length of one step (default min(20, 60 / (W + K)) s)")
    parser.add_argument("--no-single-chain", action="store_true", help="skip the C5 single-chain latency leg (N = 1 only)")
    return parser.parse_args()


def workload_config(args, world):
    return {"workload": "C2: 3D Lennard-Jones (prefactor 4, sigma 1), density 0.5, cells %d^3 nl=1, cell-veto far field, "
                        "chain_time 10, beta 1" % args.cells,
            "particles_per_chain": args.particles, "chains_per_gpu": args.chains,
            "events_per_chain_per_step": args.events, "parallelism": "chains sharded over %d GPU(s)" % world,
            "l2": "chain state (particles + cell occupancy) of one GPU = %.0f MB > 126 MB L2"
                  % ((args.chains * args.particles * 32 + args.chains * args.cells ** 3 * 4) / 1e6)}


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
        """NVML directly (what nvidia-smi reads), every 5 ms: the timed region of the default run lasts a quarter of a
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
def reference_sample(args, seconds):
    """The unmodified reference on all host cores, one chain per core, `seconds` of wall clock each."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import configs
    import reference_runner
    from jellyfysh_b200 import workloads
    if not reference_runner.available():
        return None
    cores = os.cpu_count() or 1
    length = float((args.particles / 0.5) ** (1.0 / 3.0))
    # one handler instance per possible surplus particle: the reference raises when the list outgrows its pool
    ini = configs.lennard_jones_ini(args.particles, length, args.cells, chain_time=10.0, surplus_handlers=400)
    positions = workloads.lattice_start(cores, args.particles, args.cells, length)
    rate, processes, events, init_seconds = reference_runner.run(ini, list(positions), warmup_seconds=2.0,
                                                                  budget_seconds=seconds)
    return {"value": rate, "unit": UNIT, "cores": processes, "kind": "reference",
            "sample": "unmodified JeLLyFysh 1.1 (baseline/_ref, CPython %d.%d) single_process_mediator, one chain of the "
                      "same workload per core for %.0f s after 2 s warm-up: %d events; init %.1f s per process not counted"
                      % (sys.version_info[0], sys.version_info[1], seconds, events, init_seconds)}


def _port_worker(job):
    particles, cells, chain, events = job
    from jellyfysh_b200 import workloads
    from oracle import oracle
    builder, length = _PORT_STATE
    positions = workloads.lattice_start(1, particles, cells, length, first_chain=chain)[0]
    chain_object = oracle.OracleChain(builder)
    chain_object.set_positions(positions)
    chain_object.start(stream=chain)
    chain_object.run(max_events=events // 10)
    t0 = time.perf_counter()
    n, _ = chain_object.run(max_events=events)
    return n, time.perf_counter() - t0


_PORT_STATE = None


def port_sample(args, builder, length, seconds):
    """The oracle's C port of the same algorithm on all host cores (fork: the program's tables are inherited)."""
    import multiprocessing
    global _PORT_STATE
    _PORT_STATE = (builder, length)
    cores = os.cpu_count() or 1
    events = max(1000, int(seconds * 2.5e5))  # ~4 us per event and core
    context = multiprocessing.get_context("fork")
    with context.Pool(cores) as pool:
        results = pool.map(_port_worker, [(args.particles, args.cells, c, events) for c in range(cores)])
    rate = sum(n / dt for n, dt in results)
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "oracle/ecmc_oracle.c (plain-C restatement of the reference algorithm), one chain of the same "
                      "workload per core, %d events each after a 10%% warm-up" % events}


def run_reference_arm(args, rank, world):
    """`--impl reference`: the CPU reference on the host cores, rank 0 only. One pool of processes (one chain per core)
    runs the warm-up steps and the K timed steps back to back; a step is a bounded wall-clock segment."""
    if rank != 0:
        return
    steps = args.warmup + args.steps
    seconds = args.ref_seconds if args.ref_seconds else max(0.5, min(20.0, 60.0 / steps))
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import configs
    import reference_runner
    from jellyfysh_b200 import workloads
    cores = os.cpu_count() or 1
    length = float((args.particles / 0.5) ** (1.0 / 3.0))
    if reference_runner.available():
        # one handler instance per possible surplus particle: the reference raises when the list outgrows its pool
        ini = configs.lennard_jones_ini(args.particles, length, args.cells, chain_time=10.0, surplus_handlers=400)
        positions = workloads.lattice_start(cores, args.particles, args.cells, length)
        rates, processes, events, init_seconds = reference_runner.run_segments(
            ini, list(positions), warmup_seconds=0.0, budget_seconds=seconds, segments=steps)
        value = sum(rates[args.warmup:]) / args.steps
        sample = {"value": value, "unit": UNIT, "cores": processes, "kind": "reference",
                  "sample": "unmodified JeLLyFysh 1.1 (baseline/_ref, CPython %d.%d) single_process_mediator, one chain "
                            "of the same workload per core, %d warm-up + %d timed segments of %.1f s: %d events; init "
                            "%.1f s per process not counted"
                            % (sys.version_info[0], sys.version_info[1], args.warmup, args.steps, seconds, events,
                               init_seconds)}
    else:
        # no installed reference on this box: the C port of its algorithm (oracle) instead
        builder, length = workloads.lennard_jones(n_particles=args.particles, cells_per_side=args.cells)
        sample = port_sample(args, builder, length, seconds)
        value = sample["value"]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * seconds, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
            "cpu_baseline": sample,
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


def algorithmic_bytes_per_event(stats, args):
    """active position 8D + occupancy of the K nearby cells 4K + one 32-byte particle record per pair
    candidate + cell-veto target slot and record (4 + 32 when the cell is occupied; counted always) + write-back of the
    active position and time 8D + 16 + two occupancy updates 8."""
    # pair candidates = the targets gathered from the nearby cells and the surplus: the handler calls the reference makes
    pair_candidates = stats["pair_targets"] / stats["events"]
    return 24.0 + 4.0 * 27 + 32.0 * pair_candidates + 36.0 + 40.0 + 8.0, pair_candidates


def algorithmic_flops_per_event(pair_candidates):
    """Operation count of the reference formulae: ~70 fp64 operations per Lennard-Jones pair candidate
    (separation 9, energy 2 x 8, inversion 8, root displacement 7, expovariate, time), ~30 for the veto draw,
    ~6 boundary, ~25 argmin / commit."""
    return 70.0 * pair_candidates + 30.0 + 6.0 + 25.0


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
    return {"workload": "C5: single 3D Lennard-Jones chain, N = 65536, cells 48^3 nl=1, cell-veto far field",
            "ns_per_event": 1e9 * seconds / stats["events"], "events": stats["events"], "unit": "ns/event"}


def run_ours(args, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist
    from jellyfysh_b200 import engine, sharding, workloads

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    builder, length = workloads.lennard_jones(n_particles=args.particles, cells_per_side=args.cells, device=local_rank)
    first_chain, _ = sharding.chain_shard(rank, world, args.chains)
    positions = workloads.lattice_start(args.chains, args.particles, args.cells, length, first_chain=first_chain)
    eng = engine.Engine(builder, n_chains=args.chains, device=local_rank)
    eng.upload_positions(positions)
    eng.start(first_stream=first_chain)
    stream = torch.cuda.ExternalStream(eng.cuda_stream, device=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with ClockSampler(local_rank) as clocks:  # nvidia-smi needs ~0.1 s to start: it runs through the warm-up
        for _ in range(args.warmup):
            eng.run(max_events=args.events)
        eng.sync()
        launches_before = eng.kernel_launches
        kernel_seconds_before = eng.kernel_seconds
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        clocks.open_window()
        start.record(stream)
        for _ in range(args.steps):
            eng.run(max_events=args.events)
        stop.record(stream)
        stop.synchronize()
        barrier()
        clocks.close_window()
        time.sleep(0.05)
    stats = eng.sync()
    elapsed_ms = start.elapsed_time(stop)
    launches = eng.kernel_launches - launches_before
    kernel_seconds = eng.kernel_seconds - kernel_seconds_before

    # ---- end to end through the host-buffer entry point
    pinned_in = torch.from_numpy(positions).pin_memory()
    pinned_out = torch.empty_like(pinned_in).pin_memory()
    host_in, host_out = pinned_in.numpy(), pinned_out.numpy()
    for _ in range(2):  # warm-up: first-touch of the pinned buffers, stream creation
        eng.run_from_host(host_in, first_stream=first_chain, max_events=args.events, out=host_out)
    e2e_launches_before = eng.kernel_launches
    barrier()
    t0 = time.perf_counter()
    e2e_events = 0
    for _ in range(args.e2e_steps):
        _, e2e_stats = eng.run_from_host(host_in, first_stream=first_chain, max_events=args.events, out=host_out)
        e2e_events += e2e_stats["events"]
    barrier()
    e2e_seconds = time.perf_counter() - t0
    # per chain slice: pack, start, events, unpack (ecmc_kernel_launches counts the event kernels)
    e2e_launches = 4 * (eng.kernel_launches - e2e_launches_before)

    # ---- an observable reduced over ranks: pair-separation histogram of all chains, one NCCL all-reduce
    import numpy as _np
    local_histogram = eng.separation_histogram(256, 0.0, length * 3 ** 0.5 / 2.0)
    histogram = sharding.reduce_histogram(local_histogram.astype(_np.int64), device=torch.device("cuda", local_rank))

    # ---- reduce over ranks (NCCL): event counters are summed, times are the slowest rank's
    device = torch.device("cuda", local_rank)
    all_stats = sharding.reduce_counters(stats, device=device)
    total_e2e_events = int(sharding.reduce_histogram([e2e_events], device=device)[0])
    max_ms, max_e2e_seconds, _ = sharding.reduce_max([elapsed_ms, e2e_seconds, kernel_seconds], device=device).tolist()
    total_events = all_stats["events"]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = total_events / (max_ms * 1e-3)
    e2e_value = total_e2e_events / max_e2e_seconds
    peaks, peak_source = measured_peaks()
    bytes_per_event, pair_candidates = algorithmic_bytes_per_event(stats, args)
    # per launch, this rank: algorithmic bytes of one launch / average launch duration (per-launch CUDA events)
    events_per_launch = stats["events"] / launches
    launch_seconds = kernel_seconds / launches
    achieved_gbs = events_per_launch * bytes_per_event / launch_seconds * 1e-9
    traffic = None
    summary_path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    ncu = {}
    if os.path.exists(summary_path):
        with open(summary_path) as handle:
            ncu = json.load(handle)
        traffic = ncu.get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "achieved": achieved_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved_gbs / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_source,
                "kernel": "ecmc::event_kernel<LJ, 0, LJ>", "algorithmic_bytes_per_event": bytes_per_event,
                "events_per_launch": events_per_launch, "launch_ms": 1e3 * launch_seconds,
                "pair_candidates_per_event": pair_candidates,
                "note": "the chain state is L2-resident and the kernel is bound by instruction issue (a dependent chain of "
                        "~1000 warp instructions per event, a quarter of them fp64), not by HBM: see fp64 and "
                        "profiles/"}
    dfma = dfma_peak(local_rank)
    flops_per_event = algorithmic_flops_per_event(pair_candidates)
    achieved_tflops = events_per_launch * flops_per_event / launch_seconds * 1e-12
    fp64 = {"achieved_algorithmic_tflops": achieved_tflops, "peak_dfma_tflops": dfma,
            "frac_algorithmic": None if not dfma else achieved_tflops / dfma,
            "algorithmic_flops_per_event": flops_per_event,
            "ncu_fp64_pipe_utilisation_pct": ncu.get("fp64_pipe_pct"),
            "ncu_issue_slot_utilisation_pct": ncu.get("issue_active_pct"),
            "ncu_warp_instructions_per_event": ncu.get("warp_instructions_per_event"),
            "peak_source": "tools/fp64_peak.cu measured in this run (2 flop per DFMA)"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(positions.nbytes), "d2h_bytes_per_step": int(positions.nbytes) + 96,
                    "steps": args.e2e_steps, "call": "ecmc_run_from_host: pinned host positions -> H2D -> cell binning -> events -> D2H, pipelined over "
                            "chain slices on separate streams"},
            "gpu_launches": int(launches + e2e_launches),
            "roofline": roofline, "fp64": fp64,
            "observable": {"kind": "pair-separation histogram of all chains (ecmc_separation_histogram), summed over "
                                   "ranks by one all-reduce", "bins": 256, "pairs_counted": int(histogram.sum()),
                           "expected_pairs": world * args.chains * args.particles * (args.particles - 1) // 2},
            "event_mix": {k: all_stats[k] for k in ("pair_events", "veto_events", "veto_accepted", "boundary_events",
                                                    "end_of_chain_events", "bound_violations")}}
    if world == 1 and not args.no_single_chain:
        line["single_chain"] = single_chain_latency(local_rank)
    if world == 1 and not args.no_cpu_baseline:
        baseline = reference_sample(args, args.cpu_seconds)
        if baseline is None:
            baseline = port_sample(args, builder, length, args.cpu_seconds)
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
