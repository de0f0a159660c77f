"""Time the UNMODIFIED reference (JeLLyFysh, installed copy under baseline/_ref) on a bench workload.

bench.py's `--impl reference` arm and its `cpu_baseline` leg call this: one OS process per host core, every
process builds the reference's object graph from INI text with the reference's own factory
(jellyfysh/run.py:161-200 does the same), runs `mediator.run()` -- the reference's stock single-process event loop --
for a fixed wall-clock budget and counts the iterations whose winner is an interaction / cell / end-of-chain
handler (SURVEY.md 8d "unit of work") by wrapping `Scheduler.get_succeeding_event`, exactly like the survey probe.
Nothing of jellyfysh_b200 or of the oracle is on this path.

Start configuration: the reference's RandomInputHandler asks `setting.random_position()` for every particle; the
runner answers with the bench's jittered-lattice start so that both arms simulate the same system (uniform random
positions at density 0.5 would overlap Lennard-Jones cores). That is the only patch.
"""
import configparser
import contextlib
import io
import multiprocessing
import os
import random
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")
INTERACTION_HANDLERS = ("TwoLeafUnit", "CellVeto", "CellBoundary", "EndOfChain")


def available():
    return os.path.exists(os.path.join(REF_ROOT, "jellyfysh", "run.py"))


class _Stop(Exception):
    pass


def _run_one(args):
    """Worker: never lets a reference exception travel through the pool (the parent could not unpickle it and
    multiprocessing would wait forever); errors come back as text."""
    try:
        return _run_one_unguarded(args)
    except BaseException as error:  # noqa: BLE001 - reported to the parent
        import traceback
        return ("error", "".join(traceback.format_exception_only(type(error), error)).strip())


def _run_one_unguarded(args):
    ini_text, positions, seed, warmup_seconds, budget_seconds, segments = args[:6]
    composites = args[6] if len(args) > 6 else None
    # segment_events: a segment ends after that many events (the bench's "step" of the same workload) instead of after
    # budget_seconds; budget_seconds then bounds the whole run (unfinished segments are dropped)
    segment_events = args[7] if len(args) > 7 else None
    sys.path.insert(0, REF_ROOT)
    import warnings
    warnings.filterwarnings("ignore")
    from jellyfysh.base import factory
    from jellyfysh.base.strings import to_camel_case
    import jellyfysh.setting as setting
    random.seed(seed)
    config = configparser.ConfigParser()
    config.read_string(ini_text)
    t_init = time.perf_counter()
    factory.build_from_config(config, to_camel_case(config.get("Run", "setting")), "jellyfysh.setting")
    if positions is not None:
        iterator = iter([list(map(float, p)) for p in positions])
        setting.random_position = lambda: next(iterator)
    restore = None
    if composites is not None:
        # composite point objects (dipoles, molecules): the start configuration goes through the node creator of the
        # reference's random input handler, like the positions above (tests/golden/configs.py: patch_composite_start)
        sys.path.insert(0, os.path.join(HERE, "..", "tests", "golden"))
        import configs
        restore = configs.patch_composite_start(composites)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            mediator = factory.build_from_config(config, to_camel_case(config.get("Run", "mediator")), "jellyfysh.mediator")
    finally:
        if restore is not None:
            restore()
    init_seconds = time.perf_counter() - t_init
    scheduler = mediator._scheduler
    original = scheduler.get_succeeding_event
    state = {"events": 0, "t0": None, "deadline": None, "done": []}
    started = time.perf_counter()

    def get_succeeding_event_by_events():
        winner = original()
        now = time.perf_counter()
        if state["t0"] is None:
            state["t0"] = now
        if any(tag in type(winner).__name__ for tag in INTERACTION_HANDLERS):
            state["events"] += 1
            if state["events"] >= segment_events:
                state["done"].append((state["events"], now - state["t0"]))
                state["t0"], state["events"] = now, 0
        if len(state["done"]) >= segments or now - started >= budget_seconds:
            raise _Stop()
        return winner

    def get_succeeding_event():
        winner = original()
        now = time.perf_counter()
        if state["t0"] is None and now - started >= warmup_seconds:
            state["t0"] = now
            state["deadline"] = now + budget_seconds
            state["events"] = 0
        if any(tag in type(winner).__name__ for tag in INTERACTION_HANDLERS):
            state["events"] += 1
        if state["deadline"] is not None and now >= state["deadline"]:
            state["done"].append((state["events"], now - state["t0"]))  # one timed segment
            if len(state["done"]) >= segments:
                raise _Stop()
            state["t0"], state["deadline"], state["events"] = now, now + budget_seconds, 0
        return winner

    scheduler.get_succeeding_event = get_succeeding_event_by_events if segment_events else get_succeeding_event
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            mediator.run()
    except _Stop:
        pass
    return state["done"], init_seconds


def run_segments(ini_text, positions_per_process, warmup_seconds=2.0, budget_seconds=10.0, segments=1, composites=None):
    """One chain per entry of positions_per_process in parallel processes, `segments` consecutive timed segments of
    budget_seconds each after the warm-up. Returns ([events/s summed over processes per segment], processes,
    total events, mean init seconds)."""
    jobs = [(ini_text, positions, 1000 + k, warmup_seconds, budget_seconds, segments,
             None if composites is None else composites[k]) for k, positions in enumerate(positions_per_process)]
    context = multiprocessing.get_context("spawn")
    with context.Pool(len(jobs)) as pool:
        results = pool.map(_run_one, jobs)
    failures = [r[1] for r in results if r[0] == "error"]
    if failures:
        raise RuntimeError("the reference failed in %d of %d processes: %s" % (len(failures), len(jobs), failures[0]))
    rates = [sum(done[k][0] / done[k][1] for done, _ in results) for k in range(segments)]
    events = sum(events for done, _ in results for events, _ in done)
    return rates, len(jobs), events, sum(init for _, init in results) / len(results)


def run_event_segments(ini_text, positions_per_process, segment_events, segments, skip, max_seconds, composites=None):
    """One chain per process; every chain runs `segments` consecutive segments of `segment_events` events each (a bench
    step of the same workload), at most max_seconds of wall clock. The first `skip` segments are the warm-up. Returns
    (events/s = sum over processes of timed events / timed seconds, processes, timed events, mean init seconds,
    timed segments completed by the slowest process)."""
    jobs = [(ini_text, positions, 1000 + k, 0.0, max_seconds, segments,
             None if composites is None else composites[k], int(segment_events))
            for k, positions in enumerate(positions_per_process)]
    context = multiprocessing.get_context("spawn")
    with context.Pool(len(jobs)) as pool:
        results = pool.map(_run_one, jobs)
    failures = [r[1] for r in results if r[0] == "error"]
    if failures:
        raise RuntimeError("the reference failed in %d of %d processes: %s" % (len(failures), len(jobs), failures[0]))
    rate, events, completed = 0.0, 0, None
    for done, _ in results:
        timed = done[skip:]
        if not timed:
            raise RuntimeError("the reference did not finish its warm-up segments within %.0f s" % max_seconds)
        rate += sum(e for e, _ in timed) / sum(t for _, t in timed)
        events += sum(e for e, _ in timed)
        completed = len(timed) if completed is None else min(completed, len(timed))
    return rate, len(jobs), events, sum(init for _, init in results) / len(results), completed


def run(ini_text, positions_per_process, warmup_seconds=2.0, budget_seconds=10.0, composites=None):
    """Single timed segment: (events per second summed over processes, processes, events, mean init seconds)."""
    rates, processes, events, init_seconds = run_segments(ini_text, positions_per_process, warmup_seconds,
                                                          budget_seconds, 1, composites)
    return rates[0], processes, events, init_seconds


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(HERE, "..", "tests", "golden"))
    sys.path.insert(0, os.path.join(HERE, ".."))
    import configs
    from jellyfysh_b200 import workloads
    n, cells = 1024, 12
    length = float((n / 0.5) ** (1.0 / 3.0))
    ini = configs.lennard_jones_ini(n, length, cells, chain_time=10.0)
    procs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    positions = workloads.lattice_start(procs, n, cells, length)
    print(run(ini, list(positions), 1.0, float(sys.argv[2]) if len(sys.argv) > 2 else 5.0))
