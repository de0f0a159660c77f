"""Device-resident event rates of the other configurations of SURVEY.md 8(d) (C1, C3, C4, C5): python tools/bench_configs.py

bench.py measures C2, the configuration the metric is quoted on; this tool measures the rest with the same rules
(warm-up launches, then timed launches with CUDA events on the engine's stream = ecmc_kernel_seconds) and writes one
JSON object per configuration to stdout and, with --out, to a file (profiles/r1_configs.json)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for path in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if path not in sys.path:
        sys.path.insert(0, path)

import numpy as np  # noqa: E402

from jellyfysh_b200 import engine, workloads  # noqa: E402
from jellyfysh_b200.program import ProgramBuilder  # noqa: E402


def timed(eng, events, warmup=2, steps=5):
    for _ in range(warmup):
        eng.run(max_events=events)
    eng.sync()
    before, launches = eng.kernel_seconds, eng.kernel_launches
    for _ in range(steps):
        eng.run(max_events=events)
    stats = eng.sync()
    seconds = eng.kernel_seconds - before
    return stats, seconds, eng.kernel_launches - launches


def report(name, workload, n_chains, stats, seconds, launches, extra=None):
    out = {"config": name, "workload": workload, "chains": n_chains, "events": stats["events"],
           "device_seconds": seconds, "launches": launches, "events_per_sec": stats["events"] / seconds,
           "ns_per_event_per_chain": 1e9 * seconds * n_chains / stats["events"],
           "event_mix": {k: v for k, v in stats.items() if v and k != "events"}}
    out.update(extra or {})
    print(json.dumps(out), flush=True)
    return out


def c1_dipoles(n_chains):
    import trace_util as tu
    g = tu.load_trace("trace_hard_disk_dipoles")
    pb = tu.dipole_builder_of(g, ProgramBuilder)
    with engine.Engine(pb, n_chains=n_chains) as eng:
        eng.upload_positions(np.tile(g["positions0"], (n_chains, 1, 1)))
        eng.upload_roots(np.tile(g["roots0"], (n_chains, 1, 1)))
        eng.start(first_stream=0)
        stats, seconds, launches = timed(eng, 4000)
    return report("C1", "shipped hard_disk_dipoles_cells.ini: 81 hard-disk dipoles, 13^2 leaf-level cells, shipped start "
                  "configuration, every chain its own random stream", n_chains, stats, seconds, launches)


def c3_coulomb(n, n_chains):
    builder, length = workloads.coulomb_atoms(n_particles=n)
    start = workloads.uniform_start(n_chains, n, length)
    with engine.Engine(builder, n_chains=n_chains) as eng:
        eng.upload_positions(start, np.ones((n_chains, n)))
        eng.start(first_stream=0)
        stats, seconds, launches = timed(eng, 1000)
    return report("C3 N=%d" % n, "Coulomb atoms (merged-image Coulomb, inverse-power bound for nearby cells, cell veto), "
                  "L = 1, beta = 2, uniform random start", n_chains, stats, seconds, launches)


def c4_water(n_molecules, n_chains):
    import configs
    import trace_util as tu
    g = dict(tu.load_trace("trace_water"))
    g["meta_n"] = np.asarray(3 * n_molecules)
    pb = tu.water_builder_of(g, ProgramBuilder)
    roots = np.empty((n_chains, n_molecules, 3))
    leaves = np.empty((n_chains, 3 * n_molecules, 3))
    for c in range(n_chains):
        r, l = configs.water_start(n_molecules, 10.0, seed=c)
        roots[c], leaves[c] = r, l.reshape(-1, 3)
    with engine.Engine(pb, n_chains=n_chains) as eng:
        eng.upload_positions(leaves, np.tile([0.41, -0.82, 0.41], (n_chains, n_molecules)))
        eng.upload_roots(roots)
        eng.start(first_stream=0)
        stats, seconds, launches = timed(eng, 500)
    return report("C4 %d molecules" % n_molecules, "SPC/Fw water (water/coulomb_cell_veto_lj_inverted.ini), L = 10, cells 6^3 "
                  "with two neighbour layers, cell-veto tables of the reference (200 estimator trials)", n_chains, stats,
                  seconds, launches)


def c5_single_chain():
    n, cells = 65536, 48
    builder, length = workloads.lennard_jones(n_particles=n, cells_per_side=cells)
    start = workloads.lattice_start(1, n, cells, length)
    with engine.Engine(builder, n_chains=1) as eng:
        eng.upload_positions(start)
        eng.start(first_stream=0)
        stats, seconds, launches = timed(eng, 200000, warmup=1, steps=3)
    return report("C5", "single Lennard-Jones chain, N = 65536, cells 48^3 (latency of one warp)", 1, stats, seconds,
                  launches, {"ns_per_event": 1e9 * seconds / stats["events"]})


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--out")
    parser.add_argument("--only", default="")
    args = parser.parse_args()
    runs = [("c1", lambda: c1_dipoles(4096)), ("c3_64", lambda: c3_coulomb(64, 4096)),
            ("c3_128", lambda: c3_coulomb(128, 4096)), ("c3_256", lambda: c3_coulomb(256, 2048)),
            ("c3_512", lambda: c3_coulomb(512, 1024)), ("c4_2", lambda: c4_water(2, 4096)),
            ("c4_32", lambda: c4_water(32, 1024)), ("c5", c5_single_chain)]
    results = [run() for name, run in runs if not args.only or name in args.only.split(",")]
    if args.out:
        with open(args.out, "w") as handle:
            json.dump(results, handle, indent=1)


if __name__ == "__main__":
    main()
