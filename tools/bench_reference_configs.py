"""The unmodified reference (baseline/_ref, CPython) timed on the other configurations of SURVEY.md 8(d) -- C1, C3 (N = 64)
and C4 (32 molecules) -- next to tools/bench_configs.py's device rates: python tools/bench_reference_configs.py [--out F]

One process per host core, every process one chain through the reference's own factory and SingleProcessMediator
(baseline/reference_runner.py), 2 s warm-up + `--seconds` timed; events are the iterations whose winner is an interaction,
cell or end-of-chain handler. Nothing of jellyfysh_b200 or of the oracle is on this path; tests/golden/configs.py only
supplies the INI text (the shipped files with their sizes changed) and the start configurations."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for path in (ROOT, os.path.join(ROOT, "baseline"), os.path.join(ROOT, "tests", "golden")):
    if path not in sys.path:
        sys.path.insert(0, path)

import numpy as np  # noqa: E402

import configs  # noqa: E402
import reference_runner as rr  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--out")
    parser.add_argument("--seconds", type=float, default=8.0)
    parser.add_argument("--processes", type=int, default=os.cpu_count())
    parser.add_argument("--only", default="")
    args = parser.parse_args()
    ref = rr.REF_ROOT
    procs = args.processes
    runs = []

    def c1():
        roots, leaves = configs.read_pdb_dipoles(ref)
        return ("C1", "shipped hard_disk_dipoles_cells.ini, 81 dipoles, shipped start configuration",
                configs.hard_disk_dipoles_cells_ini(ref), [None] * procs, [(roots, leaves)] * procs)

    def c3():
        n = 64
        cps = [int(np.ceil((2 * n) ** (1.0 / 3.0)))] * 3
        return ("C3 N=64", "Coulomb atoms, cell veto (coulomb_atoms/cell_veto.ini shape), uniform random start",
                configs.coulomb_atoms_ini(n, cps, points_per_side=10),
                [configs.uniform_start(n, 1.0, seed=1000 + k) for k in range(procs)], None)

    def c4():
        n = 32
        starts = [configs.water_start(n, 10.0, seed=k) for k in range(procs)]
        return ("C4 32 molecules", "shipped water/coulomb_cell_veto_lj_inverted.ini sized for 32 molecules, 200 estimator trials",
                configs.water_ini(ref, n_molecules=n, number_trials=200), [None] * procs,
                [(r, l.reshape(n, 3, 3)) for r, l in starts])

    results = []
    for key, make in (("c1", c1), ("c3", c3), ("c4", c4)):
        if args.only and key not in args.only.split(","):
            continue
        name, workload, ini, positions, composites = make()
        rate, processes, events, init_seconds = rr.run(ini, positions, 2.0, args.seconds, composites=composites)
        out = {"config": name, "workload": workload, "impl": "reference", "interpreter": "CPython %d.%d" % sys.version_info[:2],
               "cores": processes, "events_per_sec": rate, "events_per_sec_per_core": rate / processes, "events": events,
               "seconds": args.seconds, "init_seconds": init_seconds}
        print(json.dumps(out), flush=True)
        results.append(out)
    if args.out:
        with open(args.out, "w") as handle:
            json.dump(results, handle, indent=1)


if __name__ == "__main__":
    main()
