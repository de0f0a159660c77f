"""Throughput probe of the molecule kernel (water): python tools/probe_water.py [n_molecules] [n_chains] [events]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
import trace_util as tu, configs
from jellyfysh_b200 import engine
from jellyfysh_b200.program import ProgramBuilder

n_mol = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n_chains = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
events = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
g = dict(tu.load_trace("trace_water"))
g["meta_n"] = np.asarray(3 * n_mol)
pb = tu.water_builder_of(g, ProgramBuilder)
roots = np.empty((n_chains, n_mol, 3)); leaves = np.empty((n_chains, 3 * n_mol, 3))
for c in range(n_chains):
    r, l = configs.water_start(n_mol, 10.0, seed=c)
    roots[c], leaves[c] = r, l.reshape(-1, 3)
charges = np.tile([0.41, -0.82, 0.41], (n_chains, n_mol))
with engine.Engine(pb, n_chains=n_chains) as eng:
    eng.upload_positions(leaves, charges); eng.upload_roots(roots); eng.start(first_stream=0)
    for step in range(3):
        t = time.time()
        eng.run(max_events=events)
        stats = eng.sync()
        dt = time.time() - t
        print(f"step {step}: {stats['events']} events in {dt:.3f} s = {stats['events'] / dt:.3e} events/s "
              f"({1e6 * dt / events:.1f} us per event per chain)", {k: v for k, v in stats.items() if v}, flush=True)
    if len(sys.argv) > 4:
        # time-limited runs, like the mediator does between sampling events
        for k in range(1, int(sys.argv[4]) + 1):
            t = time.time()
            eng.run(until=(40.0 * k, 0.0))
            stats = eng.sync()
            st = eng.chain_states()
            print(f"until {40 * k}: {stats['events']} events in {time.time() - t:.3f} s; pending kinds",
                  np.bincount(st["pending_kind"], minlength=9).tolist(), "kept", np.unique(st["kept_kind"]).tolist(), flush=True)
