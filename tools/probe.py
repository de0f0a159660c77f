"""Quick device-time probe of the event kernel (development tool, not the bench)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from jellyfysh_b200 import engine, workloads  # noqa: E402


def main():
    n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    cells = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    events = int(sys.argv[4]) if len(sys.argv) > 4 else 512
    t0 = time.time()
    builder, length = workloads.lennard_jones(n_particles=n, cells_per_side=cells)
    print(f"tables built in {time.time() - t0:.2f} s")
    positions = workloads.lattice_start(n_chains, n, cells, length)
    with engine.Engine(builder, n_chains=n_chains) as eng:
        eng.upload_positions(positions)
        eng.start()
        for step in range(6):
            before = eng.kernel_seconds
            eng.run(max_events=events)
            stats = eng.sync()
            dt = eng.kernel_seconds - before
            print(f"step {step}: {stats['events']} events in {dt * 1e3:.3f} ms -> {stats['events'] / dt:.4g} events/s; "
                  f"cand/event {stats['candidates'] / stats['events']:.2f} pair {stats['pair_events']} veto {stats['veto_events']} "
                  f"acc {stats['veto_accepted']} bnd {stats['boundary_events']} eoc {stats['end_of_chain_events']} "
                  f"viol {stats['bound_violations']}")


if __name__ == "__main__":
    main()
