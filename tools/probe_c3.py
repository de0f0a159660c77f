"""Development tool: event statistics of the C3 program (Coulomb atoms, N = 64) with the batched kernel: evaluated
candidates per event with and without pruning."""
import sys

sys.path.insert(0, ".")
from jellyfysh_b200 import engine, workloads  # noqa: E402
import numpy as np  # noqa: E402

n_chains, n = 4096, 64
builder, length = workloads.coulomb_atoms(n_particles=n)
positions = workloads.uniform_start(n_chains, n, length)
charges = np.ones((n_chains, n))
for prune in (0, 1):
    with engine.Engine(builder, n_chains=n_chains) as eng:
        eng.set_option(eng.OPTION_PRUNE_CANDIDATES, prune)
        eng.upload_positions(positions, charges)
        eng.start()
        for step in range(3):
            before = eng.kernel_seconds
            eng.run(max_events=1000)
            stats = eng.sync()
            dt = eng.kernel_seconds - before
        print(f"prune {prune}: {stats['events'] / dt:.4g} events/s; evaluated cand/event {stats['candidates'] / stats['events']:.2f} "
              f"targets/event {stats['pair_targets'] / stats['events']:.2f} pair {stats['pair_events']} veto {stats['veto_events']} "
              f"acc {stats['veto_accepted']} bnd {stats['boundary_events']}")
