"""Development tool: follow throughput and surplus-list length while the lattice start equilibrates."""
import sys
import numpy as np
sys.path.insert(0, ".")
from jellyfysh_b200 import engine, workloads

n_chains, n, cells = 4096, 1024, 12
events = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
builder, length = workloads.lennard_jones(n_particles=n, cells_per_side=cells, max_surplus=512)
positions = workloads.lattice_start(n_chains, n, cells, length)
with engine.Engine(builder, n_chains=n_chains) as eng:
    eng.upload_positions(positions)
    eng.start()
    for step in range(steps):
        before = eng.kernel_seconds
        eng.run(max_events=events)
        stats = eng.sync()
        dt = eng.kernel_seconds - before
        _, surplus = eng.cells()
        ns = np.array([len(s) for s in surplus])
        st = eng.chain_states()
        print(f"step {step}: {stats['events'] / dt:.4g} ev/s cand/ev {stats['candidates'] / stats['events']:.2f} "
              f"pair {stats['pair_events'] / stats['events']:.3f} surplus mean {ns.mean():.1f} max {ns.max()} "
              f"time {st['time_q'].mean() + st['time_r'].mean():.2f}", flush=True)
